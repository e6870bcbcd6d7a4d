"""Training-mode statements of the dynamics processors and the STFT reverb.

The fused CUDA kernels of these processors (csrc/dynamics.cu, csrc/reverb.cu) are forward-only.  When autograd expects
a gradient from them, the same mathematics is evaluated here as the reference formulates it -- parameter-side and
memoryless steps as PyTorch expressions on the device, every O(samples x taps) convolution on the differentiable FIR
engine of this package (`functional.fir_conv` -> `autograd.FirConvCausalFn`: both gradients are causal convolutions on
csrc/fir.cu) -- so that a rendered graph EQ -> Compressor -> Reverb trains end to end.  Slower than the fused forward
(several passes over the audio instead of one); the no-grad path never comes here.

  Compressor / NoiseGate   processors/dynamics.py:361-419, 443-489, 598-651, 675-721
  TruncatedOnePoleIIRFilter processors/core/envelope.py:34-60 (an FIR of iir_len taps + relu, exactly as upstream)
  STFTMaskedNoiseReverb     processors/reverb.py:161-228
Attack / release ballistics has no statement here (upstream differentiates it inside torchcomp): it raises.
"""
from __future__ import annotations

import torch
import torch.nn.functional as tF

from . import functional as F_


def _one_pole(u: torch.Tensor, z_alpha: torch.Tensor, iir_len: int) -> torch.Tensor:
    """relu(u * h), h[n] = (1 - a) a^n, n < iir_len, a = min(sigmoid(z), 1 - 1e-5); u [B, L], z [B, 1]."""
    alpha = torch.clamp(torch.sigmoid(z_alpha.reshape(-1, 1)), max=1 - 1e-5)
    n = torch.arange(iir_len, device=u.device, dtype=torch.float32)[None, :]
    h = (1 - alpha) * torch.exp(n * torch.log(alpha))
    return torch.relu(F_.fir_conv(u, h, "causal"))


def _log_gain(kind: str, knee: str, G, T, log_ratio, log_knee):
    """G_out - G of the knee (T = log_threshold - 6 already)."""
    if kind == "compressor":
        slope = 1 / (1 + torch.exp(log_ratio)) - 1                      # 1/R - 1
        if knee == "hard":
            return torch.minimum(torch.zeros_like(G), (G - T) * slope)
        if knee == "quadratic":
            W = torch.exp(log_knee) / 2
            mid = slope * (G - T + W).square() / (4 * W)
            return torch.where(G > T + W, (G - T) * slope, torch.where(G < T - W, torch.zeros_like(G), mid))
        if knee == "exponential":
            W = torch.exp(log_knee)
            return slope * tF.softplus(W * (G - T)) / W
    else:
        if knee == "hard":
            return torch.minimum(torch.zeros_like(G), torch.exp(log_ratio) * (G - T))   # (R - 1)(G - T), R = 1 + e^lr
        if knee == "quadratic":
            W = torch.exp(log_knee) / 2
            mid = -torch.exp(log_ratio) * (G - T - W).square() / (4 * W)                # (1 - R)(...)
            return torch.where(G < T - W, torch.exp(log_ratio) * (G - T), torch.where(G > T + W, torch.zeros_like(G), mid))
        if knee == "exponential":
            W = torch.exp(log_knee)
            return -torch.exp(log_ratio) * tF.softplus(W * (T - G)) / W
        if knee == "approx_gate":  # ApproxNoiseGate.compute_gain as shipped (dynamics.py:186-204)
            R = torch.exp(log_ratio)
            W = torch.exp(log_knee)
            mid = (1 - R) * (G - T - W / 2).square() / 2 / (W + 1e-3)
            return torch.where(G < T - W / 2, (R - 1) * (G - T), torch.where(G > T + W / 2, torch.zeros_like(G), mid))
    raise ValueError(f"Unknown knee: {knee}")


def dynamics_chain(x: torch.Tensor, stages: list, iir_len: int) -> torch.Tensor:
    """Differentiable statement of functional.dynamics_chain for stages whose smoothers are one-pole or absent."""
    for st in stages:
        if "ballistics" in (st.get("energy_smoother", "iir"), st.get("gain_smoother")):
            raise NotImplementedError("attack / release ballistics has no backward pass here (forward-only kernels); "
                                      "use energy_smoother='iir' for training or wrap the call in torch.no_grad()")
    y = x.to(torch.float32)
    col = lambda t: None if t is None else t.reshape(t.shape[0], -1)[:, :1].to(torch.float32)  # noqa: E731
    for st in stages:
        energy = y.square().mean(-2)
        if st.get("energy_smoother", "iir") == "iir":
            energy = _one_pole(energy, st["z_alpha_pre"], iir_len)
        G = torch.log(energy + 1e-5)
        lg = _log_gain(st["kind"], st.get("knee", "quadratic"), G, col(st["log_threshold"]) - 6, col(st["log_ratio"]),
                       col(st.get("log_knee")))
        if st.get("gain_smoother") is None:
            gain = torch.exp(lg)
        elif st.get("gain_smooth_in_log", False):
            gain = torch.exp(_one_pole(lg, st["z_alpha_post"], iir_len))
        else:
            gain = _one_pole(torch.exp(lg), st["z_alpha_post"], iir_len)
        y = gain[:, None, :] * y
    return y


def stft_reverb_ir(noise_stft: torch.Tensor, init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude, window,
                   ir_len: int, n_fft: int, hop_length: int) -> torch.Tensor:
    """compute_ir (reverb.py:161-200) as upstream states it: mask = exp((H0 - softplus(Hd) m [+ G_env]) / 8),
    istft(noise_stft * mask).  Returns the un-normalised mid/side response [B, 2, ir_len] with its autograd graph."""
    B = init_log_magnitude.shape[0]
    frames = 1 + ir_len // hop_length
    m = torch.arange(frames, device=init_log_magnitude.device, dtype=torch.float32).view(1, 1, 1, -1)
    logmag = init_log_magnitude.unsqueeze(-1) - tF.softplus(delta_log_magnitude).unsqueeze(-1) * m
    if gain_env_log_magnitude is not None:
        logmag = logmag + gain_env_log_magnitude[:, :, None, :]
    spec = noise_stft.to(torch.complex64) * torch.exp(logmag / 8)
    spec = spec.reshape(B * 2, spec.shape[-2], spec.shape[-1])
    ir = torch.istft(spec, n_fft=n_fft, hop_length=hop_length, window=window, length=ir_len)
    return ir.reshape(B, 2, ir_len)


# ------------------------------------------------------------------ memoryless processors (functional.pointwise ops)
def pointwise(op: str, x, p0=None, p1=None, p2=None, p3=None, dc=None, order: int = 0, flags: int = 0):
    """Differentiable statements of the ops of gfx_pointwise_f32 (include/grafx_b200.h lists the parameter meaning):
    processors/stereo.py:71-84, processors/nonlinear.py:64-89, 159-205, 268-285, 349-384."""
    x = x.to(torch.float32)
    col = lambda t: t.reshape(t.shape[0], -1)  # noqa: E731
    if dc is not None:
        x = x - dc.reshape(x.shape[0], x.shape[1], 1)

    def pre_post(log_pre, log_post, core):
        if log_pre is None:
            return core(x)
        pre = torch.exp(col(log_pre))[:, :1, None]
        y = core(x * pre)
        return y / pre if (flags & 1) else y * torch.exp(col(log_post))[:, :1, None]

    if op == "side_gain":
        left, right = x[:, 0], x[:, 1]
        mid, side = left + right, torch.exp(col(p0)[:, :1]) * (left - right)
        return torch.stack([(mid + side) / 2, (mid - side) / 2], 1)
    if op == "tanh":
        if p2 is None:
            return pre_post(p0, p1, torch.tanh)
        b = col(p2)[:, :1, None]
        return pre_post(p0, p1, lambda v: torch.tanh(v + b) - torch.tanh(b))
    if op == "piecewise_tanh":
        hard, thr = torch.exp(col(p0)), torch.sigmoid(col(p1))
        kn, kp = thr[:, 0:1, None], thr[:, 1:2, None]
        gp, gn = hard[:, 0:1, None], hard[:, 1:2, None]
        ap, an = (1 - torch.tanh(kp)) / gp, (1 - torch.tanh(kn)) / gn

        def core(v):
            return torch.where(v > kp, ap * torch.tanh(gp * (v - kp)) + torch.tanh(kp),
                               torch.where(v < -kn, an * torch.tanh(gn * (v + kn)) - torch.tanh(kn), torch.tanh(v)))

        return pre_post(p2, p3, core)
    if op in ("power", "chebyshev"):
        v = x if p1 is None else x * torch.exp(col(p1))[:, :1, None]
        w = torch.tanh(col(p0))
        basis = [torch.ones_like(v), v]
        for k in range(2, order):
            basis.append(basis[-1] * v if op == "power" else 2 * v * basis[-1] - basis[-2])
        y = torch.zeros_like(v)
        for k in range(order):
            y = y + w[:, k].view(-1, 1, 1) * (torch.tanh(basis[k]) if (flags & 2) else basis[k])
        return y
    raise NotImplementedError(f"pointwise op {op} has no training-mode statement")
