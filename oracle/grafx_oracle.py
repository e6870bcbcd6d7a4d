"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the grafx hot path.

A restatement, on CPU, of the algorithms of the reference (sh-lee97/grafx @ 474e5dc,
/root/reference/src/grafx) for the path named by BASELINE.json.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import it;
the product (grafx_b200/) never does.

Pinning: the reference's tests hold no golden vectors; the only
known-answer test is ssm == lfilter (tests/processors/test_filter.py:215-233).  This oracle
is therefore pinned by executing the reference's own code in the build container
(oracle/ref_loader.py, oracle/make_golden.py) and comparing: see tests/test_oracle_golden.py
and the fixtures in tests/golden/.  One third-party recurrence is not in the reference tree nor in this image:
`ballistics` = torchcomp.compressor_core.  It is pinned to the upstream kernel text as restated in
oracle/torchcomp_core.py (bit-exact agreement in float64) and by reference-side invariants (equal coefficients ==
the linear one-pole; single-branch inputs): tests/test_oracle_golden.py::test_ballistics_*.  It has NOT been compared
with an installed torchcomp (none is available offline).

Everything is dtype-generic torch code (run it in float64 for a ground truth, in float32 to
mimic the reference's arithmetic) plus scipy/numpy for the sequential loops.  The exact IIR
uses torchaudio.functional.lfilter exactly like the reference does when `use_torchaudio=True`
(that is what the CPU baseline times), or scipy.signal.lfilter (independent second opinion).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

PI = math.pi


# ------------------------------------------------------------------ core/midside.py:4-17
def ms_to_lr(x):
    mid, side = torch.split(x, (1, 1), -2)
    return torch.cat([mid + side, mid - side], -2)


def lr_to_ms(x, mult=0.5):
    left, right = torch.split(x, (1, 1), -2)
    x = torch.cat([left + right, left - right], -2)
    return x * mult if mult is not None else x


# ------------------------------------------------------------------ core/utils.py:14-18
def normalize_impulse(ir, eps=1e-12):
    e = ir.square().sum(2, keepdim=True).mean(1, keepdim=True)
    return ir / torch.sqrt(e + eps)


# ------------------------------------------------------------------ core/convolution.py:119-134
def convolve(x, h, mode="causal"):
    """Linear convolution along the last axis, broadcasting the leading axes, sliced to
    len(x).  This is the *intended* semantics (docstring convolution.py:22-25): the shipped
    code omits n= in irfft and is only correct for even Lx+Lh-1; the pad
    length is rounded up to even here, exactly like oracle/ref_loader.py's even-pad guard."""
    lx, lh = x.shape[-1], h.shape[-1]
    n = lx + lh - 1
    n += n & 1
    X = torch.fft.rfft(F.pad(x, (0, n - lx)))
    H = torch.fft.rfft(F.pad(h, (0, n - lh)))
    y = torch.fft.irfft(X * H, n=n)
    if mode == "zerophase":
        return y[..., lh // 2 : lh // 2 + lx]
    if mode == "causal":
        return y[..., :lx]
    return y[..., : lx + lh - 1]


def convolve_direct(x, h, mode="causal"):
    """Independent O(L*N) time-domain convolution (float64 numpy) for small cases."""
    xs = np.broadcast_to(x.double().numpy(), np.broadcast_shapes(x.shape[:-1], h.shape[:-1]) + (x.shape[-1],))
    hs = np.broadcast_to(h.double().numpy(), xs.shape[:-1] + (h.shape[-1],))
    lx, lh = x.shape[-1], h.shape[-1]
    out = np.zeros(xs.shape[:-1] + (lx + lh - 1,))
    for idx in np.ndindex(*xs.shape[:-1]):
        out[idx] = np.convolve(xs[idx], hs[idx])
    out = torch.from_numpy(out)
    if mode == "zerophase":
        return out[..., lh // 2 : lh // 2 + lx]
    return out[..., :lx]


# ------------------------------------------------------------------ core/iir.py:154-184
def _broadcast_channels(x, Bs, As):
    b, c_sig, L = x.shape
    c_filt = Bs.shape[1]
    if c_sig == 1 and c_filt > 1:
        x = x.repeat(1, c_filt, 1)
    elif c_sig > 1 and c_filt == 1:
        Bs = Bs.repeat(1, c_sig, 1, 1)
        As = As.repeat(1, c_sig, 1, 1)
    else:
        assert c_sig == c_filt
    return x, Bs, As


def iir_lfilter(x, Bs, As, use_torchaudio=True):
    """Exact cascade of K DF-I biquads, each normalised by its own a0, zero state, no clamp."""
    x, Bs, As = _broadcast_channels(x, Bs, As)
    b, c, L = x.shape
    K = Bs.shape[2]
    y = x.reshape(b * c, L)
    Bf, Af = Bs.reshape(b * c, K, 3), As.reshape(b * c, K, 3)
    if use_torchaudio:
        from torchaudio.functional import lfilter

        for i in range(K):
            y = lfilter(y, b_coeffs=Bf[:, i], a_coeffs=Af[:, i], batching=True, clamp=False)
    else:
        import scipy.signal

        yn = y.double().numpy().copy()
        Bn, An = Bf.double().numpy(), Af.double().numpy()
        for r in range(b * c):
            for i in range(K):
                yn[r] = scipy.signal.lfilter(Bn[r, i] / An[r, i, 0], An[r, i] / An[r, i, 0], yn[r])
        y = torch.from_numpy(yn).to(x.dtype)
    return y.reshape(b, c, L)


# ------------------------------------------------------------------ core/iir.py:147-152,263-276
def iir_fsm_fir(Bs, As, fir_len):
    """Frequency-sampled FIR of the cascade: irfft_N( prod_k B_k(w)/A_k(w) )."""
    k = torch.arange(fir_len // 2 + 1, device=Bs.device)
    d = torch.arange(3, device=Bs.device)
    phase = d[:, None] * k[None, :] / fir_len * 2 * np.pi            # [3, N/2+1]
    delays = torch.exp(-1j * phase).to(torch.complex128 if Bs.dtype == torch.float64 else torch.complex64)
    num = torch.sum(Bs.unsqueeze(-1) * delays, -2)
    den = torch.sum(As.unsqueeze(-1) * delays, -2)
    H = (num / den).prod(-2)
    return torch.fft.irfft(H, dim=-1, n=fir_len)


def iir_fsm(x, Bs, As, fir_len=4000):
    return convolve(x, iir_fsm_fir(Bs, As, fir_len), mode="causal")


# ------------------------------------------------------------------ filter.py coefficient designers
def biquad_filter_coeffs(Bs, A1_pre, A2_pre, A0=None, normalized=False):
    """filter.py:144-154."""
    a1 = 2 * torch.tanh(A1_pre)
    a1_abs = a1.abs()
    a2 = ((2 - a1_abs) * torch.tanh(A2_pre) + a1_abs) / 2
    As = torch.stack([torch.ones_like(A1_pre), a1, a2], -1)
    if normalized:
        As = As * A0.unsqueeze(-1)
    B0 = Bs[:, :, :1]
    Bs = torch.cat([B0 + torch.ones_like(B0), Bs[:, :, 1:]], -1)
    return Bs.unsqueeze(1), As.unsqueeze(1)


def peq_activations(w0, q_inv, log_gain=None):
    """filter.py:592-604 (and 373-383)."""
    w0 = PI * torch.sigmoid(w0)
    q_inv = torch.exp(q_inv)
    cos_w0 = torch.cos(w0)
    alpha = torch.sin(w0) * q_inv * 0.5
    A = torch.exp(log_gain) if log_gain is not None else None
    return cos_w0, alpha, A


def peaking_coeffs(c, alpha, A):
    """filter.py:645-656."""
    Bs = torch.stack([1 + alpha * A, -2 * c, 1 - alpha * A], -1)
    As = torch.stack([1 + alpha / A, -2 * c, 1 - alpha / A], -1)
    return Bs, As


def lowshelf_coeffs(c, alpha, A):
    """filter.py:687-705."""
    S = 2 * A.sqrt() * alpha
    b0 = A * ((A + 1) - (A - 1) * c + S)
    b1 = 2 * A * ((A - 1) - (A + 1) * c)
    b2 = A * ((A + 1) - (A - 1) * c - S)
    a0 = (A + 1) + (A - 1) * c + S
    a1 = -2 * ((A - 1) + (A + 1) * c)
    a2 = (A + 1) + (A - 1) * c - S
    return torch.stack([b0, b1, b2], -1), torch.stack([a0, a1, a2], -1)


def highshelf_coeffs(c, alpha, A):
    """filter.py:736-754."""
    S = 2 * A.sqrt() * alpha
    b0 = A * ((A + 1) + (A - 1) * c + S)
    b1 = -2 * A * ((A - 1) + (A + 1) * c)
    b2 = A * ((A + 1) + (A - 1) * c - S)
    a0 = (A + 1) - (A - 1) * c + S
    a1 = 2 * ((A - 1) - (A + 1) * c)
    a2 = (A + 1) - (A - 1) * c - S
    return torch.stack([b0, b1, b2], -1), torch.stack([a0, a1, a2], -1)


def simple_filter_coeffs(kind, c, alpha):
    """filter.py:416-556 (signs as shipped, e.g. the low-pass numerator is (cos w0 - 1)/2)."""
    a = torch.stack([1 + alpha, -2 * c, 1 - alpha], -1)
    if kind == "lowpass":
        b = torch.stack([(c - 1) / 2, c - 1, (c - 1) / 2], -1)
    elif kind == "highpass":
        b = torch.stack([(1 + c) / 2, -(1 + c), (1 + c) / 2], -1)
    elif kind == "bandpass":
        b = torch.stack([alpha, torch.zeros_like(alpha), -alpha], -1)
    elif kind == "bandreject":
        b = torch.stack([torch.ones_like(c), -2 * c, torch.ones_like(c)], -1)
    elif kind == "allpass":
        b = torch.stack([1 - alpha, -2 * c, 1 + alpha], -1)
    else:
        raise ValueError(kind)
    return b, a


def svf_coeffs(twoR, G, c_hp, c_bp, c_lp):
    """filter.py:303-338."""
    G = torch.tan(PI / 2 * torch.sigmoid(G))
    twoR = F.softplus(twoR) / math.log(2) + 1e-2
    G2 = G.square()
    b = torch.stack([c_hp + c_bp * G + c_lp * G2, -2 * c_hp + 2 * c_lp * G2, c_hp - c_bp * G + c_lp * G2], -1)
    a = torch.stack([1 + G2 + twoR * G, 2 * G2 - 2, 1 + G2 - twoR * G], -1)
    return b, a


def polezero_coeffs(poles, zeros):
    """filter.py:218-238."""
    poles = torch.view_as_complex(poles.contiguous())
    pr = poles.abs()
    poles = poles * torch.tanh(pr) / (pr + 1e-5)
    zeros = torch.view_as_complex(zeros.contiguous())
    zr = zeros.abs()
    ones = torch.ones_like(pr)
    # NB (as shipped): a2 uses the radius BEFORE the tanh re-parameterisation.
    Bs = torch.stack([ones, -2 * zeros.real, zr.square()], -1)
    As = torch.stack([ones, -2 * poles.real, pr.square()], -1)
    return Bs, As


def peq_coeffs(w0, q_inv, log_gain, use_shelving_filters=True):
    """eq.py:290-314.  Inputs [B, n_ch, K] -> Bs, As [B, n_ch, K, 3]."""
    c, alpha, A = peq_activations(w0, q_inv, log_gain)
    K = w0.shape[-1]
    if not use_shelving_filters:
        return peaking_coeffs(c, alpha, A)
    sp = [1, K - 2, 1]
    c_ls, c_pk, c_hs = torch.split(c, sp, 2)
    al_ls, al_pk, al_hs = torch.split(alpha, sp, 2)
    A_ls, A_pk, A_hs = torch.split(A, sp, 2)
    Bl, Al = lowshelf_coeffs(c_ls, al_ls, A_ls)
    Bp, Ap = peaking_coeffs(c_pk, al_pk, A_pk)
    Bh, Ah = highshelf_coeffs(c_hs, al_hs, A_hs)
    return torch.cat([Bl, Bp, Bh], 2), torch.cat([Al, Ap, Ah], 2)


def _iir(x, Bs, As, backend, fsm_fir_len, use_torchaudio=True):
    if backend == "fsm":
        return iir_fsm(x, Bs, As, fsm_fir_len)
    return iir_lfilter(x, Bs, As, use_torchaudio=use_torchaudio)


def parametric_equalizer(x, w0, q_inv, log_gain, processor_channel="mono", use_shelving_filters=True,
                         backend="lfilter", fsm_fir_len=4000, use_torchaudio=True):
    """eq.py:273-322."""
    Bs, As = peq_coeffs(w0, q_inv, log_gain, use_shelving_filters)
    if processor_channel == "midside":
        return ms_to_lr(_iir(lr_to_ms(x), Bs, As, backend, fsm_fir_len, use_torchaudio))
    return _iir(x, Bs, As, backend, fsm_fir_len, use_torchaudio)


def biquad_filter(x, Bs, A1_pre, A2_pre, A0=None, normalized=False, backend="lfilter", fsm_fir_len=4000,
                  use_torchaudio=True):
    Bs, As = biquad_filter_coeffs(Bs, A1_pre, A2_pre, A0, normalized)
    return _iir(x, Bs, As, backend, fsm_fir_len, use_torchaudio)


# ------------------------------------------------------------------ filter.py:52-77 (FIRFilter, ctor broken upstream: R3)
def fir_filter(x, fir, processor_channel="mono"):
    fir = normalize_impulse(torch.tanh(fir))
    if processor_channel == "midside":
        return ms_to_lr(convolve(lr_to_ms(x), fir, "causal"))
    return convolve(x, fir, "causal")


# ------------------------------------------------------------------ core/envelope.py:34-60
def one_pole_alpha(z_alpha):
    return torch.clamp(torch.sigmoid(z_alpha), max=1 - 1e-5)


def truncated_one_pole(u, z_alpha, iir_len=16384):
    """relu(causal_conv(u, (1-a) a^n, n < iir_len)); u [B, L], z_alpha [B, 1]."""
    alpha = one_pole_alpha(z_alpha)
    n = torch.arange(iir_len, device=u.device)[None, :]
    h = (1 - alpha) * torch.exp(n * torch.log(alpha))
    return F.relu(convolve(u, h, "causal"))


def truncated_one_pole_recursive(u, z_alpha, iir_len=16384):
    """Same quantity via the exact recursion and the tail identity
    y_trunc[n] = y[n] - a^N y[n-N]  (float64 numpy; independent of the FFT path)."""
    import scipy.signal

    alpha = one_pole_alpha(z_alpha).double().numpy()[:, 0]
    un = u.double().numpy()
    out = np.empty_like(un)
    for r in range(un.shape[0]):
        y = scipy.signal.lfilter([1 - alpha[r]], [1, -alpha[r]], un[r])
        tail = np.zeros_like(y)
        if iir_len < y.shape[0]:
            tail[iir_len:] = y[:-iir_len]
        out[r] = np.maximum(y - alpha[r] ** iir_len * tail, 0.0)
    return torch.from_numpy(out).to(u.dtype)


# ------------------------------------------------------------------ core/envelope.py:84-101
def ballistics(u, z_alpha):
    """Ballistics.forward (core/envelope.py:84-101) = torchcomp.compressor_core(u, zi=1, at, rt), whose kernel
    (torchcomp/core.py: compressor_kernel, restated in oracle/torchcomp_core.py) is
        at, rt = sigmoid(z[:,0]), sigmoid(z[:,1]);  y[-1] = 1
        c = at if u[t] < y[t-1] else rt;  y[t] = (1 - c) y[t-1] + c u[t]
    Sequential; evaluated in the dtype of u with numpy.  Pins: tests/test_oracle_golden.py::test_ballistics_*."""
    ts = torch.sigmoid(z_alpha)
    at, rt = ts[..., 0].numpy(), ts[..., 1].numpy()
    un = u.numpy()
    y = np.empty_like(un)
    one = un.dtype.type(1)
    prev = np.ones(un.shape[0], dtype=un.dtype)
    for t in range(un.shape[1]):
        x_t = un[:, t]
        c = np.where(x_t < prev, at, rt).astype(un.dtype)
        prev = (one - c) * prev + c * x_t
        y[:, t] = prev
    return torch.from_numpy(y)


def envelope_follower(x, z_alpha, smoother="iir", detect_with="energy", iir_len=16384):
    """BaseEnvelopeFollower.forward (dynamics.py:745-767): detector over the channel axis -> smoother ->
    log(envelope + 1e-5).  x [B, C, L] -> [B, L]."""
    loud = x.square().mean(-2) if detect_with == "energy" else x.abs().mean(-2)
    env = truncated_one_pole(loud, z_alpha, iir_len) if smoother == "iir" else ballistics(loud, z_alpha)
    return torch.log(env + 1e-5)


# ------------------------------------------------------------------ dynamics.py:361-419,443-489 / 598-651,675-721
def _knee_gain(kind, knee, G, T, log_ratio, log_knee):
    if kind == "compressor":
        ratio = 1 + torch.exp(log_ratio)
        if knee == "hard":
            return torch.minimum(G, T + (G - T) / ratio) - G
        if knee == "quadratic":
            W = torch.exp(log_knee) / 2
            below_m = G < (T - W)
            above_m = G > (T + W)
            mid_m = ~below_m & ~above_m
            above = T + (G - T) / ratio
            mid = G + (1 / ratio - 1) * (G - T + W).square() / (4 * W)
            return G * below_m + above * above_m + mid * mid_m - G
        if knee == "exponential":
            W = torch.exp(log_knee)
            return (1 / ratio - 1) * F.softplus(W * (G - T)) / W
    else:
        if knee == "hard":
            ratio = 1 + torch.exp(log_ratio)
            return torch.minimum(G, ratio * (G - T) + T) - G
        if knee == "quadratic":
            ratio = 1 + torch.exp(log_ratio)
            W = torch.exp(log_knee) / 2
            below_m = G < (T - W)
            above_m = G > (T + W)
            mid_m = ~below_m & ~above_m
            below = ratio * (G - T) + T
            mid = G + (1 - ratio) * (G - T - W).square() / (4 * W)
            return below * below_m + G * above_m + mid * mid_m - G
        if knee == "exponential":
            W = torch.exp(log_knee)
            return -torch.exp(log_ratio) * F.softplus(W * (T - G)) / W
        if knee == "approx_gate":
            # ApproxNoiseGate.compute_gain as shipped (dynamics.py:186-204): the ratio has no "+1" and the knee
            # denominator is 2 (W + 1e-3) with the FULL knee width W
            ratio = torch.exp(log_ratio)
            W = torch.exp(log_knee)
            below_m = G < (T - W / 2)
            above_m = G > (T + W / 2)
            mid_m = ~below_m & ~above_m
            below = ratio * (G - T) + T
            mid = G + (1 - ratio) * (G - T - W / 2).square() / 2 / (W + 1e-3)
            return below * below_m + G * above_m + mid * mid_m - G
    raise ValueError(knee)


def _smooth(kind, u, z_alpha, iir_len):
    if kind == "iir":
        return truncated_one_pole(u, z_alpha, iir_len)
    if kind == "ballistics":
        return ballistics(u, z_alpha)
    raise ValueError(kind)


def dynamics(kind, x, log_threshold, log_ratio, log_knee=None, z_alpha_pre=None, z_alpha_post=None,
             energy_smoother="iir", gain_smoother=None, gain_smooth_in_log=False, knee="quadratic",
             iir_len=16384):
    """Compressor (kind="compressor") / NoiseGate (kind="noisegate") forward."""
    energy = x.square().mean(-2)
    if energy_smoother is not None:
        energy = _smooth(energy_smoother, energy, z_alpha_pre, iir_len)
    G = torch.log(energy + 1e-5)
    lg = _knee_gain(kind, knee, G, log_threshold - 6, log_ratio, log_knee)
    if gain_smoother is None:
        gain = torch.exp(lg)
    elif gain_smooth_in_log:
        gain = torch.exp(_smooth(gain_smoother, lg, z_alpha_post, iir_len))
    else:
        gain = _smooth(gain_smoother, torch.exp(lg), z_alpha_post, iir_len)
    return gain[:, None, :] * x


def approx_compressor(x, z_alpha, log_threshold, log_ratio, log_knee, iir_len=16384):
    """ApproxCompressor.forward (dynamics.py:86-110): IIREnvelopeFollower (:753-786) + Compressor.gain_quad_knee."""
    return dynamics("compressor", x, log_threshold, log_ratio, log_knee, z_alpha_pre=z_alpha, iir_len=iir_len)


def approx_noisegate(x, z_alpha, log_threshold, log_ratio, log_knee, iir_len=16384):
    """ApproxNoiseGate.forward (dynamics.py:161-184) with its own compute_gain (:186-204)."""
    return dynamics("noisegate", x, log_threshold, log_ratio, log_knee, z_alpha_pre=z_alpha, knee="approx_gate",
                    iir_len=iir_len)


def rms_difference(X, Y, eps=1e-7):
    """processors/core/utils.py:7-11 (the GainStagingRegularization term, container.py:289)."""
    return (torch.log(X.square().mean((-1, -2)) + eps) - torch.log(Y.square().mean((-1, -2)) + eps)).abs().sum()


def compressor(x, **kw):
    return dynamics("compressor", x, **kw)


def noisegate(x, **kw):
    return dynamics("noisegate", x, **kw)


# ------------------------------------------------------------------ reverb.py:101-114,161-200,225-228
def reverb_noise_stft(ir_len, n_fft=384, hop=192):
    rng = np.random.RandomState(0)
    noise = torch.tensor(rng.uniform(size=(2, ir_len)) * 2 - 1).float()
    return torch.stft(noise, n_fft=n_fft, hop_length=hop, window=torch.hann_window(n_fft), return_complex=True)[None]


def istft_restated(spec, n_fft, hop, window, length):
    """torch.istft(center=True) spelled out: per-frame irfft * window, overlap-add, divide by the
    window-square envelope, drop n_fft//2 samples, keep `length`."""
    frames = torch.fft.irfft(spec, n=n_fft, dim=-2) * window[:, None]          # [..., n_fft, T]
    T = spec.shape[-1]
    total = n_fft + hop * (T - 1)
    out = frames.new_zeros(spec.shape[:-2] + (total,))
    env = frames.new_zeros(total)
    w2 = window.square()
    for t in range(T):
        out[..., t * hop : t * hop + n_fft] += frames[..., :, t]
        env[t * hop : t * hop + n_fft] += w2
    s = n_fft // 2
    return out[..., s : s + length] / env[s : s + length]


def reverb_ir(init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude=None, ir_len=60000, n_fft=384,
              hop=192, noise_stft=None, use_torch_istft=True):
    """compute_ir (reverb.py:161-200): [B,2,bins] x2 (+[B,2,frames]) -> ir [B,2,ir_len] (mid/side)."""
    dt = init_log_magnitude.dtype
    frames = 1 + ir_len // hop
    if noise_stft is None:
        noise_stft = reverb_noise_stft(ir_len, n_fft, hop)
    noise_stft = noise_stft.to(torch.complex128 if dt == torch.float64 else torch.complex64)
    m = torch.arange(frames).view(1, 1, 1, -1)
    logmag = init_log_magnitude[:, :, :, None] - F.softplus(delta_log_magnitude)[:, :, :, None] * m
    if gain_env_log_magnitude is not None:
        logmag = logmag + gain_env_log_magnitude[:, :, None, :]
    spec = noise_stft * torch.exp(logmag / 8)
    B = spec.shape[0]
    spec = spec.reshape(B * 2, spec.shape[2], spec.shape[3])
    window = torch.hann_window(n_fft).to(dt)   # fp32-computed then cast, as the registered buffer is
    if use_torch_istft:
        ir = torch.istft(spec, n_fft=n_fft, hop_length=hop, window=window, length=ir_len)
    else:
        ir = istft_restated(spec, n_fft, hop, window, ir_len)
    return ir.reshape(B, 2, ir_len)


def stft_masked_noise_reverb(x, init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude=None,
                             ir_len=60000, processor_channel="pseudo_midside", n_fft=384, hop=192,
                             use_torch_istft=True):
    ir = reverb_ir(init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude, ir_len, n_fft, hop,
                   use_torch_istft=use_torch_istft)
    if processor_channel == "pseudo_midside":
        return convolve(x, normalize_impulse(ms_to_lr(ir)), "causal")
    if processor_channel == "midside":
        return ms_to_lr(convolve(lr_to_ms(x), normalize_impulse(ir), "causal"))
    return convolve(x, normalize_impulse(ir), "causal")


# ------------------------------------------------------------------ render/graph.py:56-177 + render/core.py
def render_plan(processors, input_signals, per_type_parameters, plan):
    """Restatement of render_grafx for the slice/sum/index/scatter plans of render/prepare.py.
    `plan` is a list of dicts: {type, reads:[(method, idx)], aggs:[(method, idx)], param:(method, idx),
    write:(method, idx)}; processors maps type -> callable(*inputs, **params)."""
    four_d = input_signals.ndim == 4
    nd = 1 if four_d else 0
    num_nodes = plan["num_nodes"]
    if four_d:
        Bn, V0, C, L = input_signals.shape
        buf = input_signals.new_zeros(Bn, num_nodes, C, L)
        buf[:, :V0] = input_signals
    else:
        V0, C, L = input_signals.shape
        buf = input_signals.new_zeros(num_nodes, C, L)
        buf[:V0] = input_signals

    def read(t, acc):
        method, idx = acc
        if method == "slice":
            return t.narrow(nd, idx[0], idx[1] - idx[0])
        return t.index_select(nd, torch.as_tensor(idx))

    out = None
    for it in plan["iters"][1:]:
        ins = []
        for acc, (am, aidx) in zip(it["reads"], it["aggs"]):
            s = read(buf, acc)
            if am == "sum":
                s = s.sum(nd, keepdim=True)
            elif am == "scatter":
                aidx = torch.as_tensor(aidx)
                n_out = int(aidx.max()) + 1
                shape = list(s.shape)
                shape[nd] = n_out
                s = s.new_zeros(shape).index_add_(nd, aidx, s)
            if four_d:
                s = s.reshape(-1, *s.shape[2:])
            ins.append(s)
        t = it["type"]
        if t in processors:
            params = {}
            for k, v in per_type_parameters[t].items():
                v = v.unsqueeze(0).expand(Bn, *v.shape) if four_d else v
                v = read(v, it["param"])
                params[k] = v.reshape(-1, *v.shape[2:]) if four_d else v
            out = processors[t](*ins, **params)
            if isinstance(out, tuple):
                out = out[0]
        else:
            out = ins[0]
        if four_d:
            out = out.reshape(Bn, -1, C, L)
        method, idx = it["write"]
        if method == "slice":
            if four_d:
                buf[:, idx[0] : idx[1]] = out
            else:
                buf[idx[0] : idx[1]] = out
        else:
            ii = torch.as_tensor(idx)
            if four_d:
                buf[:, ii] = out
            else:
                buf[ii] = out
    return out, buf


# ====================================================================================================
# "next" rows of SURVEY.md section 8(f): equalizers on the conv / cascade engines, memoryless processors
# ====================================================================================================
def geq_coeffs(log_gains, fc, fB, sr, neighbour_exponent=0.4):
    """core/geq.py:176-209.  fc / fB: the band centre frequencies / bandwidths [K] (Hz) the reference registers."""
    g = torch.exp(log_gains)
    gt2 = torch.exp(neighbour_exponent * log_gains) ** 2
    tan_half = torch.tan(math.pi * fB.to(log_gains.dtype) / sr)
    mult = torch.sqrt((torch.abs(1 - gt2) + 1e-7) / (torch.abs(g ** 2 - gt2) + 1e-7))
    beta = torch.where(log_gains.abs() >= 1e-3, tan_half * mult, tan_half.expand_as(g))
    m2c = (-2 * torch.cos(2 * math.pi * fc.to(log_gains.dtype) / sr)).expand_as(g)
    Bs = torch.stack([1 + g * beta, m2c, 1 - g * beta], -1)
    As = torch.stack([1 + beta, m2c, 1 - beta], -1)
    return Bs, As


def graphic_equalizer(x, log_gains, fc, fB, sr=44100, processor_channel="mono", backend="lfilter", fsm_fir_len=4000,
                      use_torchaudio=True):
    """eq.py:407-420."""
    Bs, As = geq_coeffs(log_gains, fc, fB, sr)
    if processor_channel == "midside":
        return ms_to_lr(_iir(lr_to_ms(x), Bs, As, backend, fsm_fir_len, use_torchaudio))
    return _iir(x, Bs, As, backend, fsm_fir_len, use_torchaudio)


def zerophase_fir(log_magnitude, window=None, filterbank=None, eps=1e-7):
    """core/fir.py:25-40 / :106-123: h = window * roll(irfft(|H|, 2K - 1), K - 1); with a filterbank matrix
    [K_fb, K] (synthesis direction) |H| = sqrt(exp(H_fb)^2 @ M + eps)."""
    mag = torch.exp(log_magnitude)
    if filterbank is not None:
        mag = torch.sqrt((mag ** 2) @ filterbank.to(mag.dtype) + eps)
    n = 2 * mag.shape[-1] - 1
    ir = torch.roll(torch.fft.irfft(mag, n=n), shifts=n // 2, dims=-1)
    if window is not None:
        ir = ir * window.to(ir.dtype)
    return ir


def zerophase_fir_equalizer(x, log_magnitude, window=None, filterbank=None, eps=1e-7, processor_channel="mono"):
    """eq.py:70-72 (log_magnitude [B, K] -> one filter for every channel) and eq.py:176-214 ([B, C_eq, K])."""
    fir = zerophase_fir(log_magnitude, window, filterbank, eps)
    if fir.ndim == 2:
        fir = fir[:, None, :]
    if processor_channel == "midside":
        return ms_to_lr(convolve(lr_to_ms(x), fir, "zerophase"))
    return convolve(x, fir, "zerophase")


def stereo_gain(x, log_gain):
    """stereo.py:31-38."""
    return x * torch.exp(log_gain)[..., None]


def side_gain_imager(x, log_gain):
    """stereo.py:71-84."""
    left, right = x[:, 0, :], x[:, 1, :]
    mid, side = left + right, torch.exp(log_gain) * (left - right)
    return torch.stack([(mid + side) / 2, (mid - side) / 2], 1)


def _pre_post(x, remove_dc, log_pre_gain):
    if remove_dc:
        x = x - x.mean(-1, keepdim=True)
    pre = None
    if log_pre_gain is not None:
        pre = torch.exp(log_pre_gain).unsqueeze(-1)
        x = x * pre
    return x, pre


def tanh_distortion(x, log_pre_gain=None, log_post_gain=None, bias=None, inverse_post_gain=True, remove_dc=False):
    """nonlinear.py:64-89 (log_pre_gain None <=> pre_post_gain False; bias None <=> use_bias False)."""
    x, pre = _pre_post(x, remove_dc, log_pre_gain)
    y = torch.tanh(x) if bias is None else torch.tanh(x + bias.unsqueeze(-1)) - torch.tanh(bias.unsqueeze(-1))
    if pre is not None:
        y = y * (1 / pre if inverse_post_gain else torch.exp(log_post_gain).unsqueeze(-1))
    return y


def piecewise_tanh_distortion(x, log_hardness, z_threshold, log_pre_gain=None, log_post_gain=None,
                              inverse_post_gain=True, remove_dc=False):
    """nonlinear.py:159-205."""
    x, pre = _pre_post(x, remove_dc, log_pre_gain)
    hard, thr = torch.exp(log_hardness).unsqueeze(-2), torch.sigmoid(z_threshold).unsqueeze(-2)
    kn, kp = thr[..., 0:1], thr[..., 1:2]
    gp, gn = hard[..., 0:1], hard[..., 1:2]
    ap, an = (1 - torch.tanh(kp)) / gp, (1 - torch.tanh(kn)) / gn
    above, below = x > kp, x < -kn
    y = torch.where(above, ap * torch.tanh(gp * (x - kp)) + torch.tanh(kp),
                    torch.where(below, an * torch.tanh(gn * (x + kn)) - torch.tanh(kn), torch.tanh(x)))
    if pre is not None:
        y = y * (1 / pre if inverse_post_gain else torch.exp(log_post_gain).unsqueeze(-1))
    return y


def series_distortion(kind, x, basis_weights, log_pre_gain=None, remove_dc=False, use_tanh=False):
    """PowerDistortion (nonlinear.py:268-285: basis x^k) / ChebyshevDistortion (:349-384: basis T_k(x)), k < order."""
    x, _ = _pre_post(x, remove_dc, log_pre_gain)
    w = torch.tanh(basis_weights)
    order = w.shape[-1]
    basis = [torch.ones_like(x), x]
    for k in range(2, order):
        basis.append(basis[-1] * x if kind == "power" else 2 * x * basis[-1] - basis[-2])
    y = torch.zeros_like(x)
    for k in range(order):
        b = torch.tanh(basis[k]) if use_tanh else basis[k]
        y = y + w[:, k].view(-1, 1, 1) * b
    return y


def parallel_mix_weights(parallel_weights, activation="softmax"):
    """container.py:218-222."""
    n = parallel_weights.shape[-1]
    if activation == "softmax":
        return torch.softmax(parallel_weights, -1)
    return torch.nn.functional.softplus(parallel_weights) / (math.log(2) * n)


def surrogate_delay_ir(delay_z, segment_len, straight_through=True):
    """core/delay.py:40-77 forward values: z -> z tanh|z| / |z|; irfft of (z + 1e-7)^n, n <= N/2; with the
    straight-through trick the forward value is a unit impulse at the arg-max."""
    z = torch.view_as_complex(delay_z.contiguous()).reshape(-1)
    mag = z.abs()
    loss = ((1 - torch.tanh(mag)) ** 2).sum()
    z = z * torch.tanh(mag) / (mag + 1e-7)
    n = torch.arange(segment_len // 2 + 1, device=z.device)[None, :]
    irs = torch.fft.irfft((z[:, None] + 1e-7) ** n)
    if straight_through:
        hard = torch.zeros_like(irs)
        hard[torch.arange(irs.shape[0]), irs.argmax(-1)] = 1
        irs = irs + (hard - irs)
    return irs.reshape(*delay_z.shape[:-1], -1), loss


def multitap_delay(x, delay_z, log_fir_magnitude=None, window=None, segment_len=3000, num_segments=20,
                   num_delay_per_segment=1, num_channels=2, pre_delay=0):
    """delay.py:104-140."""
    irs, loss = surrogate_delay_ir(delay_z, segment_len)
    if log_fir_magnitude is not None:
        irs = convolve(irs, zerophase_fir(log_fir_magnitude, window), "zerophase")
    b, _, t = irs.shape
    irs = irs.reshape(b, num_channels, num_segments, num_delay_per_segment, t).sum(-2).reshape(b, num_channels, -1)
    y = convolve(x, normalize_impulse(irs), "causal")
    if pre_delay:
        y = F.pad(y, (pre_delay, 0))[:, :, :-pre_delay]
    return y, loss


def noise_shaping_reverb(x, log_decay, log_gain, filtered_noise, min_decay, max_decay, log_fade_in=None,
                         z_fade_in_gain=None, processor_channel="midside"):
    """reverb.py:364-398 ("fixed" noise: filtered_noise [C, K, ir_len])."""
    t = torch.arange(filtered_noise.shape[-1], device=x.device)[None, None, None, :]
    decay = torch.sigmoid(log_decay) * (max_decay - min_decay) + min_decay
    env = torch.exp(t * decay.unsqueeze(-1))
    if log_fade_in is not None:
        fade = torch.sigmoid(log_fade_in) * (decay - min_decay) + min_decay
        env = env - torch.exp(t * fade.unsqueeze(-1)) * torch.sigmoid(z_fade_in_gain).unsqueeze(-1)
    ir = (filtered_noise.to(env.dtype)[None] * env * log_gain.unsqueeze(-1)).sum(2)
    ir = normalize_impulse(ir)
    if processor_channel == "midside":
        return ms_to_lr(convolve(lr_to_ms(x), ir, "causal"))
    return convolve(x, ir, "causal")
