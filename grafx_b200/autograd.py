"""Backward passes (first step): the biquad cascade and the causal FIR convolution.

Upstream differentiates `IIRFilter._process_lfilter` (processors/core/iir.py:154-196) through torchaudio's `lfilter`
autograd, one section at a time.  Here the same gradients come from the forward kernels themselves:

  forward   s_0 = x,  s_k = (B_k / A_k) s_{k-1},  y = s_K          (sections kept: K + 1 signals)
  backward  for k = K .. 1, with g_K = dL/dy:
      u_k     = (1 / A_k)^T g_k          the all-pole recursion of section k on the time-REVERSED gradient
      dL/db_kj =  sum_n u_k[n] s_{k-1}[n - j]          j = 0, 1, 2     (gfx_lag_dots_f32)
      dL/da_kj = -sum_n u_k[n] s_k[n - j]              j = 1, 2
      g_{k-1} = B_k^T u_k                the section's feed-forward taps, again on reversed time
  dL/dx = g_0.

Both recursions are launches of `gfx_biquad_cascade_f32` (csrc/biquad.cu) -- including its double-precision carries
for sections with poles near z = 1 -- so a training step costs about 3 K + K single-section passes over the audio.
The gradient signal stays time-reversed between sections (one flip on the way in, one on the way out).

Coefficients enter NORMALISED (a0 = 1); `functional.biquad_cascade` divides by a0 in PyTorch first, so gradients with
respect to un-normalised (Bs, As) -- and through the coefficient design formulas of processors/design.py -- are
PyTorch autograd on [B, C, K, 3] tensors.
"""
from __future__ import annotations

import torch

from . import _cabi


def _cascade_raw(x: torch.Tensor, Bs: torch.Tensor, As: torch.Tensor) -> torch.Tensor:
    """One launch of the forward kernel on matched channels, no autograd bookkeeping."""
    b, c, L = x.shape
    K = Bs.shape[2]
    y = torch.empty_like(x)
    L_ = _cabi.lib()
    ws = _cabi.workspace(L_.gfx_biquad_cascade_workspace_bytes(b, c, c, K, 4), x.device)
    with torch.cuda.device(x.device):
        code = L_.gfx_biquad_cascade_f32(x.data_ptr(), y.data_ptr(), Bs.data_ptr(), As.data_ptr(), b, c, c, K, L,
                                         ws.data_ptr(), ws.numel(), _cabi.stream_ptr())
    _cabi.check(code, "gfx_biquad_cascade_f32")
    return y


def _lag_dots(u: torch.Tensor, s0: torch.Tensor, s1: torch.Tensor | None, u_reversed: bool):
    """(sum_n u[n] s0[n-j], sum_n u[n] s1[n-j]) for j = 0..2 -> two [B, C, 3] tensors (the second None without s1)."""
    b, c, L = u.shape
    out0 = torch.empty(b, c, 3, dtype=torch.float32, device=u.device)
    out1 = torch.empty_like(out0) if s1 is not None else None
    with torch.cuda.device(u.device):
        code = _cabi.lib().gfx_lag_dots_f32(u.data_ptr(), s0.data_ptr(), _cabi.ptr(s1), out0.data_ptr(), _cabi.ptr(out1),
                                            b * c, L, int(u_reversed), _cabi.stream_ptr())
    _cabi.check(code, "gfx_lag_dots_f32")
    return out0, out1


class BiquadCascadeFn(torch.autograd.Function):
    """y = cascade(x; nb, na) for x [B, C, L], nb / na [B, C, K, 3] (float32, contiguous, na[..., 0] == 1)."""

    @staticmethod
    def forward(ctx, x, nb, na):
        K = nb.shape[2]
        if not (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            # only dL/dx is wanted: no section signal is needed, forward and backward are one fused launch each
            ctx.save_for_backward(nb, na)
            return _cascade_raw(x, nb, na)
        sections = [x]
        for k in range(K):
            sections.append(_cascade_raw(sections[-1], nb[:, :, k:k + 1].contiguous(), na[:, :, k:k + 1].contiguous()))
        ctx.save_for_backward(nb, na, *sections)
        return sections[-1]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_y):
        nb, na, *sections = ctx.saved_tensors
        K = nb.shape[2]
        need_x, need_b, need_a = ctx.needs_input_grad
        if not sections:
            # the sections are LTI and commute: the adjoint of the whole cascade is the cascade on reversed time
            return _cascade_raw(grad_y.detach().to(torch.float32).flip(-1).contiguous(), nb, na).flip(-1), None, None
        one = torch.zeros_like(nb[:, :, :1])
        one[..., 0] = 1.0
        r = grad_y.detach().to(torch.float32).flip(-1).contiguous()  # the gradient, time-reversed
        gb = torch.zeros_like(nb) if need_b else None
        ga = torch.zeros_like(na) if need_a else None
        for k in range(K - 1, -1, -1):
            bk, ak = nb[:, :, k:k + 1].contiguous(), na[:, :, k:k + 1].contiguous()
            if need_b or need_a:
                u_rev = _cascade_raw(r, one, ak)  # (1 / A_k)^T g_k, reversed time
                d_in, d_out = _lag_dots(u_rev, sections[k], sections[k + 1] if need_a else None, True)
                if need_b:
                    gb[:, :, k] = d_in
                if need_a:
                    ga[:, :, k, 1:] = -d_out[..., 1:]
            if k > 0 or need_x:
                r = _cascade_raw(r, bk, ak)  # B_k^T (1 / A_k)^T g_k, reversed time
        gx = r.flip(-1) if need_x else None
        return gx, gb, ga


def biquad_cascade_autograd(x: torch.Tensor, Bs: torch.Tensor, As: torch.Tensor) -> torch.Tensor:
    """Differentiable `functional.biquad_cascade` (float32): broadcast of the channel axis and the division by a0
    are PyTorch ops, so their gradients are autograd's; the O(samples) work is the Function above."""
    _cabi.require_cuda(x, Bs, As)
    b, c_sig, L = x.shape
    c_filt = Bs.shape[1]
    c_out = max(c_sig, c_filt)
    x = x.to(torch.float32).expand(b, c_out, L).contiguous()
    Bs, As = Bs.to(torch.float32), As.to(torch.float32)
    a0 = As[..., :1]
    nb = (Bs / a0).expand(b, c_out, *Bs.shape[2:]).contiguous()
    na = (As / a0).expand(b, c_out, *As.shape[2:]).contiguous()
    return BiquadCascadeFn.apply(x, nb, na)


# ---------------------------------------------------------------------------------------------- FIR convolution
def _fir_raw(x: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    from . import functional as F_

    with torch.no_grad():
        return F_.fir_conv(x, h, "causal")


class FirConvCausalFn(torch.autograd.Function):
    """y = convolve(x, h, "causal") for x [B, C, L], h [B, C, N] (reference: core/convolution.py:119-134, differentiated
    upstream through torch.fft).  Both gradients are causal convolutions on the FIR engine (csrc/fir.cu):
      dL/dx[i] = sum_m g[i + m] h[m]                 = flip(conv(flip(g), h))
      dL/dh[k] = sum_n g[n] x[n - k], k < N          = conv(flip(g), x)[L - 1 - k]   (x as an L-tap filter)."""

    @staticmethod
    def forward(ctx, x, h):
        ctx.save_for_backward(x, h)
        return _fir_raw(x, h)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_y):
        x, h = ctx.saved_tensors
        L, N = x.shape[-1], h.shape[-1]
        need_x, need_h = ctx.needs_input_grad
        gr = grad_y.detach().to(torch.float32).flip(-1).contiguous()
        gx = _fir_raw(gr, h).flip(-1) if need_x else None
        gh = None
        if need_h:
            lags = _fir_raw(gr, x).flip(-1)  # lags[k] = sum_n g[n] x[n - k], k = 0 .. L-1
            gh = lags[..., :N] if N <= L else torch.nn.functional.pad(lags, (0, N - L))
        return gx, gh


def fir_conv_autograd(x: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """Differentiable causal `functional.fir_conv`: the channel broadcast is a PyTorch expand (autograd sums it)."""
    _cabi.require_cuda(x, h)
    B, cx, L = x.shape
    ch, N = h.shape[1], h.shape[2]
    c = max(cx, ch)
    return FirConvCausalFn.apply(x.to(torch.float32).expand(B, c, L).contiguous(), h.to(torch.float32).expand(B, c, N).contiguous())


# ---------------------------------------------------------------------------------------------- gains and mixes
def _pointwise_raw(op, x, p0):
    from . import functional as F_

    with torch.no_grad():
        return F_.pointwise(op, x, p0)


class GainFn(torch.autograd.Function):
    """y[b, c] = x[b, c] * exp(log_gain[b, c])  (StereoGain.forward, processors/stereo.py:31-38).
    dL/dx is the same kernel on the gradient; dL/dlog_gain[b, c] = sum_t g y (a lag-0 inner product per row)."""

    @staticmethod
    def forward(ctx, x, log_gain):
        y = _pointwise_raw("gain", x, log_gain)
        ctx.save_for_backward(log_gain, y)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_y):
        log_gain, y = ctx.saved_tensors
        g = grad_y.detach().to(torch.float32).contiguous()
        gx = _pointwise_raw("gain", g, log_gain) if ctx.needs_input_grad[0] else None
        glg = _lag_dots(g, y, None, False)[0][..., 0].reshape(log_gain.shape) if ctx.needs_input_grad[1] else None
        return gx, glg


class DryWetFn(torch.autograd.Function):
    """y = w wet + (1 - w) dry with one weight per batch item (DryWet.forward, processors/container.py:62-67)."""

    @staticmethod
    def forward(ctx, dry, wet, w):
        from . import functional as F_

        with torch.no_grad():
            y = F_.drywet_mix(dry, wet, w)
        ctx.save_for_backward(dry, wet, w)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_y):
        dry, wet, w = ctx.saved_tensors
        g = grad_y.detach().to(torch.float32).contiguous()
        B = g.shape[0]
        wv = w.detach().to(torch.float32).reshape(B).contiguous()
        need_dry, need_wet, need_w = ctx.needs_input_grad
        g_dry = _pointwise_raw("scale_add", g, 1.0 - wv) if need_dry else None
        g_wet = _pointwise_raw("scale_add", g, wv) if need_wet else None
        g_w = None
        if need_w:
            flat = lambda t: t.reshape(B, 1, -1)  # noqa: E731
            d_wet, d_dry = _lag_dots(flat(g), flat(wet), flat(dry), False)
            g_w = (d_wet[..., 0] - d_dry[..., 0]).reshape(w.shape)
        return g_dry, g_wet, g_w
