"""Tensor-level wrappers around the C ABI (include/grafx_b200.h).  Each function validates
shapes the way the reference does (bare asserts / ValueError), makes the operands contiguous,
allocates the output and the scratch workspace with torch (device memory plumbing only) and
enqueues the kernel on the current CUDA stream.  Forward only: inputs are detached.
"""
from __future__ import annotations

import torch

from . import _cabi


def _prep(t: torch.Tensor, dtype=None) -> torch.Tensor:
    t = t.detach()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def biquad_cascade(x: torch.Tensor, Bs: torch.Tensor, As: torch.Tensor) -> torch.Tensor:
    """Exact cascade of K biquads (reference: IIRFilter._process_lfilter, core/iir.py:154-184).

    x [B, C, L]; Bs, As [B, Cf, K, 3] -> y [B, max(C, Cf), L].  float32 or float64."""
    _cabi.require_cuda(x, Bs, As)
    assert x.ndim == 3 and Bs.ndim == 4 and As.shape == Bs.shape and Bs.shape[-1] == 3
    b, c_sig, L = x.shape
    bb, c_filt, K, _ = Bs.shape
    assert bb == b, "batch size of the coefficients must match the signal"
    if not (c_sig == c_filt or c_sig == 1 or c_filt == 1):
        raise AssertionError("channel mismatch between signal and filter")
    dtype = x.dtype if x.dtype in (torch.float32, torch.float64) else torch.float32
    x, Bs, As = _prep(x, dtype), _prep(Bs, dtype), _prep(As, dtype)
    c_out = max(c_sig, c_filt)
    y = torch.empty(b, c_out, L, dtype=dtype, device=x.device)
    if y.numel() == 0:
        return y
    L_ = _cabi.lib()
    elem = 4 if dtype == torch.float32 else 8
    ws_bytes = L_.gfx_biquad_cascade_workspace_bytes(b, c_sig, c_filt, K, elem)
    ws = _cabi.workspace(ws_bytes, x.device)
    fn = L_.gfx_biquad_cascade_f32 if dtype == torch.float32 else L_.gfx_biquad_cascade_f64
    with torch.cuda.device(x.device):
        code = fn(x.data_ptr(), y.data_ptr(), Bs.data_ptr(), As.data_ptr(), b, c_sig, c_filt, K, L,
                  ws.data_ptr(), ws.numel(), _cabi.stream_ptr())
    _cabi.check(code, "gfx_biquad_cascade")
    return y


def _midside(x: torch.Tensor, mult: float) -> torch.Tensor:
    _cabi.require_cuda(x)
    assert x.ndim == 3 and x.shape[1] == 2, "mid/side conversion needs [B, 2, L]"
    x = _prep(x, torch.float32)
    y = torch.empty_like(x)
    if y.numel() == 0:
        return y
    with torch.cuda.device(x.device):
        code = _cabi.lib().gfx_midside_f32(x.data_ptr(), y.data_ptr(), x.shape[0], x.shape[2], mult,
                                           _cabi.stream_ptr())
    _cabi.check(code, "gfx_midside_f32")
    return y


def lr_to_ms(x: torch.Tensor, mult: float | None = 0.5) -> torch.Tensor:
    """core/midside.py:11-17."""
    return _midside(x, 1.0 if mult is None else float(mult))


def ms_to_lr(x: torch.Tensor) -> torch.Tensor:
    """core/midside.py:4-8."""
    return _midside(x, 1.0)


def iir_fsm(x: torch.Tensor, Bs: torch.Tensor, As: torch.Tensor, fir_len: int) -> torch.Tensor:
    """Frequency-sampled FIR of the cascade + causal convolution (core/iir.py:147-152,263-276)."""
    raise NotImplementedError("fsm backend: FIR convolution kernel not built yet")
