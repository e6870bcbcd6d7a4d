"""Tensor-level wrappers around the C ABI (include/grafx_b200.h).  Each function validates
shapes the way the reference does (bare asserts / ValueError), makes the operands contiguous,
allocates the output and the scratch workspace with torch (device memory plumbing only) and
enqueues the kernel on the current CUDA stream.  Without autograd (the hot path) inputs are detached and the fused
kernels run; when autograd expects a gradient the differentiable variants take over (autograd.py, training.py) or, where
none exists, the call raises.
"""
from __future__ import annotations

import threading

import torch

from . import _cabi


_DEST = threading.local()


class output_into:
    """`with output_into(view):` -- the next kernel output of exactly this shape / dtype / device is
    written straight into `view` (a contiguous slice of a caller-owned buffer, e.g. the signal buffer of
    render_grafx) instead of a fresh tensor.  The caller compares data_ptr()s afterwards: a processor
    whose last op is not one of ours simply returns its own tensor, to be copied as usual."""

    def __init__(self, dest: torch.Tensor | None):
        self.dest = dest if (dest is not None and dest.is_contiguous()) else None

    def __enter__(self):
        self.prev = getattr(_DEST, "t", None)
        _DEST.t = self.dest
        return self

    def __exit__(self, *exc):
        _DEST.t = self.prev
        return False


class shared_parameters:
    """`with shared_parameters(B):` -- tells the processors that the parameter rows they receive repeat in
    runs of B (rows n*B .. n*B+B-1 are identical): what the 4-D source path of render_grafx produces when it
    expands the per-node parameters over the batch of renders.  Parameter-side work that is O(samples)
    (the reverb impulse response and its spectra) is then done once per run."""

    def __init__(self, repeat: int | None):
        self.repeat = int(repeat) if repeat else 1

    def __enter__(self):
        self.prev = getattr(_DEST, "repeat", 1)
        _DEST.repeat = self.repeat
        return self

    def __exit__(self, *exc):
        _DEST.repeat = self.prev
        return False


def parameter_repeat() -> int:
    return getattr(_DEST, "repeat", 1)


class source_fold:
    """`with source_fold(sources, view) as f:` -- render_grafx's first render order (4-D sources).  `view` is the
    source slice of the signal buffer, flattened node-major `[V0 * B, C, L]` and NOT filled yet; `sources` is the
    caller's `[B, V0, C, L]` tensor.  An op that supports it (the biquad cascade, gfx_biquad_cascade_ex_f32) and is
    handed `view` as its signal reads `sources` instead and fills `view` from the tile it staged anyway, so the separate
    copy pass never reads the sources a second time.  `f.used` tells the caller whether that happened: if not, the
    caller fills `view` itself and runs the processor again (render/graph.py)."""

    def __init__(self, sources: torch.Tensor, view: torch.Tensor):
        self.sources, self.view, self.used = sources, view, False

    def __enter__(self):
        self.prev = getattr(_DEST, "fold", None)
        _DEST.fold = self
        return self

    def __exit__(self, *exc):
        _DEST.fold = self.prev
        return False


def _take_source_fold(x: torch.Tensor):
    f = getattr(_DEST, "fold", None)
    if (f is None or f.used or x.dtype != torch.float32 or x.data_ptr() != f.view.data_ptr()
            or tuple(x.shape) != tuple(f.view.shape) or not x.is_contiguous()):
        return None
    return f


def _new_output(shape, dtype, device) -> torch.Tensor:
    d = getattr(_DEST, "t", None)
    if d is not None and tuple(d.shape) == tuple(shape) and d.dtype == dtype and d.device == device:
        _DEST.t = None
        return d
    return torch.empty(shape, dtype=dtype, device=device)


def _wants_grad(*tensors) -> bool:
    """True when autograd is recording and one of the tensors is part of a graph: the differentiable variant of the
    op runs (grafx_b200/autograd.py for the biquad cascade, the PyTorch statements for O(parameters) design formulas);
    ops without one raise instead of silently cutting the graph."""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _no_backward(name: str, *tensors):
    if _wants_grad(*tensors):
        raise NotImplementedError(
            f"{name} has no backward pass yet (forward-only CUDA kernel); wrap the call in torch.no_grad() or detach "
            "its inputs.  Differentiable so far: the IIR family, the causal FIR convolution, StereoGain, the DryWet mix.")


def _prep(t: torch.Tensor, dtype=None) -> torch.Tensor:
    t = t.detach()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def biquad_cascade(x: torch.Tensor, Bs: torch.Tensor, As: torch.Tensor) -> torch.Tensor:
    """Exact cascade of K biquads (reference: IIRFilter._process_lfilter, core/iir.py:154-184).

    x [B, C, L]; Bs, As [B, Cf, K, 3] -> y [B, max(C, Cf), L].  float32 or float64.
    Inside `shared_parameters(R)` (render_grafx's 4-D sources) the coefficients may also come un-expanded,
    [B / R, Cf, K, 3]: runs of R consecutive batch items then share a coefficient row."""
    _cabi.require_cuda(x, Bs, As)
    assert x.ndim == 3 and Bs.ndim == 4 and As.shape == Bs.shape and Bs.shape[-1] == 3
    b, c_sig, L = x.shape
    bb, c_filt, K, _ = Bs.shape
    rep = 1
    if bb != b:
        rep = parameter_repeat()
        assert rep > 1 and bb * rep == b, "batch size of the coefficients must match the signal"
    if not (c_sig == c_filt or c_sig == 1 or c_filt == 1):
        raise AssertionError("channel mismatch between signal and filter")
    if rep > 1 and (_wants_grad(x, Bs, As) or x.dtype == torch.float64):
        Bs, As, rep = Bs.repeat_interleave(rep, 0), As.repeat_interleave(rep, 0), 1  # (differentiable; float64: no repeat form)
    if _wants_grad(x, Bs, As):
        if x.dtype == torch.float64:
            raise NotImplementedError("the backward pass of the biquad cascade is float32 only")
        from .autograd import biquad_cascade_autograd

        return biquad_cascade_autograd(x, Bs, As)
    dtype = x.dtype if x.dtype in (torch.float32, torch.float64) else torch.float32
    fold = _take_source_fold(x) if c_sig >= c_filt else None
    x, Bs, As = _prep(x, dtype), _prep(Bs, dtype), _prep(As, dtype)
    c_out = max(c_sig, c_filt)
    y = _new_output((b, c_out, L), dtype, x.device)
    if y.numel() == 0:
        return y
    L_ = _cabi.lib()
    elem = 4 if dtype == torch.float32 else 8
    ws_bytes = L_.gfx_biquad_cascade_workspace_bytes(b, c_sig, c_filt, K, elem)
    ws = _cabi.workspace(ws_bytes, x.device)
    if fold is not None or rep > 1:
        # render_grafx's batched sources: un-expanded coefficient rows and / or the first render order, which reads the
        # caller's [B, V0, C, L] sources and fills the buffer's source slice (x) on the way
        src, xcopy, n_renders, n_sources = x.data_ptr(), None, 0, 0
        if fold is not None:
            src, xcopy = fold.sources.data_ptr(), x.data_ptr()
            n_renders, n_sources = fold.sources.shape[0], fold.sources.shape[1]
        with torch.cuda.device(x.device):
            code = L_.gfx_biquad_cascade_ex_f32(src, xcopy, y.data_ptr(), Bs.data_ptr(), As.data_ptr(), b, c_sig, c_filt, K, L,
                                                n_renders, n_sources, rep, ws.data_ptr(), ws.numel(), _cabi.stream_ptr())
        _cabi.check(code, "gfx_biquad_cascade_ex")
        if fold is not None:
            fold.used = True
        return y
    fn = L_.gfx_biquad_cascade_f32 if dtype == torch.float32 else L_.gfx_biquad_cascade_f64
    with torch.cuda.device(x.device):
        code = fn(x.data_ptr(), y.data_ptr(), Bs.data_ptr(), As.data_ptr(), b, c_sig, c_filt, K, L,
                  ws.data_ptr(), ws.numel(), _cabi.stream_ptr())
    _cabi.check(code, "gfx_biquad_cascade")
    return y


def _midside(x: torch.Tensor, mult: float) -> torch.Tensor:
    _cabi.require_cuda(x)
    assert x.ndim == 3 and x.shape[1] == 2, "mid/side conversion needs [B, 2, L]"
    if _wants_grad(x):  # linear and its own adjoint up to the factor: the PyTorch statement carries the graph
        return torch.stack([(x[:, 0] + x[:, 1]) * mult, (x[:, 0] - x[:, 1]) * mult], 1)
    x = _prep(x, torch.float32)
    y = _new_output(tuple(x.shape), torch.float32, x.device)
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


_PLANS: dict = {}


def _fft_plan(device: torch.device, n: int) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), n)
    plan = _PLANS.get(key)
    if plan is None:
        L_ = _cabi.lib()
        plan = torch.empty(L_.gfx_fft_plan_bytes(n), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            _cabi.check(L_.gfx_fft_plan_init(plan.data_ptr(), n, _cabi.stream_ptr()), "gfx_fft_plan_init")
        _PLANS[key] = plan
    return plan


def fir_conv(x: torch.Tensor, h: torch.Tensor, mode: str = "causal", h_repeat: int = 1) -> torch.Tensor:
    """Linear convolution sliced to len(x) (reference: convolve(), core/convolution.py:119-134).

    x [B, Cx, L] (or [B, L]), h [B, Ch, N] (or [B, N]); channels broadcast; mode "causal" or
    "zerophase" (output shifted by N // 2)."""
    _cabi.require_cuda(x, h)
    if mode not in ("causal", "zerophase"):
        raise ValueError(f"unsupported convolution mode: {mode}")
    squeeze = x.ndim == 2 and h.ndim == 2
    x3 = x.unsqueeze(1) if x.ndim == 2 else x
    h3 = h.unsqueeze(1) if h.ndim == 2 else h
    assert x3.ndim == 3 and h3.ndim == 3 and x3.shape[0] == h3.shape[0] * h_repeat
    B, cx, L = x3.shape
    _, ch, N = h3.shape
    assert cx == ch or cx == 1 or ch == 1, "channel mismatch between signal and filter"
    if x.dtype == torch.float64 or h.dtype == torch.float64:
        raise TypeError("the FFT convolution engine is float32 only (upstream would compute float64 inputs in float64): "
                        "cast the operands explicitly; the biquad cascade has a float64 kernel")
    if _wants_grad(x, h):
        from .autograd import fir_conv_autograd

        if h_repeat != 1:  # (the expand is differentiable: autograd sums the gradient over the run)
            h3 = h3.repeat_interleave(h_repeat, 0)
        if mode == "zerophase":
            # y[n] = sum_k h[k] x[n + N//2 - k]: the causal convolution of the right-padded signal, read N//2 later
            y = fir_conv_autograd(torch.nn.functional.pad(x3, (0, N // 2)), h3)[..., N // 2: N // 2 + L]
        else:
            y = fir_conv_autograd(x3, h3)
        return y.squeeze(1) if squeeze else y
    x3, h3 = _prep(x3, torch.float32), _prep(h3, torch.float32)
    y = torch.empty(B, max(cx, ch), L, dtype=torch.float32, device=x.device) if squeeze else _new_output((B, max(cx, ch), L), torch.float32, x.device)
    if y.numel():
        L_ = _cabi.lib()
        n = L_.gfx_fir_fft_size(N)
        plan = _fft_plan(x.device, n)
        zp = int(mode == "zerophase")
        ws = _cabi.workspace(L_.gfx_fir_conv_workspace_bytes(B, cx, ch, L, N, zp), x.device)
        with torch.cuda.device(x.device):
            code = L_.gfx_fir_conv_f32(x3.data_ptr(), h3.data_ptr(), y.data_ptr(), B, cx, ch, L, N, zp, int(h_repeat),
                                       plan.data_ptr(), ws.data_ptr(), ws.numel(), _cabi.stream_ptr())
        _cabi.check(code, "gfx_fir_conv_f32")
    return y.squeeze(1) if squeeze else y


def fir_filter(x: torch.Tensor, fir: torch.Tensor) -> torch.Tensor:
    """FIRFilter semantics in one call (filter.py:65-77): causal convolution of x [B, Cx, L] with
    normalize_impulse(tanh(fir)), fir [B, Ch, N]; the activation and the unit-energy scale are folded into the filter
    spectra (csrc/fir.cu: fir_tanh_energy_kernel + TanhSrc)."""
    _cabi.require_cuda(x, fir)
    assert x.ndim == 3 and fir.ndim == 3 and x.shape[0] == fir.shape[0]
    B, cx, L = x.shape
    _, ch, N = fir.shape
    assert cx == ch or cx == 1 or ch == 1, "channel mismatch between signal and filter"
    if _wants_grad(x, fir):
        return fir_conv(x, normalize_impulse(torch.tanh(fir)), "causal")  # the differentiable statement
    x, fir = _prep(x, torch.float32), _prep(fir, torch.float32)
    y = _new_output((B, max(cx, ch), L), torch.float32, x.device)
    if y.numel():
        L_ = _cabi.lib()
        plan = _fft_plan(x.device, L_.gfx_fir_fft_size(N))
        ws = _cabi.workspace(L_.gfx_fir_conv_workspace_bytes(B, cx, ch, L, N, 0) + (4 * B + 255) // 256 * 256, x.device)
        with torch.cuda.device(x.device):
            code = L_.gfx_fir_filter_f32(x.data_ptr(), fir.data_ptr(), y.data_ptr(), B, cx, ch, L, N, plan.data_ptr(),
                                         ws.data_ptr(), ws.numel(), _cabi.stream_ptr())
        _cabi.check(code, "gfx_fir_filter_f32")
    return y


def normalize_impulse(ir: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """core/utils.py:14-18 -- O(filter taps) parameter-side math, stays in PyTorch."""
    assert ir.ndim == 3
    e = ir.square().sum(2, keepdim=True).mean(1, keepdim=True)
    return ir / torch.sqrt(e + eps)


def iir_fsm_fir(Bs: torch.Tensor, As: torch.Tensor, fir_len: int) -> torch.Tensor:
    """Frequency-sampled FIR of a biquad cascade (core/iir.py:147-150, 263-276):
    irfft_N( prod_k B_k(w_m) / A_k(w_m) ), w_m = 2 pi m / N.  O(parameters x N) filter DESIGN
    (512 rows x 4000 taps at config 2), evaluated with torch on the device."""
    m = torch.arange(fir_len // 2 + 1, device=Bs.device, dtype=torch.float32)
    ang = (-2.0 * torch.pi / fir_len) * m
    z1 = torch.polar(torch.ones_like(ang), ang)
    z2 = torch.polar(torch.ones_like(ang), 2 * ang)
    Bs, As = Bs.to(torch.float32), As.to(torch.float32)
    num = Bs[..., 0:1] + Bs[..., 1:2] * z1 + Bs[..., 2:3] * z2
    den = As[..., 0:1] + As[..., 1:2] * z1 + As[..., 2:3] * z2
    return torch.fft.irfft((num / den).prod(-2), n=fir_len, dim=-1)


def iir_fsm(x: torch.Tensor, Bs: torch.Tensor, As: torch.Tensor, fir_len: int) -> torch.Tensor:
    """Frequency-sampled FIR of the cascade + causal convolution (core/iir.py:147-152,263-276)."""
    _cabi.require_cuda(x, Bs, As)
    if _wants_grad(x, Bs, As):
        return fir_conv(x, iir_fsm_fir(Bs, As, fir_len), "causal")  # the FIR design is PyTorch: autograd's
    return fir_conv(x, iir_fsm_fir(Bs.detach(), As.detach(), fir_len), "causal")


_KNEE = {"hard": 0, "quadratic": 1, "exponential": 2, "approx_gate": 3}
_SMOOTHER = {None: 0, "iir": 1, "ballistics": 2}
_DYN_KIND = {"compressor": 0, "noisegate": 1}


def dynamics_chain(x: torch.Tensor, stages: list[dict], iir_len: int = 16384) -> torch.Tensor:
    """Compressor / NoiseGate, or several of them back to back, in one pass over the audio.

    Reference: Compressor.forward / NoiseGate.forward (processors/dynamics.py:361-419, 598-651).
    x [B, C, L].  Each stage is a dict: kind ("compressor"|"noisegate"), knee, energy_smoother,
    gain_smoother, gain_smooth_in_log, log_threshold, log_ratio, log_knee, z_alpha_pre,
    z_alpha_post (tensors with leading dim B)."""
    _cabi.require_cuda(x)
    if _wants_grad(x, *[v for st in stages for v in st.values() if isinstance(v, torch.Tensor)]):
        from . import training

        return training.dynamics_chain(x, stages, iir_len)  # (ballistics smoothers raise there)
    assert x.ndim == 3
    B, C, L = x.shape
    x = _prep(x, torch.float32)
    y = _new_output((B, C, L), torch.float32, x.device)
    if y.numel() == 0:
        return y
    L_ = _cabi.lib()
    keep = []  # keep parameter tensors alive until the launch is enqueued

    # parameter rows: B, or -- inside `shared_parameters(R)`, render_grafx's 4-D sources -- B / R un-expanded rows
    # (runs of R consecutive batch items share a row; every parameter tensor of the call the same way)
    rows = next(v.shape[0] for st in stages for v in st.values() if isinstance(v, torch.Tensor))
    rep = 1
    if rows != B:
        rep = parameter_repeat()
        assert rep > 1 and rows * rep == B, "batch size of the parameters must match the signal"

    def dev(t, cols):
        if t is None:
            return None
        _cabi.require_cuda(t)
        assert t.shape[0] == rows, "every parameter tensor of a dynamics call has the same number of rows"
        t = _prep(t, torch.float32).reshape(rows, -1)
        assert t.shape[1] == cols, f"parameter has {t.shape[1]} columns, expected {cols}"
        keep.append(t)
        return t.data_ptr()

    arr = (_cabi.DynamicsStage * len(stages))()
    for d, st in enumerate(stages):
        s = arr[d]
        s.kind = _DYN_KIND[st["kind"]]
        s.knee = _KNEE[st.get("knee", "quadratic")]
        s.energy_smoother = _SMOOTHER[st.get("energy_smoother", "iir")]
        s.gain_smoother = _SMOOTHER[st.get("gain_smoother", None)]
        s.gain_smooth_in_log = int(bool(st.get("gain_smooth_in_log", False)))
        s.log_threshold = dev(st["log_threshold"], 1)
        s.log_ratio = dev(st["log_ratio"], 1)
        s.log_knee = dev(st.get("log_knee"), 1) if s.knee != 0 else None
        if s.knee != 0 and s.log_knee is None:
            raise AssertionError("log_knee is required for the quadratic / exponential knee")
        s.z_alpha_pre = dev(st.get("z_alpha_pre"), s.energy_smoother) if s.energy_smoother else None
        s.z_alpha_post = dev(st.get("z_alpha_post"), s.gain_smoother) if s.gain_smoother else None
        if s.energy_smoother and s.z_alpha_pre is None:
            raise AssertionError("z_alpha_pre is required by the energy smoother")
        if s.gain_smoother and s.z_alpha_post is None:
            raise AssertionError("z_alpha_post is required by the gain smoother")
        if iir_len < L:
            # scratch for the truncation tail; only touched for rows whose pole is within ~100/N of 1
            if s.energy_smoother == 1 and d > 0:
                h = torch.empty(B, L, dtype=torch.float32, device=x.device)
                keep.append(h)
                s.hist_pre = h.data_ptr()
            if s.gain_smoother == 1:
                h = torch.empty(B, L, dtype=torch.float32, device=x.device)
                keep.append(h)
                s.hist_post = h.data_ptr()
    ws = _cabi.workspace(L_.gfx_dynamics_workspace_bytes(B, len(stages)), x.device)
    with torch.cuda.device(x.device):
        code = L_.gfx_dynamics_rep_f32(x.data_ptr(), y.data_ptr(), B, C, L, arr, len(stages), int(iir_len), rep,
                                       ws.data_ptr(), ws.numel(), _cabi.stream_ptr())
    _cabi.check(code, "gfx_dynamics_f32")
    # the caching allocator keeps stream order: freeing `keep`/`ws` here is safe on this stream
    return y


_DETECT = {"energy": 0, "amplitude": 1, None: 2}


def envelope(x: torch.Tensor, z: torch.Tensor, smoother: str, detect: str | None = None, log_out: bool = False,
             iir_len: int = 16384) -> torch.Tensor:
    """Stand-alone envelope smoothers / followers (core/envelope.py:34-60, 84-101; dynamics.py:745-767).

    x [B, L] with detect=None (the smoother applied to x itself) or [B, C, L] with detect "energy" |
    "amplitude"; z [B, 1] (smoother "iir") or [B, 2] ("ballistics") -> [B, L]."""
    _cabi.require_cuda(x, z)
    _no_backward("envelope", x, z)
    if detect is None:
        assert x.ndim == 2, "the smoothers take [B, L] signals"
        x3 = x.unsqueeze(1)
    else:
        assert x.ndim == 3
        x3 = x
    B, C, L = x3.shape
    sm = _SMOOTHER[smoother]
    assert sm in (1, 2)
    x3 = _prep(x3, torch.float32)
    y = torch.empty(B, L, dtype=torch.float32, device=x.device)
    if y.numel():
        z = _prep(z, torch.float32).reshape(B, -1)
        assert z.shape[1] == sm, f"z_alpha has {z.shape[1]} columns, expected {sm}"
        L_ = _cabi.lib()
        ws = _cabi.workspace(L_.gfx_dynamics_workspace_bytes(B, 1), x.device)
        with torch.cuda.device(x.device):
            code = L_.gfx_envelope_f32(x3.data_ptr(), y.data_ptr(), B, C, L, sm, z.data_ptr(), _DETECT[detect],
                                       int(bool(log_out)), int(iir_len), ws.data_ptr(), ws.numel(), _cabi.stream_ptr())
        _cabi.check(code, "gfx_envelope_f32")
    return y


def drywet_mix(dry: torch.Tensor, wet: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """y = w * wet + (1 - w) * dry, w per batch item used as given (container.py:62-67)."""
    _cabi.require_cuda(dry, wet, weight)
    if dry.shape != wet.shape and dry.ndim == 3 and wet.ndim == 3 and dry.shape[0] == wet.shape[0] and dry.shape[2] == wet.shape[2]:
        # upstream mixes by tensor broadcasting (container.py:62-65): a mono side follows the stereo one
        if dry.shape[1] == 1:
            dry = dry.expand_as(wet)
        elif wet.shape[1] == 1:
            wet = wet.expand_as(dry)
    if _wants_grad(dry, wet, weight):
        from .autograd import DryWetFn

        return DryWetFn.apply(dry.to(torch.float32).contiguous(), wet.to(torch.float32).contiguous(), weight)
    assert dry.shape == wet.shape
    dry, wet = _prep(dry, torch.float32), _prep(wet, torch.float32)
    w = _prep(weight, torch.float32).reshape(-1)
    assert w.numel() == dry.shape[0]
    y = torch.empty_like(dry)
    if y.numel() == 0:
        return y
    with torch.cuda.device(dry.device):
        code = _cabi.lib().gfx_drywet_f32(dry.data_ptr(), wet.data_ptr(), w.data_ptr(), y.data_ptr(), dry.shape[0],
                                          dry[0].numel(), _cabi.stream_ptr())
    _cabi.check(code, "gfx_drywet_f32")
    return y


def node_sum(src: torch.Tensor, node_dim: int, index: torch.Tensor | None = None, n_dst: int = 1,
             out: torch.Tensor | None = None) -> torch.Tensor:
    """Sum (index None) or scatter-sum over the node axis of a signal-buffer view
    (render/core.py:101-112).  `src` is [Q, C, L] (node_dim 0) or [B, Q, C, L] (node_dim 1) and may
    be a slice of the buffer (only the node/batch axes may be strided).  `out`, if given, is a
    view with n_dst nodes that is written in place."""
    _cabi.require_cuda(src)
    assert src.dtype == torch.float32
    if node_dim == 0:
        src4 = src.unsqueeze(0)
    else:
        src4 = src
    B, Q, C, L = src4.shape
    if src4.stride(3) != 1 or src4.stride(2) != L:
        src4 = src4.contiguous()
    shape = (B, n_dst, C, L)
    if out is None:
        out4 = torch.empty(shape, dtype=torch.float32, device=src.device)
    else:
        out4 = out.unsqueeze(0) if node_dim == 0 else out
        assert tuple(out4.shape) == shape and out4.stride(3) == 1 and out4.stride(2) == L
    idx_ptr = None
    if index is not None:
        index = index.to(device=src.device, dtype=torch.int32).contiguous()
        assert index.numel() == Q
        idx_ptr = index.data_ptr()
    if out4.numel():
        with torch.cuda.device(src.device):
            code = _cabi.lib().gfx_node_sum_f32(src4.data_ptr(), out4.data_ptr(), idx_ptr, B, Q, n_dst, C * L,
                                                src4.stride(0), src4.stride(1), out4.stride(0), out4.stride(1),
                                                _cabi.stream_ptr())
        _cabi.check(code, "gfx_node_sum_f32")
    if out is not None:
        return out
    return out4.squeeze(0) if node_dim == 0 else out4


def node_copy(src: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[i, j] = src[i, j] for [N0, N1, C, L] views whose two leading axes may be strided (a transposed
    batch/node pair): the source write of create_signal_buffer (render/core.py:6-33) in one pass."""
    _cabi.require_cuda(src)
    assert src.dtype == torch.float32 and out.dtype == torch.float32 and src.shape == out.shape and src.dim() == 4
    N0, N1, C, L = src.shape
    for t in (src, out):
        assert t.stride(3) == 1 and t.stride(2) == L, "only the two leading axes may be strided"
    if src.numel():
        with torch.cuda.device(src.device):
            code = _cabi.lib().gfx_node_copy_f32(src.data_ptr(), out.data_ptr(), N0, N1, C * L, src.stride(0),
                                                 src.stride(1), out.stride(0), out.stride(1), _cabi.stream_ptr())
        _cabi.check(code, "gfx_node_copy_f32")
    return out


def reverb_ir(noise_stft: torch.Tensor, init_log_magnitude: torch.Tensor, delta_log_magnitude: torch.Tensor,
              gain_env_log_magnitude: torch.Tensor | None, window: torch.Tensor, ir_len: int, n_fft: int,
              hop_length: int, finish: str = "unit"):
    """Masked-noise STFT -> impulse response [B, 2, ir_len]
    (reference: STFTMaskedNoiseReverb.compute_ir + _process_* prologue, reverb.py:161-228).

    finish = "unit": mid/side rows normalised to unit energy (normalize_impulse);
             "lr":   ms_to_lr, then normalised (pseudo_midside);
             "raw" / "raw_lr": returns (un-normalised mid/side | left/right response, energy[B, 2] of the
                     raw mid/side rows) for fir_conv_midside_ir, which folds the normalisation into the
                     filter spectra (the finished IR is never stored)."""
    _cabi.require_cuda(noise_stft, init_log_magnitude, delta_log_magnitude, window)
    _no_backward("reverb_ir", init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude)
    B = init_log_magnitude.shape[0]
    bins, frames = n_fft // 2 + 1, 1 + ir_len // hop_length
    assert tuple(init_log_magnitude.shape) == (B, 2, bins) and tuple(delta_log_magnitude.shape) == (B, 2, bins)
    assert noise_stft.is_complex() and tuple(noise_stft.shape[-3:]) == (2, bins, frames)
    nz = torch.view_as_real(noise_stft.detach().to(torch.complex64).contiguous())
    nb = noise_stft.shape[0] if noise_stft.ndim == 4 else 1
    assert nb in (1, B)
    bstride = 0 if nb == 1 else 2 * bins * frames
    h0, hd = _prep(init_log_magnitude, torch.float32), _prep(delta_log_magnitude, torch.float32)
    ge = None
    if gain_env_log_magnitude is not None:
        ge = _prep(gain_env_log_magnitude, torch.float32)
        assert tuple(ge.shape) == (B, 2, frames)
    win = _prep(window, torch.float32)
    mode = {"raw": 0, "unit": 1, "lr": 2, "raw_lr": 3}[finish]
    ir = torch.empty(B, 2, ir_len, dtype=torch.float32, device=h0.device)
    energy = torch.empty(B, 2, dtype=torch.float32, device=h0.device)
    if B == 0:
        return (ir, energy) if mode in (0, 3) else ir
    L_ = _cabi.lib()
    ws = _cabi.workspace(L_.gfx_reverb_ir_workspace_bytes(B, n_fft, hop_length, ir_len), h0.device)
    with torch.cuda.device(h0.device):
        code = L_.gfx_reverb_ir_f32(nz.data_ptr(), bstride, h0.data_ptr(), hd.data_ptr(), _cabi.ptr(ge),
                                    win.data_ptr(), ir.data_ptr(), energy.data_ptr(), ws.data_ptr(), ws.numel(),
                                    B, n_fft, hop_length, ir_len, mode, _cabi.stream_ptr())
    if code == -4:
        raise NotImplementedError("reverb IR synthesis supports n_fft=384 with hop_length=192, or a power-of-two n_fft in "
                                  "32..4096 with hop_length <= n_fft")
    _cabi.check(code, "gfx_reverb_ir_f32")
    return (ir, energy) if mode in (0, 3) else ir


def fir_conv_midside_ir(x: torch.Tensor, ir_raw: torch.Tensor, energy: torch.Tensor, to_lr: bool,
                        h_repeat: int = 1) -> torch.Tensor:
    """Causal convolution of x [B, 1|2, L] with a reverb response given un-normalised ([B, 2, N], rows
    mid/side, or left/right when to_lr) + the energies [B, 2] of its raw mid/side rows:
    normalize_impulse (reverb.py:215-228) is folded into the filter spectra."""
    _cabi.require_cuda(x, ir_raw, energy)
    _no_backward("fir_conv_midside_ir", x, ir_raw, energy)
    assert x.ndim == 3 and ir_raw.ndim == 3 and ir_raw.shape[1] == 2 and x.shape[0] == ir_raw.shape[0] * h_repeat
    B, cx, L = x.shape
    N = ir_raw.shape[2]
    assert cx in (1, 2), "channel mismatch between signal and filter"
    x, ir_raw, energy = _prep(x, torch.float32), _prep(ir_raw, torch.float32), _prep(energy, torch.float32)
    y = _new_output((B, 2, L), torch.float32, x.device)
    if y.numel():
        L_ = _cabi.lib()
        plan = _fft_plan(x.device, L_.gfx_fir_fft_size(N))
        ws = _cabi.workspace(L_.gfx_fir_conv_workspace_bytes(B, cx, 2, L, N, 0), x.device)
        with torch.cuda.device(x.device):
            code = L_.gfx_fir_conv_midside_ir_f32(x.data_ptr(), ir_raw.data_ptr(), energy.data_ptr(), y.data_ptr(), B,
                                                  cx, L, N, int(bool(to_lr)), int(h_repeat), plan.data_ptr(), ws.data_ptr(),
                                                  ws.numel(), _cabi.stream_ptr())
        _cabi.check(code, "gfx_fir_conv_midside_ir_f32")
    return y


DESIGN_FAMILY = {"peq": 0, "peaking": 1, "lowshelf": 2, "highshelf": 3, "lowpass": 4, "highpass": 5, "bandpass": 6,
                 "bandreject": 7, "allpass": 8, "stable": 9, "svf": 10}


def _design_torch(family: str, *params, flags: int = 0):
    """The PyTorch statement of the same formulas (processors/design.py): O(parameters) work, differentiable."""
    from .processors import design as D

    if family == "peq":
        return D.parametric_eq(*params, bool(flags & 1))
    if family in ("peaking", "lowshelf", "highshelf"):
        return D.eq_band(family, *params)
    if family in ("lowpass", "highpass", "bandpass", "bandreject", "allpass"):
        return D.simple_filter(family, *params)
    if family == "stable":
        Bs, a1, a2 = params[:3]
        a0 = params[3] if len(params) > 3 else None
        return D.stable_biquad(Bs, a1, a2, a0, bool(flags & 2))
    if family == "svf":
        return D.state_variable(*params)
    raise ValueError(family)


def biquad_design(family: str, *params: torch.Tensor, flags: int = 0):
    """Parameter activations -> (Bs, As) in one launch (reference: the coefficient designers of
    processors/filter.py and eq.py:300-314; same formulas as processors/design.py, which stays as
    the PyTorch statement of them).  All parameter tensors share the shape [..., K] (for "stable"
    the first one is [..., K, 3]); returns Bs, As of shape [..., K, 3]."""
    if _wants_grad(*params):
        return _design_torch(family, *params, flags=flags)
    fam = DESIGN_FAMILY[family]
    ref = params[1]
    _cabi.require_cuda(*[t for t in params if t is not None])
    K = ref.shape[-1]
    n_rows = ref.numel() // K
    ps = [None if t is None else _prep(t, torch.float32) for t in params]
    ps += [None] * (5 - len(ps))
    Bs = torch.empty(*ref.shape, 3, dtype=torch.float32, device=ref.device)
    As = torch.empty_like(Bs)
    if Bs.numel():
        with torch.cuda.device(ref.device):
            code = _cabi.lib().gfx_biquad_design_f32(fam, *[_cabi.ptr(t) for t in ps], Bs.data_ptr(), As.data_ptr(),
                                                     n_rows, K, int(flags), _cabi.stream_ptr())
        _cabi.check(code, "gfx_biquad_design_f32")
    return Bs, As


POINTWISE_OP = {"gain": 0, "side_gain": 1, "tanh": 2, "piecewise_tanh": 3, "power": 4, "chebyshev": 5, "scale_add": 6}


def row_mean(x: torch.Tensor) -> torch.Tensor:
    """Mean over time of every (batch, channel) row -> [B, C] (the `remove_dc` option of nonlinear.py)."""
    _cabi.require_cuda(x)
    assert x.ndim == 3
    if _wants_grad(x):  # O(1) flops per sample and its adjoint is a broadcast: the PyTorch statement carries the graph
        return x.to(torch.float32).mean(-1)
    x = _prep(x, torch.float32)
    B, C, L = x.shape
    m = torch.empty(B, C, dtype=torch.float32, device=x.device)
    if m.numel():
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib().gfx_row_mean_f32(x.data_ptr(), m.data_ptr(), B * C, L, _cabi.stream_ptr()), "gfx_row_mean_f32")
    return m


def mean_square(x: torch.Tensor) -> torch.Tensor:
    """Mean of x^2 over the channel and time axes -> [B] (x.square().mean((-1, -2)) of core/utils.py:8-9)."""
    _cabi.require_cuda(x)
    assert x.ndim == 3
    if _wants_grad(x):  # a training regulariser upstream (GainStagingRegularization): never cut the graph silently
        return x.to(torch.float32).square().mean((-1, -2))
    x = _prep(x, torch.float32)
    B, C, L = x.shape
    m = torch.empty(B, dtype=torch.float32, device=x.device)
    if m.numel():
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib().gfx_row_mean_square_f32(x.data_ptr(), m.data_ptr(), B, C * L, _cabi.stream_ptr()),
                        "gfx_row_mean_square_f32")
    return m


def rms_difference(X: torch.Tensor, Y: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """sum_b |log(mean X_b^2 + eps) - log(mean Y_b^2 + eps)|  (processors/core/utils.py:7-11): the two reductions
    over the audio run on the device in one pass each, the B-element tail in PyTorch."""
    return (torch.log(mean_square(X) + eps) - torch.log(mean_square(Y) + eps)).abs().sum()


def pointwise(op: str, x: torch.Tensor, p0=None, p1=None, p2=None, p3=None, dc=None, order: int = 0, flags: int = 0,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """Sample-wise processors (stereo.py, nonlinear.py, the ParallelMix accumulation) in one pass; see
    gfx_pointwise_f32 in include/grafx_b200.h for the parameter meaning of every op."""
    _cabi.require_cuda(x, *[t for t in (p0, p1, p2, p3, dc) if t is not None])
    if op == "gain" and out is None and dc is None and _wants_grad(x, p0):
        from .autograd import GainFn

        return GainFn.apply(x.to(torch.float32).contiguous(), p0.to(torch.float32).contiguous())
    if _wants_grad(x, p0, p1, p2, p3, dc):
        from . import training

        if op == "scale_add":  # ParallelMix accumulation (container.py:203-216)
            term = p0.reshape(-1, 1, 1) * x
            return term if (out is None or not (flags & 4)) else out + term
        return training.pointwise(op, x, p0, p1, p2, p3, dc, order, flags)
    assert x.ndim == 3
    x = _prep(x, torch.float32)
    B, C, L = x.shape
    keep = [None if t is None else _prep(t, torch.float32) for t in (p0, p1, p2, p3, dc)]
    for t in keep[:4]:
        assert t is None or t.shape[0] == B, "parameter batch size must match the signal"
    y = out if out is not None else _new_output((B, C, L), torch.float32, x.device)
    if y.numel():
        with torch.cuda.device(x.device):
            code = _cabi.lib().gfx_pointwise_f32(POINTWISE_OP[op], x.data_ptr(), y.data_ptr(), B, C, L,
                                                 *[_cabi.ptr(t) for t in keep], int(order), int(flags), _cabi.stream_ptr())
        _cabi.check(code, "gfx_pointwise_f32")
    return y


def noise_shaping_ir(noise: torch.Tensor, offset: int, decay: torch.Tensor, gain: torch.Tensor, fade=None, fade_gain=None,
                     ir_len: int = 0):
    """FilteredNoiseShapingReverb impulse response (reverb.py:364-380): band-filtered noise [C, K, T_noise] shaped by
    per-band exponential envelopes (already activated log-slopes / gains [B, C, K]).  Returns (ir [B, C, ir_len]
    un-normalised, energy [B, C])."""
    _cabi.require_cuda(noise, decay, gain)
    _no_backward("noise_shaping_ir", decay, gain, fade, fade_gain)
    assert noise.ndim == 3 and decay.ndim == 3 and decay.shape == gain.shape and decay.shape[1:] == noise.shape[:2]
    noise = _prep(noise, torch.float32)
    B, C, K = decay.shape
    ps = [_prep(decay, torch.float32), _prep(gain, torch.float32), None if fade is None else _prep(fade, torch.float32),
          None if fade_gain is None else _prep(fade_gain, torch.float32)]
    ir = torch.empty(B, C, ir_len, dtype=torch.float32, device=noise.device)
    energy = torch.empty(B, C, dtype=torch.float32, device=noise.device)
    if ir.numel():
        L_ = _cabi.lib()
        ws = _cabi.workspace(L_.gfx_noise_shaping_ir_workspace_bytes(B, C, ir_len), noise.device)
        with torch.cuda.device(noise.device):
            code = L_.gfx_noise_shaping_ir_f32(noise.data_ptr(), noise.shape[2], int(offset), *[_cabi.ptr(t) for t in ps],
                                               ir.data_ptr(), energy.data_ptr(), ws.data_ptr(), ws.numel(), B, C, K, ir_len,
                                               _cabi.stream_ptr())
        _cabi.check(code, "gfx_noise_shaping_ir_f32")
    return ir, energy
