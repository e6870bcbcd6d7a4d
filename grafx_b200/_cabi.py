"""ctypes binding of the C ABI in include/grafx_b200.h (the only way the Python host code
reaches the CUDA kernels).  There is NO CPU fallback: if the shared library is missing and
cannot be built, or a call returns an error code, this module raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GRAFX_B200_LIB") or os.path.join(_HERE, "lib", "libgrafx_b200.so")  # (override: kernel A/B experiments)

_lib = None
_lock = threading.Lock()

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_size_t = ctypes.c_size_t
c_float = ctypes.c_float


class DynamicsStage(ctypes.Structure):
    """gfx_dynamics_stage (include/grafx_b200.h)."""

    _fields_ = [
        ("kind", c_int), ("knee", c_int), ("energy_smoother", c_int), ("gain_smoother", c_int),
        ("gain_smooth_in_log", c_int), ("reserved", c_int),
        ("log_threshold", c_void_p), ("log_ratio", c_void_p), ("log_knee", c_void_p),
        ("z_alpha_pre", c_void_p), ("z_alpha_post", c_void_p),
        ("hist_pre", c_void_p), ("hist_post", c_void_p),
    ]


# name -> (restype, argtypes); must list every symbol declared in include/grafx_b200.h
SIGNATURES = {
    "gfx_version": (c_int, []),
    "gfx_last_cuda_error": (c_int, []),
    "gfx_error_string": (ctypes.c_char_p, [c_int]),
    "gfx_kernel_launch_count": (ctypes.c_ulonglong, []),
    "gfx_device_sm_count": (c_int, []),
    "gfx_fma_probe_f32": (c_ll, [c_void_p, c_int, c_void_p]),
    "gfx_biquad_cascade_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "gfx_biquad_cascade_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_int, c_ll, c_void_p, c_size_t, c_void_p]),
    "gfx_biquad_cascade_ex_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_ll,
                                          c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "gfx_biquad_cascade_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_int, c_ll, c_void_p, c_size_t, c_void_p]),
    "gfx_dynamics_workspace_bytes": (c_size_t, [c_int, c_int]),
    "gfx_dynamics_set_tuning": (c_int, [c_int]),
    "gfx_dynamics_set_ballistics_mode": (c_int, [c_int]),
    "gfx_dynamics_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_ll, ctypes.POINTER(DynamicsStage), c_int,
                                 c_int, c_void_p, c_size_t, c_void_p]),
    "gfx_dynamics_rep_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_ll, ctypes.POINTER(DynamicsStage), c_int,
                                     c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "gfx_envelope_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_ll, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                                 c_size_t, c_void_p]),
    "gfx_fir_fft_size": (c_int, [c_int]),
    "gfx_fft_plan_bytes": (c_size_t, [c_int]),
    "gfx_fft_plan_init": (c_int, [c_void_p, c_int, c_void_p]),
    "gfx_fir_conv_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_ll, c_int, c_int]),
    "gfx_fir_conv_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_ll, c_int, c_int, c_int, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "gfx_fir_filter_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_ll, c_int, c_void_p, c_void_p, c_size_t,
                                   c_void_p]),
    "gfx_fir_set_tuning": (c_int, [c_int, c_int]),
    "gfx_fir_set_long_mode": (c_int, [c_int, c_int]),
    "gfx_fir_set_mac_form": (c_int, [c_int]),
    "gfx_fir_set_sweep_mb": (c_int, [c_int]),
    "gfx_fir_conv_midside_ir_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_int, c_int,
                                            c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gfx_reverb_ir_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "gfx_reverb_ir_f32": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "gfx_noise_shaping_ir_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gfx_noise_shaping_ir_f32": (c_int, [c_void_p, c_ll, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_void_p]),
    "gfx_drywet_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_void_p]),
    "gfx_node_sum_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_ll,
                                 c_void_p]),
    "gfx_node_copy_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_ll, c_void_p]),
    "gfx_lag_dots_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p]),
    "gfx_biquad_design_f32": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_int, c_void_p]),
    "gfx_row_mean_f32": (c_int, [c_void_p, c_void_p, c_int, c_ll, c_void_p]),
    "gfx_row_mean_square_f32": (c_int, [c_void_p, c_void_p, c_int, c_ll, c_void_p]),
    "gfx_pointwise_f32": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int, c_int, c_void_p]),
    "gfx_midside_f32": (c_int, [c_void_p, c_void_p, c_int, c_ll, c_float, c_void_p]),
}


class GrafxB200Error(RuntimeError):
    pass


def lib():
    """Loads (building first if the sources are newer / the .so is absent) the CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.environ.get("GRAFX_B200_LIB"):
            from . import build as _build

            try:
                stale = _build.needs_rebuild()
            except OSError:  # sources not shipped next to a prebuilt library
                stale = not os.path.exists(LIB_PATH)
            if stale:
                try:
                    _build.build()
                except Exception:
                    if not os.path.exists(LIB_PATH):
                        raise
        if not os.path.exists(LIB_PATH):
            raise GrafxB200Error(
                f"{LIB_PATH} is missing and could not be built; grafx_b200 has no CPU fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        return _lib


def check(code: int, what: str):
    if code != 0:
        L = lib()
        msg = L.gfx_error_string(code).decode()
        extra = f" (cudaError {L.gfx_last_cuda_error()})" if code == -3 else ""
        raise GrafxB200Error(f"{what}: {msg}{extra}")


def require_cuda(*tensors: torch.Tensor):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise GrafxB200Error(
                "grafx_b200 processes CUDA tensors only (no CPU fallback); got a tensor on "
                f"{t.device}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)
