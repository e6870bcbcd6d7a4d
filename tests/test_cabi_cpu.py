"""CPU: the C-ABI shared library builds, loads and exports every symbol include/grafx_b200.h
declares (no compute calls without a GPU); the Python binding lists exactly those symbols."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "grafx_b200.h")).read()
    return sorted(set(re.findall(r"GFX_API\s+[\w\s\*]+?\b(gfx_\w+)\s*\(", text)))


def test_header_declares_symbols():
    syms = declared_symbols()
    assert "gfx_biquad_cascade_f32" in syms and "gfx_version" in syms


def test_library_exports_every_declared_symbol():
    from grafx_b200 import _cabi, build

    path = build.build()
    handle = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert sorted(_cabi.SIGNATURES) == declared_symbols()
    L = _cabi.lib()
    assert L.gfx_version() >= 100
    assert L.gfx_error_string(-2).decode().startswith("workspace")


def test_argument_validation_without_gpu():
    from grafx_b200 import _cabi

    L = _cabi.lib()
    # null pointers / bad sizes are rejected before any CUDA call
    assert L.gfx_biquad_cascade_f32(None, None, None, None, 1, 1, 1, 1, 16, None, 0, None) == -1
    assert L.gfx_biquad_cascade_workspace_bytes(4, 2, 2, 5, 4) >= 4 * 2 * 2 * 5 * 4


def test_no_cpu_fallback():
    import torch
    from grafx_b200 import functional as F_
    from grafx_b200._cabi import GrafxB200Error

    with pytest.raises(GrafxB200Error):
        F_.biquad_cascade(torch.zeros(1, 1, 8), torch.ones(1, 1, 1, 3), torch.ones(1, 1, 1, 3))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "grafx_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dirpath, f)
