"""TEST INFRASTRUCTURE ONLY -- never imported by the product (grafx_b200/).

Imports the *unmodified* reference package from /root/reference/src in THIS build
container so that (a) the oracle restatement in oracle/grafx_oracle.py can be pinned
against the reference's own code and (b) tests/golden/*.npz can be generated
(oracle/make_golden.py).  /root/reference does not exist on the GPU box; nothing that
runs there may call this module.

The reference needs four third-party modules that are not installable here
(SURVEY.md R5).  They are replaced by minimal stand-ins registered in sys.modules:

  torch_geometric.utils.{scatter, sort_edge_index}   call sites render/core.py:3,106,
                                                     render/prepare.py:5,115-119,
                                                     render/order/tensor.py:4,94,159,198
  torchlpc.sample_wise_lpc                           core/iir.py:11,282
  torchcomp.compressor_core                          core/envelope.py:5,100
  matplotlib{,.pyplot,.patches}                      grafx/__init__.py:1 -> draw/

and one documented deviation, the *even-pad guard* (SURVEY.md R1): the reference's
`convolve` (core/convolution.py:119-134) calls irfft without n=, which is only a linear
convolution when Lx+Lh-1 is even.  With `even_pad_guard=True` compute_pad_len is
replaced by a version that rounds the pad length up to the next even number; everything
else is the reference's own code.
"""
from __future__ import annotations

import importlib
import sys
import types

import numpy as np
import torch

REFERENCE_SRC = "/root/reference/src"


# --------------------------------------------------------------------------- shims
def _scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    """torch_geometric.utils.scatter restated with Tensor.scatter_reduce_."""
    dim = dim if dim >= 0 else src.dim() + dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.shape)
    size[dim] = dim_size
    view = [1] * src.dim()
    view[dim] = -1
    idx = index.view(view).expand_as(src)
    if reduce in ("sum", "add"):
        return src.new_zeros(size).scatter_add_(dim, idx, src)
    red = {"mul": "prod", "min": "amin", "max": "amax", "mean": "mean"}[reduce]
    init = {"prod": 1, "amin": 0, "amax": 0, "mean": 0}[red]
    out = src.new_full(size, init)
    return out.scatter_reduce_(dim, idx, src, reduce=red, include_self=(red == "prod"))


def _sort_edge_index(edge_index, edge_attr=None, num_nodes=None, sort_by_row=True):
    key = edge_index[0] if sort_by_row else edge_index[1]
    other = edge_index[1] if sort_by_row else edge_index[0]
    n = int(edge_index.max()) + 1 if edge_index.numel() else 1
    perm = torch.argsort(key * n + other, stable=True)
    ei = edge_index[:, perm]
    if edge_attr is None:
        return ei
    return ei, edge_attr[perm]


def _sample_wise_lpc(x, a, zi=None):
    """y[t] = x[t] - sum_k a[t,k] * y[t-1-k]   (torchlpc>=0.4 semantics, complex-capable)."""
    B, T = x.shape
    order = a.shape[-1]
    y = torch.zeros(B, T + order, dtype=torch.result_type(x, a), device=x.device)
    if zi is not None:
        y[:, :order] = zi.flip(-1)
    for t in range(T):
        hist = y[:, t : t + order].flip(-1)
        y[:, t + order] = x[:, t] - (a[:, t, :] * hist).sum(-1)
    return y[:, order:]


def compressor_core_loop(x, zi, at, rt):
    """torchcomp.compressor_core (the stand-in installed for the reference's import): c = at if x[t] < y[t-1]
    else rt; y[t] = (1-c) y[t-1] + c x[t]; y[-1] = zi -- the upstream kernel restated in oracle/torchcomp_core.py,
    vectorised over the batch (bit-identical in float64: tests/test_oracle_golden.py::test_ballistics_*)."""
    x_np = x.detach().cpu().numpy()
    y = np.empty_like(x_np)
    atn, rtn = at.detach().cpu().numpy(), rt.detach().cpu().numpy()
    prev = zi.detach().cpu().numpy().astype(x_np.dtype).copy()
    one = x_np.dtype.type(1)
    for t in range(x_np.shape[1]):
        u = x_np[:, t]
        c = np.where(u < prev, atn, rtn).astype(x_np.dtype)
        prev = (one - c) * prev + c * u
        y[:, t] = prev
    return torch.from_numpy(y).to(x.device)


def _install_shims():
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tgu = types.ModuleType("torch_geometric.utils")
        tgu.scatter = _scatter
        tgu.sort_edge_index = _sort_edge_index
        tg.utils = tgu
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.utils"] = tgu
    if "torchlpc" not in sys.modules:
        m = types.ModuleType("torchlpc")
        m.sample_wise_lpc = _sample_wise_lpc
        sys.modules["torchlpc"] = m
    if "torchcomp" not in sys.modules:
        m = types.ModuleType("torchcomp")
        m.compressor_core = compressor_core_loop
        sys.modules["torchcomp"] = m
    try:
        import matplotlib  # noqa: F401
    except Exception:
        class _Anything:
            def __init__(self, *a, **k):
                pass

            def __call__(self, *a, **k):
                return _Anything()

            def __getattr__(self, name):
                return _Anything()

            def __iter__(self):
                return iter(())

            def __getitem__(self, i):
                return _Anything()

        class _Mod(types.ModuleType):
            def __getattr__(self, name):
                if name.startswith("__"):
                    raise AttributeError(name)
                return _Anything()

        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                     "matplotlib.colors", "matplotlib.cm", "matplotlib.path",
                     "matplotlib.lines", "matplotlib.collections"):
            sys.modules[name] = _Mod(name)


def _even_pad_len(x, y, pad_mode="min"):
    n = x.shape[-1] + y.shape[-1] - 1
    return n + (n & 1)


def load_reference(even_pad_guard: bool = True):
    """Returns the imported reference package `grafx` (shimmed as documented above)."""
    _install_shims()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    grafx = importlib.import_module("grafx")
    conv = importlib.import_module("grafx.processors.core.convolution")
    if not hasattr(conv, "_compute_pad_len_as_shipped"):
        conv._compute_pad_len_as_shipped = conv.compute_pad_len
    conv.compute_pad_len = _even_pad_len if even_pad_guard else conv._compute_pad_len_as_shipped
    return grafx
