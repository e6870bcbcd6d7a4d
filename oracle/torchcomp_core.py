"""TEST INFRASTRUCTURE ONLY -- never imported by the product (grafx_b200/).

The attack/release recursion behind `torchcomp.compressor_core`, the third-party call the reference makes
in Ballistics.forward (/root/reference/src/grafx/processors/core/envelope.py:5,97-100:
`compressor_core(input_signals, zi=ones, at, rt)`; dependency `torchcomp`, unpinned in the reference's
pyproject.toml:19, absent from /root/reference and from this image -- no wheel in /opt/wheelhouse, no network).

Source restated: torchcomp (github.com/DiffAPF/torchcomp, "Differentiable All-pole Filters for Time-varying
Audio Systems", Yu et al. 2024), file `torchcomp/core.py`, the CPU kernel `compressor_kernel` (numba
`@njit(parallel=True)`) and its CUDA twin `compressor_cuda_kernel`: per batch row

    g = zi[b]
    for t in range(T):
        f = x[b, t]
        flag = f < g
        coeff = at[b] if flag else rt[b]      # "attack" when the input falls below the state
        g *= 1 - coeff
        g += coeff * f
        y[b, t] = g

The text above is written down from the published algorithm (the package cannot be fetched here, so the exact
upstream commit cannot be quoted); `compressor_core` below is that loop, compiled with numba when it is importable.
What pins it besides the text (tests/test_oracle_golden.py::test_ballistics_*):
  * at == rt  =>  the recursion is the linear one-pole y[t] = (1 - a) y[t-1] + a x[t] with y[-1] = zi, checked
    against scipy.signal.lfilter with that initial state (reference-side invariant: no branch involved);
  * an input that stays below (above) the state uses only `at` (`rt`): each branch alone is the same one-pole;
  * the reference's own call convention (zi = 1, at = sigmoid(z[..., 0]), rt = sigmoid(z[..., 1])) through the
    unmodified Ballistics module with this function installed as `torchcomp.compressor_core`
    (oracle/ref_loader.py) -- that is how every `*_ballistics_*` fixture in tests/golden/ was produced.
"""
from __future__ import annotations

import numpy as np

try:  # numba is in the image; the pure-Python loop is the same arithmetic
    from numba import njit, prange

    _HAVE_NUMBA = True
except Exception:  # pragma: no cover
    _HAVE_NUMBA = False

    def njit(*a, **k):
        def wrap(f):
            return f
        return wrap

    prange = range


@njit(parallel=True, cache=False)
def compressor_kernel(x, zi, at, rt):
    B, T = x.shape
    y = np.empty_like(x)
    at_mask = np.zeros(x.shape, dtype=np.bool_)
    for b in prange(B):
        g = zi[b]
        at_b = at[b]
        rt_b = rt[b]
        for t in range(T):
            f = x[b, t]
            flag = f < g
            if flag:
                coeff = at_b
                at_mask[b, t] = 1
            else:
                coeff = rt_b
            g *= 1 - coeff
            g += coeff * f
            y[b, t] = g
    return y, at_mask


def compressor_core(x, zi, at, rt):
    """torch-in / torch-out wrapper with the signature the reference calls (core/envelope.py:100)."""
    import torch

    xn = np.ascontiguousarray(x.detach().cpu().numpy())
    dt = xn.dtype
    y, _ = compressor_kernel(xn, zi.detach().cpu().numpy().astype(dt), at.detach().cpu().numpy().astype(dt),
                             rt.detach().cpu().numpy().astype(dt))
    return torch.from_numpy(y).to(x.device)
