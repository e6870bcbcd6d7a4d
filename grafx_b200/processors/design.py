"""Parameter activations and biquad coefficient design (O(parameters) work, stays in PyTorch on
the device).  Formulas follow the reference CODE, not its docstrings:
  grafx/processors/filter.py:144-154 (BiquadFilter), :218-238 (PoleZeroFilter), :303-338 (SVF),
  :373-383 / :592-604 (w0, 1/q, gain activations), :416-556 (LP/HP/BP/BR/AP),
  :645-656 / :687-705 / :736-754 (peaking / low shelf / high shelf), eq.py:300-314 (PEQ layout).
Every designer returns (Bs, As) with the three taps stacked on the last axis.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

_LN2 = math.log(2.0)


def _taps(*coeffs):
    return torch.stack(coeffs, dim=-1)


def cutoff_and_alpha(w0, q_inv):
    """w0 = pi*sigmoid(.), 1/q = exp(.), alpha = sin(w0)/(2q).  Returns (cos w0, alpha)."""
    w = math.pi * torch.sigmoid(w0)
    return torch.cos(w), torch.sin(w) * torch.exp(q_inv) * 0.5


def _common_denominator(c, alpha):
    return _taps(1 + alpha, -2 * c, 1 - alpha)


def simple_filter(kind: str, w0, q_inv):
    c, alpha = cutoff_and_alpha(w0, q_inv)
    den = _common_denominator(c, alpha)
    if kind == "lowpass":  # numerator sign as shipped: (cos w0 - 1)/2
        h = (c - 1) * 0.5
        num = _taps(h, c - 1, h)
    elif kind == "highpass":
        h = (1 + c) * 0.5
        num = _taps(h, -(1 + c), h)
    elif kind == "bandpass":
        num = _taps(alpha, torch.zeros_like(alpha), -alpha)
    elif kind == "bandreject":
        one = torch.ones_like(c)
        num = _taps(one, -2 * c, one)
    elif kind == "allpass":
        num = den.flip(-1)
    else:
        raise ValueError(kind)
    return num, den


def _shelf(c, alpha, A, sign):
    """sign=+1: low shelf, sign=-1: high shelf."""
    ap, am = A + 1, A - 1
    root = 2 * torch.sqrt(A) * alpha
    num = _taps(A * (ap - sign * am * c + root), sign * 2 * A * (am - sign * ap * c), A * (ap - sign * am * c - root))
    den = _taps(ap + sign * am * c + root, -sign * 2 * (am + sign * ap * c), ap + sign * am * c - root)
    return num, den


def peaking(c, alpha, A):
    return _taps(1 + alpha * A, -2 * c, 1 - alpha * A), _taps(1 + alpha / A, -2 * c, 1 - alpha / A)


def low_shelf(c, alpha, A):
    return _shelf(c, alpha, A, +1.0)


def high_shelf(c, alpha, A):
    return _shelf(c, alpha, A, -1.0)


def eq_band(kind: str, w0, q_inv, log_gain):
    c, alpha = cutoff_and_alpha(w0, q_inv)
    A = torch.exp(log_gain)
    return {"peaking": peaking, "lowshelf": low_shelf, "highshelf": high_shelf}[kind](c, alpha, A)


def parametric_eq(w0, q_inv, log_gain, use_shelving_filters=True):
    """[.., K] -> [.., K, 3] x2.  Band 0 low shelf, 1..K-2 peaking, K-1 high shelf."""
    c, alpha = cutoff_and_alpha(w0, q_inv)
    A = torch.exp(log_gain)
    if not use_shelving_filters:
        return peaking(c, alpha, A)
    K = w0.shape[-1]
    parts = []
    for sl, fn in ((slice(0, 1), low_shelf), (slice(1, K - 1), peaking), (slice(K - 1, K), high_shelf)):
        parts.append(fn(c[..., sl], alpha[..., sl], A[..., sl]))
    return torch.cat([p[0] for p in parts], -2), torch.cat([p[1] for p in parts], -2)


def stable_biquad(Bs, A1_pre, A2_pre, A0=None, scale_by_a0=False):
    """Stability-constrained direct coefficients (filter.py:144-154).  NB upstream multiplies by
    A0 when its `normalized` flag is True (inverted w.r.t. its docstring); kept as shipped."""
    a1 = 2 * torch.tanh(A1_pre)
    mag = a1.abs()
    a2 = ((2 - mag) * torch.tanh(A2_pre) + mag) * 0.5
    den = _taps(torch.ones_like(a1), a1, a2)
    if scale_by_a0:
        den = den * A0.unsqueeze(-1)
    num = torch.cat([Bs[..., :1] + 1, Bs[..., 1:]], -1)
    return num, den


def state_variable(twoR, G, c_hp, c_bp, c_lp):
    g = torch.tan(0.5 * math.pi * torch.sigmoid(G))
    r2 = (1.0 / _LN2) * F.softplus(twoR) + 1e-2
    g2 = g * g
    num = _taps(c_hp + c_bp * g + c_lp * g2, -c_hp * 2 + c_lp * 2 * g2, c_hp - c_bp * g + c_lp * g2)
    den = _taps(1 + g2 + r2 * g, 2 * g2 - 2, 1 + g2 - r2 * g)
    return num, den


def pole_zero(poles, zeros):
    """Conjugate pole/zero pairs -> biquads (filter.py:218-238).  As shipped, a2 uses the pole
    radius BEFORE the tanh re-parameterisation."""
    p = torch.view_as_complex(poles.contiguous())
    rp = p.abs()
    p = p * torch.tanh(rp) / (rp + 1e-5)
    z = torch.view_as_complex(zeros.contiguous())
    one = torch.ones_like(rp)
    return _taps(one, -2 * z.real, z.abs().square()), _taps(one, -2 * p.real, rp.square())
