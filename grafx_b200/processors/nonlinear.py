"""Memoryless distortions -- drop-ins for grafx.processors.nonlinear (nonlinear.py:6-413).

Same class names, constructor kwargs, forward signatures and parameter_size(); one streaming CUDA pass each
(csrc/pointwise.cu), plus a per-row mean reduction when `remove_dc` is set."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F_

_INVERSE_POST, _USE_TANH = 1, 2


class TanhDistortion(nn.Module):
    """nonlinear.py:6-103."""

    def __init__(self, pre_post_gain=True, inverse_post_gain=True, remove_dc=False, use_bias=False):
        super().__init__()
        self.pre_post_gain = pre_post_gain
        self.inverse_post_gain = inverse_post_gain
        self.remove_dc = remove_dc
        self.use_bias = use_bias

    def forward(self, input_signals, log_pre_gain=None, log_post_gain=None, bias=None):
        dc = F_.row_mean(input_signals) if self.remove_dc else None
        pre = log_pre_gain if self.pre_post_gain else None
        inverse = self.pre_post_gain and self.inverse_post_gain
        post = log_post_gain if (self.pre_post_gain and not self.inverse_post_gain) else None
        return F_.pointwise("tanh", input_signals, pre, post, bias if self.use_bias else None, dc=dc,
                            flags=_INVERSE_POST if inverse else 0)

    def parameter_size(self):
        size = {}
        if self.pre_post_gain:
            size["log_pre_gain"] = 1
            if not self.inverse_post_gain:
                size["log_post_gain"] = 1
        if self.use_bias:
            size["bias"] = 1
        return size


class PiecewiseTanhDistortion(nn.Module):
    """nonlinear.py:106-218: tanh between the thresholds, scaled tanh lobes of adjustable hardness beyond them."""

    def __init__(self, pre_post_gain=True, inverse_post_gain=True, remove_dc=False):
        super().__init__()
        self.pre_post_gain = pre_post_gain
        self.inverse_post_gain = inverse_post_gain
        self.remove_dc = remove_dc

    def forward(self, input_signals, log_hardness, z_threshold, log_pre_gain=None, log_post_gain=None):
        dc = F_.row_mean(input_signals) if self.remove_dc else None
        pre = log_pre_gain if self.pre_post_gain else None
        inverse = self.pre_post_gain and self.inverse_post_gain
        post = log_post_gain if (self.pre_post_gain and not self.inverse_post_gain) else None
        return F_.pointwise("piecewise_tanh", input_signals, log_hardness, z_threshold, pre, post, dc=dc,
                            flags=_INVERSE_POST if inverse else 0)

    def parameter_size(self):
        size = {"log_hardness": 2, "z_threshold": 2}
        if self.pre_post_gain:
            size["log_pre_gain"] = 1
            if not self.inverse_post_gain:
                size["log_post_gain"] = 1
        return size


class _SeriesDistortion(nn.Module):
    op = None

    def __init__(self, max_order=10, pre_gain=True, remove_dc=False, use_tanh=False):
        super().__init__()
        assert max_order > 1
        self.pre_gain = pre_gain
        self.max_order = max_order
        self.remove_dc = remove_dc
        self.use_tanh = use_tanh
        if self.op == "power":  # (upstream buffer, nonlinear.py:268-270; the kernel forms the powers in registers)
            self.register_buffer("arange", torch.arange(max_order)[:, None, None, None])

    def forward(self, input_signals, basis_weights, log_pre_gain=None):
        dc = F_.row_mean(input_signals) if self.remove_dc else None
        return F_.pointwise(self.op, input_signals, basis_weights, log_pre_gain if self.pre_gain else None, dc=dc,
                            order=self.max_order, flags=_USE_TANH if self.use_tanh else 0)

    def parameter_size(self):
        size = {"basis_weights": self.max_order}
        if self.pre_gain:
            size["log_pre_gain"] = 1
        return size


class PowerDistortion(_SeriesDistortion):
    """Weighted sum of the monomials x^k, k < max_order, weights tanh(basis_weights) (nonlinear.py:221-299)."""

    op = "power"


class ChebyshevDistortion(_SeriesDistortion):
    """Weighted sum of the Chebyshev polynomials T_k(x), k < max_order (nonlinear.py:302-413)."""

    op = "chebyshev"
