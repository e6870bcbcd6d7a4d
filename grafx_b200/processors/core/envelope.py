"""Stand-alone envelope smoothers -- drop-ins for grafx.processors.core.envelope
(core/envelope.py:10-101): TruncatedOnePoleIIRFilter and Ballistics as modules of their own.

Both run on the fused dynamics kernel (csrc/dynamics.cu, gfx_envelope_f32): the truncated one-pole response
h[n] = (1 - a) a^n, n < iir_len, is evaluated by the exact recursion T[n] = a T[n-1] + (1 - a)(u[n] - a^N u[n-N])
(then relu, as upstream) instead of a 16384-tap FFT convolution; Ballistics walks the attack / release
recursion of torchcomp.compressor_core (y[-1] = 1; coefficient `at` when u[t] < y[t-1], `rt` otherwise)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import functional as F_


class TruncatedOnePoleIIRFilter(nn.Module):
    """core/envelope.py:10-60.  forward(input_signals [B, L], z_alpha [B, 1]) -> [B, L]."""

    def __init__(self, iir_len=16384, **backend_kwargs):
        super().__init__()
        self.iir_len = iir_len
        self.register_buffer("arange", torch.arange(iir_len)[None, :])  # (kept for upstream checkpoints)

    def forward(self, input_signals, z_alpha):
        return F_.envelope(input_signals, z_alpha, "iir", iir_len=self.iir_len)

    def compute_impulse(self, z_alpha):
        """The truncated impulse response itself (core/envelope.py:51-60), O(parameters x iir_len) in PyTorch."""
        alpha = torch.clamp(torch.sigmoid(z_alpha), max=1 - 1e-5)
        return (1 - alpha) * torch.exp(self.arange * torch.log(alpha))

    def parameter_size(self):
        return {"z_alpha": 1}


class Ballistics(nn.Module):
    """core/envelope.py:63-101.  forward(input_signals [B, L], z_alpha [B, 2] = (attack, release)) -> [B, L]."""

    def __init__(self):
        super().__init__()

    def forward(self, input_signals, z_alpha):
        return F_.envelope(input_signals, z_alpha, "ballistics")

    def parameter_size(self):
        return {"z_alpha": 2}
