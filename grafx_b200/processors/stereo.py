"""Stereo utilities -- drop-ins for grafx.processors.stereo (stereo.py:9-205).

Same class names, forward signatures and parameter_size(); the sample loops run in csrc/pointwise.cu
(gain, side gain) and csrc/elementwise.cu (mid/side butterflies)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import functional as F_


class StereoGain(nn.Module):
    """Per-channel gain exp(log_gain) (stereo.py:9-41)."""

    def forward(self, input_signals, log_gain):
        assert input_signals.ndim == 3
        if input_signals.shape[1] == 1 and log_gain.shape[-1] == 2:
            # mono in, stereo out: upstream's `input * exp(log_gain)[:, :, None]` broadcasts the channel axis (stereo.py:38-41)
            input_signals = input_signals.expand(-1, 2, -1)
        assert log_gain.shape[-1] == input_signals.shape[1]
        return F_.pointwise("gain", input_signals, log_gain)

    def parameter_size(self):
        return {"log_gain": 2}


class SideGainImager(nn.Module):
    """Scales the side channel of the mid/side decomposition by exp(log_gain) (stereo.py:44-87)."""

    def forward(self, input_signals, log_gain):
        b, c, t = input_signals.shape
        assert c == 2
        return F_.pointwise("side_gain", input_signals, log_gain.reshape(b, -1)[:, :1])

    def parameter_size(self):
        return {"log_gain": 1}


class MonoToStereo(nn.Module):
    """Repeats a mono signal on two channels (stereo.py:90-115); a plain device copy."""

    def forward(self, input_signals):
        b, c, t = input_signals.shape
        assert c == 1
        return input_signals.repeat(1, 2, 1)

    def parameter_size(self):
        return {}


class StereoToMidSide(nn.Module):
    """(left, right) -> (mid, side) = (l + r, l - r) [/ sqrt(2) if normalize] as two mono signals (stereo.py:118-154)."""

    def __init__(self, normalize=True):
        super().__init__()
        self.normalize = normalize

    def forward(self, input_signals):
        _, c, _ = input_signals.shape
        assert c == 2
        ms = F_.lr_to_ms(input_signals, mult=1.0 / math.sqrt(2) if self.normalize else 1.0)
        return ms[:, :1, :], ms[:, 1:, :]

    def parameter_size(self):
        return {}


class MidSideToStereo(nn.Module):
    """(mid, side) -> (mid + side, mid - side) * (1/sqrt(2) if normalize else 1/2) (stereo.py:157-205)."""

    def __init__(self, normalize=True):
        super().__init__()
        self.normalization_const = 1.0 / math.sqrt(2) if normalize else 0.5

    def forward(self, mid, side):
        b, c, t = mid.shape
        assert c == 1
        return F_.lr_to_ms(torch.cat([mid, side], 1), mult=self.normalization_const)

    def parameter_size(self):
        return {}
