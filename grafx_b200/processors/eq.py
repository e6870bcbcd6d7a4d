"""Equalizers -- drop-ins for grafx.processors.eq (eq.py:217-336 ParametricEqualizer)."""
from __future__ import annotations

import torch.nn as nn

from .. import functional as F_
from .core.iir import IIRFilter


class ParametricEqualizer(nn.Module):
    """K-band PEQ: low shelf, K-2 peaking bands, high shelf (or all peaking), in
    mono / stereo / mid-side channel modes.  Parameters w0, q_inv, log_gain: [N, n_ch, K]."""

    def __init__(self, num_filters=10, processor_channel="mono", use_shelving_filters=True, **backend_kwargs):
        super().__init__()
        if processor_channel not in ("mono", "stereo", "midside"):
            raise ValueError(f"Invalid processor_channel: {processor_channel}")
        self.num_filters = num_filters
        self.use_shelving_filters = use_shelving_filters
        self.processor_channel = processor_channel
        self.biquad = IIRFilter(order=2, **backend_kwargs)

    def forward(self, input_signals, w0, q_inv, log_gain):
        Bs, As = F_.biquad_design("peq", w0, q_inv, log_gain, flags=int(self.use_shelving_filters))
        if self.processor_channel == "midside":
            return F_.ms_to_lr(self.biquad(F_.lr_to_ms(input_signals), Bs, As))
        return self.biquad(input_signals, Bs, As)

    def parameter_size(self):
        n_channels = 1 if self.processor_channel == "mono" else 2
        return {k: (n_channels, self.num_filters) for k in ("w0", "q_inv", "log_gain")}
