"""Filter processors -- drop-ins for grafx.processors.filter (filter.py:20-754).

Same class names, constructor kwargs, `forward(input_signals, **params)` signatures and
`parameter_size()` as the reference.  Coefficient design is O(parameters) PyTorch
(processors/design.py); the O(samples) filtering runs in the CUDA kernels behind IIRFilter.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F_
from . import design
from .core.convolution import FIRConvolution
from .core.iir import IIRFilter


class FIRFilter(nn.Module):
    """Learnable FIR filter (filter.py:20-84): tanh -> unit-energy normalisation -> causal
    convolution, in mono / stereo / mid-side modes.  The upstream constructor raises (it forwards
    `fir_len` to FIRConvolution and reads `processor_channel` before storing it);
    this class implements the documented signature."""

    def __init__(self, fir_len=1023, processor_channel="mono", **backend_kwargs):
        super().__init__()
        if processor_channel not in ("mono", "stereo", "midside"):
            raise ValueError(f"Unknown channel type: {processor_channel}")
        self.fir_len = fir_len
        self.processor_channel = processor_channel
        self.num_channels = 1 if processor_channel == "mono" else 2
        self.conv = FIRConvolution(mode="causal", **backend_kwargs)

    def forward(self, input_signals, fir):
        # tanh + normalize_impulse happen inside the convolution while the filter spectra are formed
        if self.processor_channel == "midside":
            return F_.ms_to_lr(F_.fir_filter(F_.lr_to_ms(input_signals), fir))
        return F_.fir_filter(input_signals, fir)

    def parameter_size(self):
        return {"fir": (self.num_channels, self.fir_len)}


class _BiquadStack(nn.Module):
    """Shared plumbing: design -> add the broadcast channel axis -> IIRFilter."""

    def __init__(self, **backend_kwargs):
        super().__init__()
        self.biquad = IIRFilter(order=2, **backend_kwargs)

    def _run(self, input_signals, num, den):
        return self.biquad(input_signals, num.unsqueeze(1), den.unsqueeze(1))

    def folds_source_read(self):
        """render_grafx: the first kernel to read `input_signals` is the biquad cascade (F_.source_fold)."""
        return self.biquad.backend != "fsm"

    def accepts_parameter_repeat(self):
        """render_grafx (4-D sources): un-expanded per-node parameter rows are fine (F_.shared_parameters)."""
        return self.biquad.backend != "fsm"


class BiquadFilter(_BiquadStack):
    """filter.py:87-168."""

    def __init__(self, num_filters=1, normalized=False, **backend_kwargs):
        super().__init__(**backend_kwargs)
        self.num_filters = num_filters
        self.normalized = normalized

    def forward(self, input_signals, Bs, A1_pre, A2_pre, A0=None):
        num, den = F_.biquad_design("stable", Bs, A1_pre, A2_pre, A0 if self.normalized else None,
                                    flags=2 if self.normalized else 0)
        return self._run(input_signals, num, den)

    def parameter_size(self):
        size = {"Bs": (self.num_filters, 3), "A1_pre": self.num_filters, "A2_pre": self.num_filters}
        if self.normalized:
            size["A0"] = self.num_filters
        return size


class PoleZeroFilter(_BiquadStack):
    """filter.py:171-255.  Upstream forgets the channel axis on (Bs, As) and cannot run; the
    documented behaviour (one filter shared by all channels, overall gain exp(log_gain)) is
    implemented here."""

    def __init__(self, num_filters=1, **backend_kwargs):
        super().__init__(**backend_kwargs)
        self.num_filters = num_filters

    def forward(self, input_signals, log_gain, poles, zeros):
        num, den = design.pole_zero(poles, zeros)
        # fold the gain into the first section's numerator: no extra pass over the audio
        gain = torch.exp(log_gain).reshape(-1, 1, 1)
        num = torch.cat([num[:, :1] * gain, num[:, 1:]], 1)
        return self._run(input_signals, num, den)

    def parameter_size(self):
        return {"log_gain": 1, "poles": (self.num_filters, 2), "zeros": (self.num_filters, 2)}


class StateVariableFilter(_BiquadStack):
    """filter.py:258-338."""

    def __init__(self, num_filters=1, **backend_kwargs):
        super().__init__(**backend_kwargs)
        self.num_filters = num_filters

    def forward(self, input_signals, twoR, G, c_hp, c_bp, c_lp):
        return self._run(input_signals, *F_.biquad_design("svf", twoR, G, c_hp, c_bp, c_lp))

    def parameter_size(self):
        return {k: self.num_filters for k in ("twoR", "G", "c_hp", "c_bp", "c_lp")}


class BaseParametricFilter(_BiquadStack):
    """filter.py:341-390: two parameters (cutoff, inverse Q), one biquad."""

    kind = None

    def __init__(self, **backend_kwargs):
        super().__init__(**backend_kwargs)

    def forward(self, input_signals, w0, q_inv):
        return self._run(input_signals, *F_.biquad_design(self.kind, w0, q_inv))

    def parameter_size(self):
        return {"w0": 1, "q_inv": 1}


class LowPassFilter(BaseParametricFilter):
    kind = "lowpass"


class HighPassFilter(BaseParametricFilter):
    kind = "highpass"


class BandPassFilter(BaseParametricFilter):
    kind = "bandpass"


class BandRejectFilter(BaseParametricFilter):
    kind = "bandreject"


class AllPassFilter(BaseParametricFilter):
    kind = "allpass"


class BaseParametricEqualizerFilter(_BiquadStack):
    """filter.py:559-616: three parameters per band (cutoff, inverse Q, log gain), K bands."""

    kind = None

    def __init__(self, num_filters=1, **backend_kwargs):
        super().__init__(**backend_kwargs)
        self.num_filters = num_filters

    def forward(self, input_signals, w0, q_inv, log_gain):
        return self._run(input_signals, *F_.biquad_design(self.kind, w0, q_inv, log_gain))

    def parameter_size(self):
        return {k: self.num_filters for k in ("w0", "q_inv", "log_gain")}


class PeakingFilter(BaseParametricEqualizerFilter):
    kind = "peaking"


class LowShelf(BaseParametricEqualizerFilter):
    kind = "lowshelf"


class HighShelf(BaseParametricEqualizerFilter):
    kind = "highshelf"
