"""Equalizers -- drop-ins for grafx.processors.eq: ParametricEqualizer (eq.py:217-336), GraphicEqualizer
(eq.py:339-436), ZeroPhaseFIREqualizer / NewZeroPhaseFIREqualizer (eq.py:25-214)."""
from __future__ import annotations

import torch.nn as nn

from .. import functional as F_
from .core.fir import ZeroPhaseFilterBankFIR, ZeroPhaseFIR
from .core.geq import GraphicEqualizerBiquad
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

    def folds_source_read(self):
        """render_grafx: the first kernel to read `input_signals` is the biquad cascade (F_.source_fold)."""
        return self.processor_channel != "midside" and self.biquad.backend != "fsm"

    def accepts_parameter_repeat(self):
        """render_grafx (4-D sources): the per-node parameter rows may arrive un-expanded (F_.shared_parameters): the
        design runs on them as they are and the cascade shares a coefficient row over each run of batch items."""
        return self.biquad.backend != "fsm"

    def parameter_size(self):
        n_channels = 1 if self.processor_channel == "mono" else 2
        return {k: (n_channels, self.num_filters) for k in ("w0", "q_inv", "log_gain")}


class GraphicEqualizer(nn.Module):
    """24-band (Bark) or 31-band (third-octave) graphic equalizer: one peaking biquad per band, all bands in ONE
    launch of the cascade kernel (upstream: K x lfilter or the frequency-sampled FIR).  log_gains: [N, n_ch, K]."""

    def __init__(self, processor_channel="mono", scale="bark", sr=44100, **backend_kwargs):
        super().__init__()
        if processor_channel not in ("mono", "stereo", "midside"):
            raise ValueError(f"Invalid processor_channel: {processor_channel}")
        self.geq = GraphicEqualizerBiquad(scale=scale, sr=sr)
        self.biquad = IIRFilter(**backend_kwargs)
        self.processor_channel = processor_channel

    def forward(self, input_signals, log_gains):
        Bs, As = self.geq(log_gains)
        if self.processor_channel == "midside":
            return F_.ms_to_lr(self.biquad(F_.lr_to_ms(input_signals), Bs, As))
        return self.biquad(input_signals, Bs, As)

    def folds_source_read(self):
        return self.processor_channel != "midside" and self.biquad.backend != "fsm"

    def accepts_parameter_repeat(self):
        return self.biquad.backend != "fsm"

    def parameter_size(self):
        n_channels = 1 if self.processor_channel == "mono" else 2
        return {"log_gains": (n_channels, self.geq.num_bands)}


class ZeroPhaseFIREqualizer(nn.Module):
    """Single-channel zero-phase FIR from K log-magnitudes (hann-windowed, 2K - 1 taps), applied to every channel
    with the zero-phase slice of the convolution (eq.py:25-94)."""

    def __init__(self, num_magnitude_bins=1024):
        super().__init__()
        self.num_magnitude_bins = num_magnitude_bins
        self.fir = ZeroPhaseFIR(num_magnitude_bins)

    def forward(self, input_signals, log_magnitude):
        return F_.fir_conv(input_signals, self.fir(log_magnitude)[:, None, :], "zerophase")

    def parameter_size(self):
        return {"log_magnitude": self.num_magnitude_bins}


class NewZeroPhaseFIREqualizer(nn.Module):
    """Zero-phase FIR equalizer with channel modes and an optional perceptual filterbank parameterisation
    (eq.py:97-214).  `flashfftconv` is accepted and ignored (one convolution engine here)."""

    def __init__(self, num_frequency_bins=1024, processor_channel="mono", use_filterbank=False, filterbank_kwargs={},
                 window="hann", window_kwargs={}, eps=1e-7, flashfftconv=False):
        super().__init__()
        if processor_channel not in ("mono", "stereo", "midside"):
            raise ValueError(f"Invalid processor_channel: {processor_channel}")
        self.num_frequency_bins = num_frequency_bins
        self.processor_channel = processor_channel
        self.use_filterbank = use_filterbank
        self.fir = ZeroPhaseFilterBankFIR(num_frequency_bins=num_frequency_bins, use_filterbank=use_filterbank,
                                          filterbank_kwargs=filterbank_kwargs, window=window,
                                          window_kwargs=window_kwargs, eps=eps)

    def forward(self, input_signals, log_magnitude):
        fir = self.fir(log_magnitude)
        if self.processor_channel == "midside":
            return F_.ms_to_lr(F_.fir_conv(F_.lr_to_ms(input_signals), fir, "zerophase"))
        return F_.fir_conv(input_signals, fir, "zerophase")

    def parameter_size(self):
        n_bins = self.fir.filterbank.num_filters if self.use_filterbank else self.num_frequency_bins
        n_channels = 1 if self.processor_channel == "mono" else 2
        return {"log_magnitude": (n_channels, n_bins)}
