"""MultitapDelay -- drop-in for grafx.processors.delay.MultitapDelay (delay.py:12-177) and the surrogate delay line of
grafx.processors.core.delay (core/delay.py:16-142).

The impulse response is designed in PyTorch on the device (O(taps x segment) work: one surrogate
"delay" per tap from a complex pole, an optional zero-phase colouring FIR per tap, taps of a segment summed,
segments concatenated, unit-energy normalisation); the two O(samples) steps -- the zero-phase colouring convolution
of the tap responses and the causal convolution of the audio -- run on the FIR engine (csrc/fir.cu).  With
`straight_through` (the default) the forward value of every tap is a unit impulse at the arg-max of its soft
response, exactly as upstream.  Returns `(output, {"radii_reg": loss})` like the reference."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F_
from .core.fir import ZeroPhaseFIR


class _UnitModulusGradient(torch.autograd.Function):
    """Identity whose backward rescales every (complex) gradient entry to modulus ~1 -- the
    `normalize_gradients` option of the surrogate delay (core/delay.py:5-13)."""

    @staticmethod
    def forward(ctx, z):
        return z

    @staticmethod
    def backward(ctx, g):
        return g / (g.abs() + 1e-7)


class SurrogateDelay(nn.Module):
    """core/delay.py:16-142: soft delays from complex poles; with `straight_through` the forward value is the hard
    unit impulse while the gradient is the soft response's; `normalize_gradients` as upstream."""

    def __init__(self, N, straight_through=True, radii_loss=True, normalize_gradients=True):
        super().__init__()
        self.straight_through = straight_through
        self.radii_loss = radii_loss
        self.normalize_gradients = normalize_gradients
        self.register_buffer("arange_sin", torch.arange(N // 2 + 1)[None, :])

    def forward(self, z):
        assert z.dtype == torch.cfloat
        shape = z.shape
        z = z.reshape(-1)
        mag = torch.abs(z)
        loss = (1 - torch.tanh(mag)).square().sum()
        if self.normalize_gradients and z.requires_grad and torch.is_grad_enabled():
            z = _UnitModulusGradient.apply(z)
            mag = torch.abs(z)
        z = z * torch.tanh(mag) / (mag + 1e-7)
        irs = torch.fft.irfft((z[:, None] + 1e-7) ** self.arange_sin)
        if self.straight_through:
            with torch.no_grad():
                hard = torch.zeros_like(irs)
                hard[torch.arange(irs.shape[0], device=irs.device), torch.argmax(irs, -1)] = 1
            irs = irs + (hard - irs).detach()
        return irs.reshape(*shape, -1), loss


class MultitapDelay(nn.Module):
    def __init__(self, segment_len=3000, num_segments=20, num_delay_per_segment=1, processor_channel="stereo",
                 zp_filter_per_tap=True, zp_filter_bins=20, flashfftconv=True, max_input_len=2**17, pre_delay=0,
                 **surrogate_delay_kwargs):
        super().__init__()
        if processor_channel not in ("mono", "stereo", "midside"):
            raise ValueError(f"Invalid processor_channel: {processor_channel}")
        self.segment_len = segment_len
        self.num_segments = num_segments
        self.num_delay_per_segment = num_delay_per_segment
        self.zp_filter_per_tap = zp_filter_per_tap
        self.zp_filter_bins = zp_filter_bins
        self.register_buffer("window", torch.hann_window(zp_filter_bins * 2 - 1).view(1, 1, -1))  # (upstream buffer, delay.py:83-85)
        if zp_filter_per_tap:
            self.zp_filter = ZeroPhaseFIR(zp_filter_bins)
        self.delay = SurrogateDelay(N=segment_len, **surrogate_delay_kwargs)
        self.pre_delay = pre_delay
        self.processor_channel = processor_channel
        self.num_channels = 1 if processor_channel == "mono" else 2

    def get_ir(self, delay_z, log_fir_magnitude=None):
        irs, radii_loss = self.delay(torch.view_as_complex(delay_z.contiguous()))
        if self.zp_filter_per_tap:
            irs = F_.fir_conv(irs, self.zp_filter(log_fir_magnitude), "zerophase")
        b, _, t = irs.shape
        irs = irs.reshape(b, self.num_channels, self.num_segments, self.num_delay_per_segment, t).sum(-2)
        irs = irs.reshape(b, self.num_channels, self.num_segments * t)
        return F_.normalize_impulse(irs), {"radii_reg": radii_loss}

    def forward(self, input_signals, delay_z, log_fir_magnitude=None):
        ir, radii_loss = self.get_ir(delay_z, log_fir_magnitude)
        # (upstream applies the convolution directly for every channel mode, delay.py:107)
        output_signals = F_.fir_conv(input_signals, ir, "causal")
        if self.pre_delay != 0:
            shifted = torch.zeros_like(output_signals)
            shifted[:, :, self.pre_delay:] = output_signals[:, :, : -self.pre_delay]
            output_signals = shifted
        return output_signals, radii_loss

    def parameter_size(self):
        num_delay = self.num_segments * self.num_delay_per_segment * self.num_channels
        size = {"delay_z": (num_delay, 2)}
        if self.zp_filter_per_tap:
            size["log_fir_magnitude"] = (num_delay, self.zp_filter_bins)
        return size
