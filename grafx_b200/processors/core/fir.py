"""Zero-phase FIR design from a log-magnitude response -- stands in for grafx.processors.core.fir
(core/fir.py:7-123) and the triangular filterbank of core/fft_filterbank.py / core/scale.py.

    h = window * roll(irfft(exp(H_log), n = 2K - 1), K - 1)          (a symmetric, zero-centred response)

With a filterbank the K_fb parameters are log-magnitudes on a perceptual scale:
    |H| = sqrt(M exp(H_fb)^2 + eps),  M = triangular filterbank [K_fb -> K] (synthesis direction).
O(parameters) design math, PyTorch on the device."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

_WINDOWS = {"hann": torch.hann_window, "hamming": torch.hamming_window, "blackman": torch.blackman_window,
            "bartlett": torch.bartlett_window, "kaiser": torch.kaiser_window}


def get_window(window_type, window_length, **kwargs):
    if window_type in ("rectangular", "none", "boxcar", None):
        return None
    if window_type not in _WINDOWS:
        raise ValueError(f"Unsupported window type: {window_type}")
    return _WINDOWS[window_type](window_length, **kwargs)


def log_magnitude_to_zerophase_fir(log_magnitude, fir_len, window=None, magnitude=None):
    lead, k = log_magnitude.shape[:-1], log_magnitude.shape[-1]
    mag = torch.exp(log_magnitude.reshape(-1, k)) if magnitude is None else magnitude
    ir = torch.roll(torch.fft.irfft(mag, n=fir_len), shifts=fir_len // 2, dims=-1)
    if window is not None:
        ir = ir * window[None, :]
    return ir.reshape(*lead, -1)


class ZeroPhaseFIR(nn.Module):
    def __init__(self, num_magnitude_bins=1024, window="hann", **window_kwargs):
        super().__init__()
        self.num_magnitude_bins = num_magnitude_bins
        self.fir_len = 2 * num_magnitude_bins - 1
        w = window if isinstance(window, torch.Tensor) else get_window(window, self.fir_len, **window_kwargs)
        if w is None:
            self.window = None
        else:
            self.register_buffer("window", w)

    def forward(self, log_magnitude):
        return log_magnitude_to_zerophase_fir(log_magnitude, self.fir_len, self.window)


# ---- frequency scales (core/scale.py).  The inverse Traunmuller map applies its low-end correction, and the
# high-end one ONLY when no point needed the low-end one -- as shipped (core/scale.py:40-45): the filterbank is a
# constant of the processor, so the quirk is part of the parity contract.
def _to_scale(f, scale):
    kind, _, variant = scale.partition("_")
    if kind == "bark":
        if variant == "wang":
            return 6.0 * math.asinh(f / 600.0)
        if variant == "schroeder":
            return 7.0 * math.asinh(f / 650.0)
        b = 26.81 * f / (1960.0 + f) - 0.53
        if b < 2:
            b += 0.15 * (2 - b)
        elif b > 20.1:
            b += 0.22 * (b - 20.1)
        return b
    if kind == "mel":
        if variant == "htk":
            return 2595.0 * math.log10(1.0 + f / 700.0)
        mel = f / (200.0 / 3)
        if f >= 1000.0:
            mel = 15.0 + math.log(f / 1000.0) / (math.log(6.4) / 27.0)
        return mel
    if scale == "linear":
        return f
    if scale == "log":
        return math.log(f)
    raise ValueError(f"Unsupported scale: {scale}")


def _from_scale(s, scale):
    kind, _, variant = scale.partition("_")
    if kind == "bark":
        if variant == "wang":
            return 600.0 * torch.sinh(s / 6.0)
        if variant == "schroeder":
            return 650.0 * torch.sinh(s / 7.0)
        s = s.clone()
        if bool((s < 2).any()):
            s = torch.where(s < 2, (s - 0.3) / 0.85, s)
        elif bool((s > 20.1).any()):
            s = torch.where(s > 20.1, (s + 4.422) / 1.22, s)
        return 1960 * ((s + 0.53) / (26.28 - s))
    if kind == "mel":
        if variant == "htk":
            return 700.0 * (10.0 ** (s / 2595.0) - 1.0)
        f = (200.0 / 3) * s
        step = math.log(6.4) / 27.0
        return torch.where(s >= 15.0, 1000.0 * torch.exp(step * (s - 15.0)), f)
    if scale == "linear":
        return s
    if scale == "log":
        return torch.exp(s)
    raise ValueError(f"Unsupported scale: {scale}")


def triangular_filterbank(num_frequency_bins, num_filters=50, scale="bark_traunmuller", f_min=40, f_max=None, sr=44100,
                          low_half_triangle=True):
    """[num_frequency_bins, num_filters] matrix of triangular filters equally spaced on `scale`
    (core/fft_filterbank.py:52-108); with low_half_triangle the first column is what is left below the first peak."""
    if f_max is None or f_max > sr // 2:
        f_max = sr // 2
    n = num_filters - 1 if low_half_triangle else num_filters
    freqs = torch.linspace(0, sr // 2, num_frequency_bins)
    pts = _from_scale(torch.linspace(_to_scale(f_min, scale), _to_scale(f_max, scale), n + 2), scale)
    width = pts[1:] - pts[:-1]
    dist = pts.unsqueeze(0) - freqs.unsqueeze(1)          # [bins, n + 2]
    rising = -dist[:, :-2] / width[:-1]
    falling = dist[:, 2:] / width[1:]
    fb = torch.clamp(torch.minimum(rising, falling), min=0.0)
    if low_half_triangle:
        fb = torch.cat([(1 - fb.sum(-1))[:, None], fb], -1)
    return fb


class TriangularFilterBank(nn.Module):
    def __init__(self, num_frequency_bins, num_filters=50, scale="bark_traunmuller", f_min=40, f_max=None, sr=44100,
                 low_half_triangle=True):
        super().__init__()
        fb = triangular_filterbank(num_frequency_bins, num_filters, scale, f_min, f_max, sr, low_half_triangle)
        self.num_filters = num_filters
        self.register_buffer("filterbank", fb.T.contiguous())                       # synthesis: [filters, bins]
        self.register_buffer("filterbank_normalized", fb / fb.sum(0, keepdim=True))  # analysis: [bins, filters]

    def forward(self, energy, mode="synthesis"):
        return energy @ (self.filterbank if mode == "synthesis" else self.filterbank_normalized)


class ZeroPhaseFilterBankFIR(nn.Module):
    def __init__(self, num_frequency_bins=1024, use_filterbank=False, filterbank_kwargs={}, window="hann",
                 window_kwargs={}, eps=1e-7):
        super().__init__()
        self.num_frequency_bins = num_frequency_bins
        self.fir_len = 2 * num_frequency_bins - 1
        self.eps = eps
        self.use_filterbank = use_filterbank
        if use_filterbank:
            self.filterbank = TriangularFilterBank(num_frequency_bins=num_frequency_bins, **filterbank_kwargs)
        w = window if isinstance(window, torch.Tensor) else get_window(window, self.fir_len, **window_kwargs)
        if w is None:
            self.window = None
        else:
            self.register_buffer("window", w)

    def forward(self, log_magnitude):
        lead, k = log_magnitude.shape[:-1], log_magnitude.shape[-1]
        mag = torch.exp(log_magnitude.reshape(-1, k))
        if self.use_filterbank:
            mag = torch.sqrt(self.filterbank(mag.square()) + self.eps)
        return log_magnitude_to_zerophase_fir(log_magnitude, self.fir_len, self.window, magnitude=mag)
