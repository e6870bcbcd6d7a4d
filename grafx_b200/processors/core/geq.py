"""Graphic-equalizer band design -- stands in for grafx.processors.core.geq.GraphicEqualizerBiquad
(core/geq.py:139-209): one second-order peaking section per band (Liski & Valimaki 2017),

    H_k(z) = (1 + g b - 2 cos(w) z^-1 + (1 - g b) z^-2) / (1 + b - 2 cos(w) z^-1 + (1 - b) z^-2),
    b = tan(B/2) * sqrt((|g~^2 - 1| + 1e-7) / (|g^2 - g~^2| + 1e-7)),  g = exp(log_gain),  g~ = g^0.4,

with b = tan(B/2) alone when |log_gain| < 1e-3.  O(parameters) design math, PyTorch on the device."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

# (centre frequency, bandwidth) in Hz of the two band layouts (core/geq.py:5-136)
_BARK = ((50, 133.3), (150, 160.0), (250, 171.4), (350, 177.8), (450, 214.7), (570, 235.9), (700, 256.7), (840, 294.4),
         (1000, 315.5), (1170, 370.8), (1370, 426.9), (1600, 466.2), (1850, 558.1), (2150, 651.0), (2500, 744.8),
         (2900, 926.5), (3400, 1110.0), (4000, 1467.0), (4800, 1828.0), (5800, 2194.0), (7000, 2735.0), (8500, 3619.0),
         (10500, 5333.0), (13500, 6000.0))
_THIRD_OCTAVE = ((19.69, 9.178), (24.80, 11.56), (31.25, 14.57), (39.37, 18.36), (49.61, 23.13), (62.50, 29.14),
                 (78.75, 36.71), (99.21, 46.25), (125.0, 58.28), (157.5, 73.43), (198.4, 92.51), (250.0, 116.6),
                 (315.0, 146.9), (396.9, 185.0), (500.0, 233.1), (630.0, 293.7), (793.7, 370.0), (1000.0, 466.2),
                 (1260.0, 587.4), (1587.0, 740.1), (2000.0, 932.4), (2520.0, 1175.0), (3175.0, 1480.0), (4000.0, 1865.0),
                 (5040.0, 2350.0), (6350.0, 2846.0), (8000.0, 3502.0), (10080.0, 4253.0), (12700.0, 5038.0),
                 (16000.0, 5689.0), (20160.0, 5573.0))
_NEIGHBOUR_EXPONENT = 0.4


class GraphicEqualizerBiquad(nn.Module):
    def __init__(self, scale="bark", sr=44100):
        super().__init__()
        if scale == "bark":
            bands = _BARK
        elif scale == "third_octave":
            bands = _THIRD_OCTAVE
        else:
            raise ValueError(f"Unsupported scale: {scale}")
        bands = [b for b in bands if b[0] < sr / 2]
        fc = torch.tensor([b[0] for b in bands], dtype=torch.float32)
        fb = torch.tensor([b[1] for b in bands], dtype=torch.float32)
        self.num_bands = len(bands)
        # (upstream builds `fc` from a Python list: an all-integer table gives an int64 buffer, core/geq.py:148-167)
        self.register_buffer("fc", fc.to(torch.int64) if bool((fc == fc.round()).all()) else fc)
        self.register_buffer("fB", fb)
        self.register_buffer("m2_cos_wc", -2 * torch.cos(2 * math.pi * fc / sr))
        self.register_buffer("tan_B_half", torch.tan(math.pi * fb / sr))
        self.register_buffer("c", torch.tensor([0.4] * len(fc)))  # neighbour exponents as upstream registers them (core/geq.py:171)

    def forward(self, log_gains):
        g = torch.exp(log_gains)
        g2 = g.square()
        gn2 = torch.exp(_NEIGHBOUR_EXPONENT * log_gains).square()
        mult = torch.sqrt(((1 - gn2).abs() + 1e-7) / ((g2 - gn2).abs() + 1e-7))
        beta = self.tan_B_half * torch.where(log_gains.abs() >= 1e-3, mult, torch.ones_like(mult))
        gb = g * beta
        mid = self.m2_cos_wc.expand_as(g)
        return torch.stack([1 + gb, mid, 1 - gb], -1), torch.stack([1 + beta, mid, 1 - beta], -1)
