"""FIRConvolution / convolve -- drop-ins for grafx.processors.core.convolution (convolution.py:17-134).

One engine (csrc/fir.cu): `flashfftconv` and `max_input_len` are accepted and ignored.  The
result is the true linear convolution (the reference's documented intent); the reference's
shipped native path deviates from it whenever Lx+Lh-1 is odd."""
from __future__ import annotations

import torch.nn as nn

from ... import functional as F_


def convolve(x, h, mode="zerophase", pad_mode="min"):
    """convolution.py:119-134.  mode: "zerophase" (default upstream) | "causal"."""
    return F_.fir_conv(x, h, mode)


class FIRConvolution(nn.Module):
    def __init__(self, mode="causal", flashfftconv=True, max_input_len=2**17):
        super().__init__()
        if mode not in ("causal", "zerophase"):
            raise ValueError(f"unsupported mode: {mode}")
        self.mode = mode

    def forward(self, input_signals, fir):
        return F_.fir_conv(input_signals, fir, self.mode)
