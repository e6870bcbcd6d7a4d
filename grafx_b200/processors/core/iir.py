"""IIRFilter -- drop-in for grafx.processors.core.iir.IIRFilter (core/iir.py:25-276).

Backends:
  "lfilter", "ssm"  exact recursion.  Both run the same time-parallel CUDA kernel
                    (csrc/biquad.cu): the reference's `ssm` path is a parallel-scan formulation
                    of the same difference equation (and is wrong for K>=2 at the surveyed
                    commit); its own test asserts ssm == lfilter.
  "fsm"             frequency-sampled FIR of length `fsm_fir_len` followed by a causal FIR
                    convolution (core/iir.py:147-152,263-276), i.e. the time-aliased impulse
                    response -- evaluated by csrc/fir.cu.
`flashfftconv` / `fsm_max_input_len` are accepted for signature compatibility and ignored
(there is one convolution engine here).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import functional as F_


class IIRFilter(nn.Module):
    def __init__(self, order=2, backend="fsm", flashfftconv=True, fsm_fir_len=4000,
                 fsm_max_input_len=2**17, fsm_regularization=False):
        super().__init__()
        if order != 2:
            raise ValueError("only second-order sections are supported")
        if backend not in ("fsm", "lfilter", "ssm"):
            raise ValueError(f"Unsupported backend: {backend}")
        if fsm_regularization:
            raise AssertionError("fsm_regularization is not supported (asserts False upstream)")
        self.backend = backend
        self.fsm_fir_len = fsm_fir_len
        if backend == "fsm":
            # same buffer as upstream (core/iir.py:115-117, 269-276) so that its checkpoints load with strict=True; the
            # frequency-sampled design itself evaluates the same phases on the fly (functional.iir_fsm_fir)
            k = torch.arange(fsm_fir_len // 2 + 1)
            phase = torch.arange(order + 1).unsqueeze(-1) * k / fsm_fir_len * 2 * torch.pi
            self.register_buffer("delays", torch.exp(-1j * phase))

    def forward(self, input_signal, Bs, As):
        if self.backend == "fsm":
            return F_.iir_fsm(input_signal, Bs, As, self.fsm_fir_len)
        return F_.biquad_cascade(input_signal, Bs, As)
