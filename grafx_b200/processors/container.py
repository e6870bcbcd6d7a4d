"""Container processors -- drop-ins for grafx.processors.container (container.py:10-148).

SerialChain keeps the reference contract (returns `(output, intermediates)`, takes one kwargs
dict per wrapped processor).  When every wrapped processor is a Compressor / NoiseGate with the
same `iir_len`, the chain is executed by ONE fused kernel launch (the intermediate signal never
goes to HBM); otherwise the processors run one after another.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as tF

from .. import functional as F_
from .dynamics import _DynamicsBase

_MAX_FUSED = 4


class DryWet(nn.Module):
    """y = w * f(u) + (1 - w) * u with w used as given (container.py:45-71)."""

    def __init__(self, processor, external_param=True):
        super().__init__()
        self.processor = processor
        self.external_param = external_param

    def forward(self, input_signals, drywet_weight, **processor_kwargs):
        out = self.processor(input_signals, **processor_kwargs)
        wet, inter = out if isinstance(out, tuple) else (out, None)
        mixed = F_.drywet_mix(input_signals, wet, drywet_weight)
        return (mixed, inter) if isinstance(out, tuple) else mixed

    def parameter_size(self):
        size = self.processor.parameter_size()
        if not self.external_param:
            size["drywet_weight"] = (1,)
        return size


class SerialChain(nn.Module):
    def __init__(self, processors):
        super().__init__()
        self.processors = nn.ModuleDict(processors)

    def _fusable(self):
        procs = list(self.processors.values())
        return (1 < len(procs) <= _MAX_FUSED and all(isinstance(p, _DynamicsBase) for p in procs)
                and len({p.iir_len for p in procs}) == 1)

    def forward(self, input_signals, **processors_kwargs):
        intermediates = {}
        if self._fusable():
            stages = [p.stage(**processors_kwargs[k]) for k, p in self.processors.items()]
            iir_len = next(iter(self.processors.values())).iir_len
            return F_.dynamics_chain(input_signals, stages, iir_len), intermediates
        output_signals = input_signals
        for k, processor in self.processors.items():
            out = processor(output_signals, **processors_kwargs[k])
            if isinstance(out, tuple):
                output_signals, intermediates[k] = out
            else:
                output_signals = out
        return output_signals, intermediates

    def parameter_size(self):
        return {k: v.parameter_size() for k, v in self.processors.items()}


class ParallelMix(nn.Module):
    """Weighted sum of processors fed with the same input (container.py:151-225); weights = softmax or
    softplus / (ln 2 * n) of `parallel_weights` [N, n].  The accumulation y (+)= w_i * out_i is one streaming
    pass per branch (csrc/pointwise.cu)."""

    def __init__(self, processors, activation="softmax"):
        super().__init__()
        self.processors = nn.ModuleDict(processors)
        if activation not in ("softmax", "softplus"):
            raise ValueError(f"Unsupported activation: {activation}")
        self.activation = activation
        self.mult = 1 / (math.log(2) * len(self.processors))

    def forward(self, input_signals, parallel_weights, **processors_kwargs):
        if self.activation == "softmax":
            weights = torch.softmax(parallel_weights, dim=-1)
        else:
            weights = tF.softplus(parallel_weights) * self.mult
        mix, intermediates = None, {}
        for i, (k, processor) in enumerate(self.processors.items()):
            out = processor(input_signals, **processors_kwargs[k])
            if isinstance(out, tuple):
                out, intermediates[k] = out
            w = weights[..., i].reshape(-1, 1)
            if F_._wants_grad(out, w):  # training mode: no in-place accumulation
                mix = F_.pointwise("scale_add", out, w, flags=0 if mix is None else 4, out=mix)
            elif mix is None:
                mix = torch.empty_like(out)
                F_.pointwise("scale_add", out, w, out=mix)
            else:
                F_.pointwise("scale_add", out, w, flags=4, out=mix)
        return mix, intermediates

    def parameter_size(self):
        size = {k: v.parameter_size() for k, v in self.processors.items()}
        size["parallel_weights"] = len(self.processors)
        return size


class GainStagingRegularization(nn.Module):
    """grafx.processors.container.GainStagingRegularization (container.py:231-299): runs the wrapped processor and
    adds `key` -> rms_difference(input, output) (core/utils.py:7-11) to the intermediates.  The two energy
    reductions are one pass each over the audio (gfx_row_mean_square_f32)."""

    def __init__(self, processor, key="gain_reg"):
        super().__init__()
        self.processor = processor
        self.key = key

    def forward(self, input_signals, **processor_kwargs):
        out = self.processor(input_signals, **processor_kwargs)
        if isinstance(out, tuple):
            output_signals, intermediates = out
        else:
            output_signals, intermediates = out, {}
        gain_reg = F_.rms_difference(input_signals, output_signals)
        assert self.key not in intermediates
        intermediates[self.key] = gain_reg
        return output_signals, intermediates

    def parameter_size(self):
        return self.processor.parameter_size()
