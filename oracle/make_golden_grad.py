"""TEST INFRASTRUCTURE -- gradient fixtures for the backward passes, produced by
EXECUTING THE REFERENCE's own code under PyTorch autograd (torchaudio lfilter / torch.fft backward on the CPU).
Run:  python -m oracle.make_golden_grad      Writes tests/golden/grad_*.npz:
x, p_<param>, y, e_w (the loss is sum(w * y)), e_gx (dL/dx), e_g_<param> (dL/dparam), meta."""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle.make_golden import _params, _save
    from oracle.ref_loader import load_reference

    load_reference(even_pad_guard=True)
    import grafx.processors as P

    gen = torch.Generator().manual_seed(7)
    cases = [
        ("grad_peq_lfilter_stereo", "ParametricEqualizer", dict(num_filters=5, processor_channel="stereo", backend="lfilter", flashfftconv=False), 2, 2048),
        ("grad_peq_lfilter_midside", "ParametricEqualizer", dict(num_filters=4, processor_channel="midside", backend="lfilter", flashfftconv=False), 2, 2048),
        ("grad_peq_lfilter_mono", "ParametricEqualizer", dict(num_filters=3, processor_channel="mono", backend="lfilter", flashfftconv=False), 2, 3000),
        ("grad_peq_fsm_stereo", "ParametricEqualizer", dict(num_filters=4, processor_channel="stereo", backend="fsm", fsm_fir_len=1024, flashfftconv=False), 2, 2048),
        ("grad_biquadfilter_lfilter", "BiquadFilter", dict(num_filters=2, backend="lfilter", flashfftconv=False), 3, 2048),
        ("grad_lowpassfilter_lfilter", "LowPassFilter", dict(backend="lfilter", flashfftconv=False), 3, 2048),
        ("grad_statevariablefilter_lfilter", "StateVariableFilter", dict(num_filters=2, backend="lfilter", flashfftconv=False), 2, 2048),
        ("grad_firfilter_stereo", "FIRFilter", dict(fir_len=255, processor_channel="stereo", flashfftconv=False), 2, 2048),
        ("grad_firfilter_midside", "FIRFilter", dict(fir_len=100, processor_channel="midside", flashfftconv=False), 2, 1500),
    ]
    from grafx.processors.core.convolution import FIRConvolution
    from grafx.processors.core.midside import lr_to_ms, ms_to_lr
    from grafx.processors.core.utils import normalize_impulse

    class ComposedFIRFilter:
        """FIRFilter cannot be constructed upstream: the reference's own functions in the order of
        filter.py:65-77, as in oracle/make_golden.py."""

        def __init__(self, fir_len, processor_channel, flashfftconv):
            self.n, self.ch = fir_len, processor_channel
            self.conv = FIRConvolution(mode="causal", flashfftconv=False)

        def parameter_size(self):
            return {"fir": (1 if self.ch == "mono" else 2, self.n)}

        def __call__(self, x, fir):
            f = normalize_impulse(torch.tanh(fir))
            return ms_to_lr(self.conv(lr_to_ms(x), f)) if self.ch == "midside" else self.conv(x, f)

    for name, cls, kw, B, L in cases:
        proc = ComposedFIRFilter(**kw) if cls == "FIRFilter" else getattr(P, cls)(**kw)
        x = torch.randn(B, 2, L, generator=gen).requires_grad_(True)
        prm = {k: v.requires_grad_(True) for k, v in _params(proc.parameter_size(), B, 0.5, gen).items()}
        w = torch.randn(B, 2, L, generator=gen)
        y = proc(x, **prm)
        (y * w).sum().backward()
        extra = {"w": w.numpy(), "gx": x.grad.numpy()}
        for k, v in prm.items():
            extra["g_" + k] = v.grad.numpy()
        _save(name, x.detach(), {k: v.detach() for k, v in prm.items()}, dict(kw, cls=cls), y.detach(), extra=extra)


if __name__ == "__main__":
    main()
