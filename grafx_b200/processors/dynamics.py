"""Compressor / NoiseGate -- drop-ins for grafx.processors.dynamics (dynamics.py:213-721).

Same constructor kwargs, forward signature and parameter_size() as the reference; the whole
forward pass (energy, envelope smoother, log, knee, optional gain smoother, gain multiply) is one
fused CUDA kernel (csrc/dynamics.cu).  `flashfftconv` / `max_input_len` are accepted and ignored.
Quirks reproduced as shipped: the relu inside the one-pole smoother also
applies when it smooths a log-gain; ballistics always starts from 1; parameter_size() omits
log_knee for the hard knee.
"""
from __future__ import annotations

import torch.nn as nn

from .. import functional as F_

_SMOOTHERS = ("iir", "ballistics", None)
_KNEES = ("hard", "quadratic", "exponential")


class _DynamicsBase(nn.Module):
    kind = None

    def __init__(self, energy_smoother="iir", gain_smoother=None, gain_smooth_in_log=False, knee="quadratic",
                 iir_len=16384, flashfftconv=True, max_input_len=2**17):
        super().__init__()
        if energy_smoother not in _SMOOTHERS:
            raise ValueError(f"Unknown energy_smoother: {energy_smoother}")
        if gain_smoother not in _SMOOTHERS:
            raise ValueError(f"Unknown gain_smoother: {gain_smoother}")
        if knee not in _KNEES:
            raise ValueError(f"Unknown knee: {knee}")
        self.energy_smoother = energy_smoother
        self.gain_smoother = gain_smoother
        self.gain_smooth_in_log = gain_smooth_in_log
        self.knee = knee
        self.iir_len = iir_len
        # the smoother sub-modules of upstream (dynamics.py:314-340): same names and buffers, so that its checkpoints load
        # with strict=True; the fused kernel does their work
        from .core.envelope import Ballistics, TruncatedOnePoleIIRFilter

        for name, kind in (("energy_smoother_module", energy_smoother), ("gain_smoother_module", gain_smoother)):
            if kind == "iir":
                setattr(self, name, TruncatedOnePoleIIRFilter(iir_len=iir_len))
            elif kind == "ballistics":
                setattr(self, name, Ballistics())

    def stage(self, log_threshold, log_ratio, log_knee=None, z_alpha_pre=None, z_alpha_post=None):
        """Descriptor of this processor for functional.dynamics_chain (also used by SerialChain)."""
        return dict(kind=self.kind, knee=self.knee, energy_smoother=self.energy_smoother,
                    gain_smoother=self.gain_smoother, gain_smooth_in_log=self.gain_smooth_in_log,
                    log_threshold=log_threshold, log_ratio=log_ratio, log_knee=log_knee,
                    z_alpha_pre=z_alpha_pre, z_alpha_post=z_alpha_post)

    def forward(self, input_signals, log_threshold, log_ratio, log_knee=None, z_alpha_pre=None, z_alpha_post=None):
        st = self.stage(log_threshold, log_ratio, log_knee, z_alpha_pre, z_alpha_post)
        return F_.dynamics_chain(input_signals, [st], self.iir_len)

    def accepts_parameter_repeat(self):
        """render_grafx (4-D sources): un-expanded per-node parameter rows are fine (F_.shared_parameters,
        gfx_dynamics_rep_f32)."""
        return True

    def parameter_size(self):
        size = {"log_threshold": 1, "log_ratio": 1}
        if self.knee != "hard":
            size["log_knee"] = 1
        for name, kind in (("z_alpha_pre", self.energy_smoother), ("z_alpha_post", self.gain_smoother)):
            if kind == "iir":
                size[name] = 1
            elif kind == "ballistics":
                size[name] = 2
        return size


class Compressor(_DynamicsBase):
    kind = "compressor"


class NoiseGate(_DynamicsBase):
    kind = "noisegate"


class ApproxCompressor(nn.Module):
    """grafx.processors.dynamics.ApproxCompressor (dynamics.py:8-115): one-pole energy follower, quadratic knee,
    no gain smoother -- the same fused kernel as Compressor(energy_smoother="iir", gain_smoother=None)."""

    def __init__(self, iir_len=16384, flashfftconv=True, max_input_len=2**17):
        super().__init__()
        self.iir_len = iir_len
        self.env_follower = IIREnvelopeFollower(iir_len=iir_len)  # (upstream sub-module, dynamics.py:79: buffers for checkpoints)

    def stage(self, z_alpha, log_threshold, log_ratio, log_knee):
        return dict(kind="compressor", knee="quadratic", energy_smoother="iir", gain_smoother=None,
                    gain_smooth_in_log=False, log_threshold=log_threshold, log_ratio=log_ratio, log_knee=log_knee,
                    z_alpha_pre=z_alpha, z_alpha_post=None)

    def forward(self, input_signals, z_alpha, log_threshold, log_ratio, log_knee=None):
        return F_.dynamics_chain(input_signals, [self.stage(z_alpha, log_threshold, log_ratio, log_knee)], self.iir_len)

    def parameter_size(self):
        return {"z_alpha": 1, "log_threshold": 1, "log_ratio": 1, "log_knee": 1}


class ApproxNoiseGate(nn.Module):
    """grafx.processors.dynamics.ApproxNoiseGate (dynamics.py:118-210), knee as shipped (compute_gain, :186-204):
    ratio = exp(log_ratio) and the knee term divides by 2 (W + 1e-3) -- `knee="approx_gate"` of the fused kernel."""

    def __init__(self, freq_sample_n=16384, flashfftconv=True, max_input_len=2**17):
        super().__init__()
        self.iir_len = freq_sample_n
        self.env_follower = IIREnvelopeFollower(iir_len=freq_sample_n)  # (upstream sub-module, dynamics.py:155)

    def stage(self, z_alpha, log_threshold, log_ratio, log_knee):
        return dict(kind="noisegate", knee="approx_gate", energy_smoother="iir", gain_smoother=None,
                    gain_smooth_in_log=False, log_threshold=log_threshold, log_ratio=log_ratio, log_knee=log_knee,
                    z_alpha_pre=z_alpha, z_alpha_post=None)

    def forward(self, input_signals, z_alpha, log_threshold, log_ratio, log_knee):
        return F_.dynamics_chain(input_signals, [self.stage(z_alpha, log_threshold, log_ratio, log_knee)], self.iir_len)

    def parameter_size(self):
        return {"z_alpha": 1, "log_threshold": 1, "log_ratio": 1, "log_knee": 1}


class BaseEnvelopeFollower(nn.Module):
    """dynamics.py:745-767: detector (mean over channels of x^2 or |x|) -> smoother -> log(envelope + 1e-5), one
    fused kernel launch (gfx_envelope_f32).  `detect_with="rms_channel"` cannot run upstream (it reads an undefined
    `self.eps`) and is rejected here."""

    _smoother_kind = None

    def __init__(self, smoother, detect_with="energy"):
        super().__init__()
        if detect_with not in ("energy", "amplitude"):
            raise ValueError(f"Unknown detect_with: {detect_with}")
        self.detect_with = detect_with
        self.smoother = smoother

    def forward(self, signal, z_alpha):
        return F_.envelope(signal, z_alpha, self._smoother_kind, detect=self.detect_with, log_out=True,
                           iir_len=getattr(self.smoother, "iir_len", 16384))

    def parameter_size(self):
        return self.smoother.parameter_size()


class IIREnvelopeFollower(BaseEnvelopeFollower):
    """dynamics.py:770-784."""

    _smoother_kind = "iir"

    def __init__(self, detect_with="energy", iir_len=16384, flashfftconv=True, max_input_len=2**17):
        from .core.envelope import TruncatedOnePoleIIRFilter

        super().__init__(TruncatedOnePoleIIRFilter(iir_len=iir_len), detect_with=detect_with)


class BallisticsEnvelopeFollower(BaseEnvelopeFollower):
    """dynamics.py:787-790."""

    _smoother_kind = "ballistics"

    def __init__(self, detect_with="energy"):
        from .core.envelope import Ballistics

        super().__init__(Ballistics(), detect_with=detect_with)
