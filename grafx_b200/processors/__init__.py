"""Drop-in mirror of the public names of `grafx.processors` (processors/__init__.py:1-36) for the hot path.

The export table below maps every class name the reference exposes at package level to the module of this
package that implements it; the names are bound eagerly so that `from grafx_b200.processors import X`, `dir()` and
star-imports behave like the reference package."""
from importlib import import_module

from . import core  # noqa: F401

_EXPORTS = {
    "container": "DryWet GainStagingRegularization ParallelMix SerialChain",
    "delay": "MultitapDelay",
    "dynamics": "ApproxCompressor ApproxNoiseGate Compressor NoiseGate",
    "eq": "GraphicEqualizer NewZeroPhaseFIREqualizer ParametricEqualizer ZeroPhaseFIREqualizer",
    "filter": "AllPassFilter BandPassFilter BandRejectFilter BiquadFilter FIRFilter HighPassFilter HighShelf "
              "LowPassFilter LowShelf PeakingFilter PoleZeroFilter StateVariableFilter",
    "nonlinear": "ChebyshevDistortion PiecewiseTanhDistortion PowerDistortion TanhDistortion",
    "reverb": "FilteredNoiseShapingReverb STFTMaskedNoiseReverb",
    "stereo": "MidSideToStereo MonoToStereo SideGainImager StereoGain StereoToMidSide",
}

__all__ = ["core"]
for _module, _names in _EXPORTS.items():
    _m = import_module(f".{_module}", __name__)
    for _name in _names.split():
        globals()[_name] = getattr(_m, _name)
        __all__.append(_name)
del _module, _names, _m, _name
