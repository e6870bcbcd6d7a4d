"""Drop-in mirror of grafx.processors (processors/__init__.py:1-36) for the hot path."""
from . import core  # noqa: F401
from .container import DryWet, GainStagingRegularization, ParallelMix, SerialChain  # noqa: F401
from .dynamics import ApproxCompressor, ApproxNoiseGate, Compressor, NoiseGate  # noqa: F401
from .eq import (  # noqa: F401
    GraphicEqualizer,
    NewZeroPhaseFIREqualizer,
    ParametricEqualizer,
    ZeroPhaseFIREqualizer,
)
from .filter import (  # noqa: F401
    AllPassFilter,
    BandPassFilter,
    BandRejectFilter,
    BiquadFilter,
    FIRFilter,
    HighPassFilter,
    HighShelf,
    LowPassFilter,
    LowShelf,
    PeakingFilter,
    PoleZeroFilter,
    StateVariableFilter,
)
from .delay import MultitapDelay  # noqa: F401
from .reverb import FilteredNoiseShapingReverb, STFTMaskedNoiseReverb  # noqa: F401
from .nonlinear import (  # noqa: F401
    ChebyshevDistortion,
    PiecewiseTanhDistortion,
    PowerDistortion,
    TanhDistortion,
)
from .stereo import (  # noqa: F401
    MidSideToStereo,
    MonoToStereo,
    SideGainImager,
    StereoGain,
    StereoToMidSide,
)
