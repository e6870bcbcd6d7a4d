from .iir import IIRFilter  # noqa: F401
from .midside import lr_to_ms, ms_to_lr  # noqa: F401
from .convolution import FIRConvolution, convolve  # noqa: F401
from .envelope import Ballistics, TruncatedOnePoleIIRFilter  # noqa: F401
