"""neural_polar_decoder_b200 -- B200-native (sm_100a) Monte-Carlo decode path of CRISP
(hebbarashwin/neural_polar_decoder) behind the reference's own Python method surface.

Importing the package does not touch CUDA; the first hot-path call loads libnpd.so and requires a
CUDA device (there is no CPU fallback)."""
from . import _lib, rng  # noqa: F401
from .polar import PolarCode  # noqa: F401
from .pac_code import PAC  # noqa: F401
from .utils import errors_ber, errors_bler, errors_bitwise_ber, snr_db2sigma  # noqa: F401

__version__ = "0.1.0"
