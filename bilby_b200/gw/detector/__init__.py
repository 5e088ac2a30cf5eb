from .psd import PowerSpectralDensity  # noqa: F401
from .geometry import InterferometerGeometry  # noqa: F401
from .interferometer import Interferometer  # noqa: F401
from .networks import InterferometerList, get_empty_interferometer  # noqa: F401
from . import calibration  # noqa: F401
