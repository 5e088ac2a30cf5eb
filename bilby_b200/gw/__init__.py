from . import conversion, source, detector, likelihood, waveform_generator  # noqa: F401
from .waveform_generator import WaveformGenerator  # noqa: F401
from .likelihood import GravitationalWaveTransient  # noqa: F401
