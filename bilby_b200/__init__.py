"""bilby_b200 - B200-native (sm_100a) implementation of bilby's compact-binary likelihood hot path.

Python host code mirrors the reference's GravitationalWaveTransient / WaveformGenerator /
Interferometer API for this path only; all arithmetic on the path runs in
hand-written CUDA kernels behind the C ABI of include/bilby_b200.h.  There is no CPU fallback.
"""
from . import core, gw  # noqa: F401

__version__ = "0.1.0"
