"""PowerSpectralDensity (set-up time, host).  Mirrors bilby/gw/detector/psd.py: constructor
(asd_file / psd_file / arrays), linear interpolation with +inf outside the tabulated range (:236-258),
Gaussian noise realisation (:350-376)."""
import os

import numpy as np

from ...core.utils import create_white_noise

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "data", "noise_curves.npz")
_CURVES = None


def _packed_curve(name):
    global _CURVES
    if _CURVES is None:
        _CURVES = dict(np.load(_DATA))
    key = os.path.basename(name)
    if key + ":frequency" not in _CURVES:
        return None
    return _CURVES[key + ":frequency"], _CURVES[key + ":value"]


class PowerSpectralDensity:
    def __init__(self, frequency_array=None, psd_array=None, asd_array=None, psd_file=None, asd_file=None):
        self.psd_file = psd_file
        self.asd_file = asd_file
        if asd_file is not None:
            frequency_array, asd_array = self._read(asd_file)
        elif psd_file is not None:
            frequency_array, psd_array = self._read(psd_file)
        if asd_array is not None:
            psd_array = np.asarray(asd_array) ** 2
        if frequency_array is None or psd_array is None:
            raise ValueError("PowerSpectralDensity needs a file or (frequency_array, psd_array/asd_array)")
        self.frequency_array = np.asarray(frequency_array, dtype=float)
        self.psd_array = np.asarray(psd_array, dtype=float)
        if len(self.frequency_array) != len(self.psd_array):
            raise ValueError("Provided spectral density does not match frequency array.")

    @staticmethod
    def _read(path):
        if os.path.exists(path):
            f, v = np.genfromtxt(path).T
            return f, v
        packed = _packed_curve(path)
        if packed is None:
            raise FileNotFoundError(f"noise curve '{path}' not found (packed curves: aLIGO_O4_high_asd.txt, "
                                    "AdV_psd.txt, aLIGO_ZERO_DET_high_P_{psd,asd}.txt, AdV_asd.txt)")
        return packed

    @property
    def asd_array(self):
        return self.psd_array ** 0.5

    def power_spectral_density_interpolated(self, frequencies):
        """scipy interp1d(kind='linear', bounds_error=False, fill_value=inf) (psd.py:236-247)."""
        frequencies = np.asarray(frequencies, dtype=float)
        out = np.interp(frequencies, self.frequency_array, self.psd_array)
        outside = (frequencies < self.frequency_array[0]) | (frequencies > self.frequency_array[-1])
        # np.interp and scipy's linear interp1d use the same two-point formula; pin exact agreement
        lo = np.clip(np.searchsorted(self.frequency_array, frequencies) - 1, 0, len(self.frequency_array) - 2)
        x0, x1 = self.frequency_array[lo], self.frequency_array[lo + 1]
        y0, y1 = self.psd_array[lo], self.psd_array[lo + 1]
        slope = (y1 - y0) / (x1 - x0)
        out = slope * (frequencies - x0) + y0
        out = np.where(outside, np.inf, out)
        return out

    def get_power_spectral_density_array(self, frequency_array):
        return self.power_spectral_density_interpolated(frequency_array)

    def get_amplitude_spectral_density_array(self, frequency_array):
        return self.power_spectral_density_interpolated(frequency_array) ** 0.5

    def get_noise_realisation(self, sampling_frequency, duration, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        white_noise, frequencies = create_white_noise(sampling_frequency, duration, rng)
        with np.errstate(invalid="ignore"):
            fd = self.power_spectral_density_interpolated(frequencies) ** 0.5 * white_noise
        out_of_bounds = (frequencies < self.frequency_array.min()) | (frequencies > self.frequency_array.max())
        fd[out_of_bounds] = 0j
        return np.nan_to_num(fd), frequencies
