"""Interferometer with the reference's interface for the hot path
(bilby/gw/detector/interferometer.py: constructor :37-106, antenna_response :267-301,
get_detector_response :303-368, inject_signal_from_waveform_polarizations :492-523,
power_spectral_density_array :551-564, time_delay_from_geocenter :571-590, optimal_snr_squared :607-622,
inner_product :624-640; strain data bilby/gw/detector/strain_data.py:102-159, 212-233).

Set-up (PSD interpolation, masks, noise draws) is host numpy like the reference; every per-parameter
computation (antenna patterns, delays, projection, inner products) is a CUDA kernel behind the C ABI.
"""
import ctypes
import os

import numpy as np

from ...core.utils import create_frequency_series, create_time_series, logger
from .calibration import Recalibrate
from .geometry import InterferometerGeometry
from .psd import PowerSpectralDensity


class _StrainData:
    """The attributes of InterferometerStrainData the likelihood path reads."""

    def __init__(self, minimum_frequency, maximum_frequency, notch_list=None):
        self.__dict__["_version"] = 0
        self.duration = None
        self.sampling_frequency = None
        self.start_time = None
        self._minimum_frequency = minimum_frequency
        self._maximum_frequency = maximum_frequency
        self.notch_list = list(notch_list or [])
        self._frequency_domain_strain = None
        self.window_factor = 1

    def __setattr__(self, name, value):
        """Every change of the data, the band or the notches bumps ``_version``: the device tiles built from this
        object (Interferometer._device, GravitationalWaveTransient.device_network) are rebuilt on the next call."""
        old = self.__dict__.get(name, self)
        changed = old is not value and not (np.isscalar(value) and np.isscalar(old) and old == value)
        object.__setattr__(self, name, value)
        if changed:
            self.__dict__["_version"] += 1

    @property
    def minimum_frequency(self):
        return self._minimum_frequency

    @minimum_frequency.setter
    def minimum_frequency(self, value):
        self._minimum_frequency = value

    @property
    def maximum_frequency(self):
        if self.sampling_frequency is not None:
            self._maximum_frequency = min(self._maximum_frequency, self.sampling_frequency / 2)
        return self._maximum_frequency

    @maximum_frequency.setter
    def maximum_frequency(self, value):
        self._maximum_frequency = value

    @property
    def frequency_array(self):
        return create_frequency_series(self.sampling_frequency, self.duration)

    @property
    def time_array(self):
        return create_time_series(self.sampling_frequency, self.duration, self.start_time)

    @property
    def frequency_mask(self):
        f = self.frequency_array
        mask = (f >= self.minimum_frequency) & (f <= self.maximum_frequency)
        for lo, hi in self.notch_list:
            mask[(f >= lo) & (f <= hi)] = False
        return mask

    @property
    def frequency_domain_strain(self):
        if self._frequency_domain_strain is None:
            raise ValueError("frequency domain strain data not yet set")
        return self._frequency_domain_strain * self.frequency_mask

    @frequency_domain_strain.setter
    def frequency_domain_strain(self, value):
        if len(value) != len(self.frequency_array):
            raise ValueError("The frequency_array and the set strain have different lengths")
        self._frequency_domain_strain = np.asarray(value, dtype=complex)

    def time_within_data(self, time):
        return self.start_time <= time <= self.start_time + self.duration

    def set(self, strain, sampling_frequency, duration, start_time):
        self.sampling_frequency = float(sampling_frequency)
        self.duration = float(duration)
        self.start_time = float(start_time)
        self.frequency_domain_strain = strain


class Interferometer:
    def __init__(self, name, power_spectral_density, minimum_frequency, maximum_frequency, length, latitude,
                 longitude, elevation, xarm_azimuth, yarm_azimuth, xarm_tilt=0., yarm_tilt=0.,
                 calibration_model=None):
        self.__dict__["_own_version"] = 0
        self.name = name
        self.geometry = InterferometerGeometry(length, latitude, longitude, elevation, xarm_azimuth, yarm_azimuth,
                                               xarm_tilt, yarm_tilt)
        self.power_spectral_density = power_spectral_density
        self._calibration_model = Recalibrate() if calibration_model is None else calibration_model
        self.strain_data = _StrainData(minimum_frequency, maximum_frequency)
        self.meta_data = dict(name=name)
        self.reference_time = None
        self._handle = None

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        if name in ("power_spectral_density", "strain_data", "geometry", "reference_time"):
            self.__dict__["_own_version"] += 1

    @property
    def _data_version(self):
        """Identity of everything the device tiles are built from (PSD object, strain, band, notches, calibration
        model): replaced or edited inputs are re-uploaded on the next evaluation."""
        return self._own_version + self.strain_data._version

    @_data_version.setter
    def _data_version(self, value):
        self.__dict__["_own_version"] = value - self.strain_data._version

    def __repr__(self):
        return f"Interferometer(name='{self.name}', minimum_frequency={self.minimum_frequency}, " \
               f"maximum_frequency={self.maximum_frequency})"

    # ---- forwarded strain-data attributes (interferometer.py:22-35 PropertyAccessor list)
    duration = property(lambda self: self.strain_data.duration)
    sampling_frequency = property(lambda self: self.strain_data.sampling_frequency)
    start_time = property(lambda self: self.strain_data.start_time)
    frequency_array = property(lambda self: self.strain_data.frequency_array)
    time_array = property(lambda self: self.strain_data.time_array)
    frequency_mask = property(lambda self: self.strain_data.frequency_mask)
    frequency_domain_strain = property(lambda self: self.strain_data.frequency_domain_strain)

    @property
    def calibration_model(self):
        return self._calibration_model

    @calibration_model.setter
    def calibration_model(self, model):
        self._calibration_model = model
        self._data_version += 1

    @property
    def minimum_frequency(self):
        return self.strain_data.minimum_frequency

    @minimum_frequency.setter
    def minimum_frequency(self, value):
        self.strain_data.minimum_frequency = value
        self._data_version += 1

    @property
    def maximum_frequency(self):
        return self.strain_data.maximum_frequency

    @maximum_frequency.setter
    def maximum_frequency(self, value):
        self.strain_data.maximum_frequency = value
        self._data_version += 1

    @property
    def vertex(self):
        return self.geometry.vertex

    @property
    def detector_tensor(self):
        return self.geometry.detector_tensor

    # ---- data set-up (host)
    def set_strain_data_from_frequency_domain_strain(self, frequency_domain_strain, sampling_frequency=None,
                                                     duration=None, start_time=0, frequency_array=None):
        if frequency_array is not None and (sampling_frequency is None or duration is None):
            df = frequency_array[1] - frequency_array[0]
            duration = 1 / df
            sampling_frequency = 2 * frequency_array[-1]
        self.strain_data.set(np.array(frequency_domain_strain, dtype=complex), sampling_frequency, duration,
                             start_time)
        self._data_version += 1

    def set_strain_data_from_zero_noise(self, sampling_frequency, duration, start_time=0):
        n = len(create_frequency_series(sampling_frequency, duration))
        self.strain_data.set(np.zeros(n, dtype=complex), sampling_frequency, duration, start_time)
        self._data_version += 1

    def set_strain_data_from_power_spectral_density(self, sampling_frequency, duration, start_time=0, rng=None):
        fd, _ = self.power_spectral_density.get_noise_realisation(sampling_frequency, duration, rng=rng)
        self.strain_data.set(fd, sampling_frequency, duration, start_time)
        self._data_version += 1

    @property
    def _window_power_correction(self):
        """interferometer.py:525-536."""
        flag = os.environ.get("BILBY_INCORRECT_PSD_NORMALIZATION", "FALSE").upper()
        return self.strain_data.window_factor if flag in ("TRUE", "1", "YES") else 1

    @property
    def power_spectral_density_array(self):
        return self.power_spectral_density.get_power_spectral_density_array(
            self.strain_data.frequency_array) * self._window_power_correction

    @property
    def amplitude_spectral_density_array(self):
        return self.power_spectral_density_array ** 0.5

    # ---- device-evaluated per-parameter methods
    def _device(self):
        from ..likelihood import DeviceNetwork
        if self._handle is None or self._handle.version != self._data_version:
            self._handle = DeviceNetwork([self])
            self._handle.version = self._data_version
        return self._handle

    def antenna_response(self, ra, dec, time, psi, mode):
        """interferometer.py:267-301 (tensor modes plus/cross on the device; other modes -> 0/1 like the
        reference for names, scalar/vector GR-violating modes are out of scope)."""
        if mode in ("plus", "cross"):
            fp, fc, _ = self._device().antenna(ra, dec, time, psi)[0]
            return fp if mode == "plus" else fc
        if mode in ("x", "y", "breathing", "longitudinal"):
            raise NotImplementedError("non-tensor polarisation modes are outside the hot path")
        return 1 if mode == self.name else 0

    def time_delay_from_geocenter(self, ra, dec, time):
        return self._device().antenna(ra, dec, time, 0.0)[0][2]

    def get_detector_response(self, waveform_polarizations, parameters, frequencies=None):
        """interferometer.py:303-368 for host polarisation arrays (injection / tests)."""
        if frequencies is not None:
            raise NotImplementedError("custom frequency nodes belong to the ROQ / relative-binning kernels")
        return self._device().project(waveform_polarizations, parameters, 0)

    def inner_product(self, signal):
        return self._device().inner_product_arrays(0, signal, None)

    def optimal_snr_squared(self, signal):
        return self._device().inner_product_arrays(0, signal, signal)

    def matched_filter_snr(self, signal):
        return self.inner_product(signal) / self.optimal_snr_squared(signal).real ** 0.5

    def inject_signal_from_waveform_polarizations(self, parameters, injection_polarizations):
        """interferometer.py:492-523."""
        if not self.strain_data.time_within_data(parameters["geocent_time"]):
            logger.warning("Injecting signal outside segment, start_time={}, merger time={}.".format(
                self.strain_data.start_time, parameters["geocent_time"]))
        signal_ifo = self.get_detector_response(injection_polarizations, parameters)
        self.strain_data._frequency_domain_strain = self.strain_data._frequency_domain_strain + signal_ifo
        self._data_version += 1
        self.meta_data["optimal_SNR"] = self.optimal_snr_squared(signal=signal_ifo).real ** 0.5
        self.meta_data["matched_filter_SNR"] = self.matched_filter_snr(signal=signal_ifo)
        self.meta_data["parameters"] = parameters
        return signal_ifo

    def inject_signal(self, parameters, injection_polarizations=None, waveform_generator=None, raise_error=True):
        """interferometer.py:412-490 (without the astropy/lalsimulation duration check)."""
        if injection_polarizations is None:
            if waveform_generator is None:
                raise ValueError("inject_signal needs one of waveform_generator or injection_polarizations.")
            injection_polarizations = waveform_generator.frequency_domain_strain(parameters)
        self.inject_signal_from_waveform_polarizations(parameters, injection_polarizations)
        return injection_polarizations
