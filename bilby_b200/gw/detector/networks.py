"""InterferometerList and the built-in detectors (bilby/gw/detector/networks.py:13-353, 420-484;
site constants from bilby/gw/detector/detectors/{H1,L1,V1}.interferometer)."""
import numpy as np

from .interferometer import Interferometer
from .psd import PowerSpectralDensity

_SITES = {
    "H1": dict(power_spectral_density=dict(asd_file="aLIGO_O4_high_asd.txt"), minimum_frequency=20,
               maximum_frequency=2048, length=4, latitude=46 + 27. / 60 + 18.528 / 3600,
               longitude=-(119 + 24. / 60 + 27.5657 / 3600), elevation=142.554, xarm_azimuth=125.9994,
               yarm_azimuth=215.9994, xarm_tilt=-6.195e-4, yarm_tilt=1.25e-5),
    "L1": dict(power_spectral_density=dict(asd_file="aLIGO_O4_high_asd.txt"), minimum_frequency=20,
               maximum_frequency=2048, length=4, latitude=30 + 33. / 60 + 46.4196 / 3600,
               longitude=-(90 + 46. / 60 + 27.2654 / 3600), elevation=-6.574, xarm_azimuth=197.7165,
               yarm_azimuth=287.7165, xarm_tilt=-3.121e-4, yarm_tilt=-6.107e-4),
    "V1": dict(power_spectral_density=dict(psd_file="AdV_psd.txt"), minimum_frequency=20,
               maximum_frequency=2048, length=3, latitude=43 + 37. / 60 + 53.0921 / 3600,
               longitude=10 + 30. / 60 + 16.1878 / 3600, elevation=51.884, xarm_azimuth=70.5674,
               yarm_azimuth=160.5674),
}


def get_empty_interferometer(name):
    """networks.py:420-484."""
    if name not in _SITES:
        raise ValueError(f"Interferometer {name} not implemented (available: {sorted(_SITES)})")
    spec = dict(_SITES[name])
    psd = PowerSpectralDensity(**spec.pop("power_spectral_density"))
    return Interferometer(name=name, power_spectral_density=psd, **spec)


class InterferometerList(list):
    def __init__(self, interferometers):
        super().__init__()
        if isinstance(interferometers, str):
            raise TypeError("Input must not be a string")
        for ifo in interferometers:
            if isinstance(ifo, str):
                ifo = get_empty_interferometer(ifo)
            if not isinstance(ifo, Interferometer):
                raise TypeError("Input list of interferometers are not all Interferometer objects")
            self.append(ifo)
        self._check_interferometers()

    def _check_interferometers(self):
        """networks.py:52-79: all detectors must share duration / sampling_frequency / start_time."""
        for attr in ("duration", "sampling_frequency", "start_time"):
            vals = [getattr(ifo.strain_data, attr) for ifo in self]
            if any(v is None for v in vals):
                continue
            if not all(abs(v - vals[0]) < 1e-5 for v in vals):
                raise ValueError(f"The {attr} of all interferometers are not the same")

    def set_strain_data_from_power_spectral_densities(self, sampling_frequency, duration, start_time=0, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        for ifo in self:
            ifo.set_strain_data_from_power_spectral_density(sampling_frequency, duration, start_time, rng=rng)

    def set_strain_data_from_zero_noise(self, sampling_frequency, duration, start_time=0):
        for ifo in self:
            ifo.set_strain_data_from_zero_noise(sampling_frequency, duration, start_time)

    def inject_signal(self, parameters=None, injection_polarizations=None, waveform_generator=None,
                      raise_error=True):
        if injection_polarizations is None:
            if waveform_generator is None:
                raise ValueError("inject_signal needs one of waveform_generator or injection_polarizations.")
            injection_polarizations = waveform_generator.frequency_domain_strain(parameters)
        out = []
        for ifo in self:
            out.append(ifo.inject_signal(parameters, injection_polarizations=injection_polarizations))
        return out

    @property
    def number_of_interferometers(self):
        return len(self)

    @property
    def duration(self):
        return self[0].strain_data.duration

    @property
    def start_time(self):
        return self[0].strain_data.start_time

    @property
    def sampling_frequency(self):
        return self[0].strain_data.sampling_frequency

    @property
    def frequency_array(self):
        return self[0].strain_data.frequency_array

    @property
    def meta_data(self):
        return {ifo.name: ifo.meta_data for ifo in self}
