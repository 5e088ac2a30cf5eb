"""Host-side bookkeeping that needs no GPU: device-state invalidation counters and prior side effects
(bilby/gw/likelihood/base.py:166, 183-229)."""
import numpy as np

from bilby_b200.gw.detector import InterferometerList
from bilby_b200.gw.detector.psd import PowerSpectralDensity


def test_every_input_of_the_device_tiles_bumps_the_data_version():
    ifo = InterferometerList(["H1"])[0]
    ifo.set_strain_data_from_zero_noise(2048.0, 4.0, 0.0)
    _ = ifo.frequency_mask          # first read clamps maximum_frequency to Nyquist (a real change)
    seen = [ifo._data_version]

    def bumped():
        seen.append(ifo._data_version)
        return seen[-1] != seen[-2]

    ifo.power_spectral_density = PowerSpectralDensity(frequency_array=np.array([10.0, 2000.0]),
                                                      psd_array=np.array([1e-46, 1e-46]))
    assert bumped()
    ifo.strain_data.frequency_domain_strain = np.ones(len(ifo.frequency_array), dtype=complex)
    assert bumped()
    ifo.strain_data.minimum_frequency = 30.0
    assert bumped()
    ifo.minimum_frequency = 25.0
    assert bumped()
    ifo.strain_data.notch_list = [(59.0, 61.0)]
    assert bumped()
    assert not ifo.frequency_mask[np.searchsorted(ifo.frequency_array, 60.0)]
    # reading properties (the maximum-frequency clamp writes the same scalar back) must NOT invalidate
    _ = ifo.maximum_frequency, ifo.frequency_mask, ifo.power_spectral_density_array
    assert not bumped()
