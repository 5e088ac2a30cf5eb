"""Host logic of bilby_b200.gw.multiband (bands, banded points, linear / quadratic coefficients; vectorised numpy)
against the golden vectors of the UNMODIFIED reference class - runs without a GPU (the device upload is skipped by
building the object without its constructor)."""
import types

import numpy as np
import pytest

from oracle import cbc_likelihood as ocl

import reduced_common as rc


def _host_only(g, bns):
    from bilby_b200.core.prior import PriorDict, Uniform
    from bilby_b200.gw.detector import InterferometerList
    from bilby_b200.gw.multiband import MBGravitationalWaveTransient
    inj = rc.injection_of(g)
    wa = dict(waveform_approximant=str(g["approximant"]), reference_frequency=50.0, minimum_frequency=20.0)
    oifos = rc.oracle_ifos(g, inj, ocl.lal_binary_neutron_star if bns else ocl.lal_binary_black_hole, wa, lambdas=bns)
    ifos = InterferometerList([o.name for o in oifos])
    for ifo, o in zip(ifos, oifos):
        ifo.minimum_frequency, ifo.maximum_frequency = 20.0, o.sampling_frequency / 2
        ifo.set_strain_data_from_frequency_domain_strain(o.frequency_domain_strain, sampling_frequency=o.sampling_frequency,
                                                         duration=o.duration, start_time=o.start_time)
    like = MBGravitationalWaveTransient.__new__(MBGravitationalWaveTransient)
    like.interferometers = ifos
    tmin, tmax = (float(x) for x in g["geocent_time_prior"])
    like.priors = PriorDict(dict(geocent_time=Uniform(tmin, tmax, "geocent_time")))
    like.time_reference = "geocent"
    like.waveform_generator = types.SimpleNamespace(waveform_arguments={})
    like.reference_chirp_mass = float(g["reference_chirp_mass"])
    like.highest_mode = 2
    like.linear_interpolation = True
    like.accuracy_factor = 5
    like.time_offset = None
    like.delta_f_end = None
    like.maximum_banding_frequency = None
    like.minimum_banding_duration = 0.
    like.setup_multibanding()
    return like


@pytest.mark.parametrize("name,bns", [("multiband_bbh_8s_H1L1V1", False), ("multiband_bns_32s_H1L1V1", True)])
def test_multiband_setup_vs_reference(name, bns):
    g, _ = rc.load(name)
    like = _host_only(g, bns)
    assert like.time_offset == float(g["time_offset"]) and like.delta_f_end == float(g["delta_f_end"])
    assert like.maximum_banding_frequency == float(g["maximum_banding_frequency"])
    for key in ("durations", "fb_dfb", "Nbs", "Mbs", "Ks_Ke", "banded_frequency_points", "start_end_idxs",
                "unique_to_original_frequencies"):
        assert np.array_equal(np.asarray(getattr(like, key)), g[key]), key
    assert np.array_equal(like.waveform_generator.waveform_arguments["frequencies"], np.unique(g["banded_frequency_points"]))
    for ifo in like.interferometers:
        for kind in ("linear_coeffs", "quadratic_coeffs"):
            ref = g[f"{kind}_{ifo.name}"]
            got = getattr(like, kind)[ifo.name]
            assert np.allclose(got, ref, rtol=1e-10, atol=1e-12 * np.abs(ref).max()), kind
    w = like.weights
    assert set(w) >= {"linear_coeffs", "quadratic_coeffs", "banded_frequency_points", "waveform_frequencies"}


def test_multiband_setting_validation():
    from bilby_b200.gw.multiband import MBGravitationalWaveTransient
    like = MBGravitationalWaveTransient.__new__(MBGravitationalWaveTransient)
    like._prior = None
    with pytest.raises(TypeError):
        like.highest_mode = "2"
    with pytest.raises(TypeError):
        like.linear_interpolation = 1
    with pytest.raises(TypeError):
        like.accuracy_factor = None
    with pytest.raises(TypeError):
        like.reference_chirp_mass = None            # no prior to take the minimum chirp mass from
