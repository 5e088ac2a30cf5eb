"""Pins the oracle's restatement of the marginalised-parameter reconstruction (base.py:502-773) and of the
per-detector SNRs (conversion.py:2215-2288) against golden vectors from the UNMODIFIED reference
(oracle/tools/make_golden_recon.py)."""
import os

import numpy as np
import pytest

from oracle import cbc_likelihood as ocl

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
WA = dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0)
MODES = {
    "phase": dict(phase_marginalization=True),
    "distance": dict(distance_marginalization=True),
    "distance_phase": dict(distance_marginalization=True, phase_marginalization=True),
    "time": dict(time_marginalization=True),
    "time_phase": dict(time_marginalization=True, phase_marginalization=True),
    "time_distance_phase": dict(time_marginalization=True, distance_marginalization=True, phase_marginalization=True),
}


def load():
    g = np.load(os.path.join(GOLDEN, "recon_4s_H1L1V1.npz"))
    names = [str(x) for x in g["detectors"]]
    ifos = [ocl.OracleInterferometer(n, 2048.0, 4.0, float(g["start_time"])) for n in names]
    rng = np.random.default_rng(int(g["noise_seed"]))
    for ifo in ifos:
        ifo.set_gaussian_noise(rng)
    inj = dict(ocl.INJECTION)
    conv = ocl.convert_to_lal_binary_black_hole_parameters(inj)
    pols = ocl.lal_binary_black_hole(ifos[0].frequency_array, *[conv[k] for k in ocl.SOURCE_ARGS], **WA)
    for ifo in ifos:
        ifo.frequency_domain_strain = ifo.frequency_domain_strain + ifo.get_detector_response(pols, conv)
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_")}
    return g, ifos, draws


def make_oracle_likelihood(ifos, mode, lookup_table=None):
    kw = dict(MODES[mode])
    if kw.get("distance_marginalization"):
        kw["distance_prior"] = ocl.OraclePowerLaw(2, 100.0, 5000.0)
        # the distance sample itself never touches the lookup table; only time + distance needs a real one
        kw["lookup_table"] = lookup_table if lookup_table is not None else np.zeros((400, 800))
    if kw.get("time_marginalization"):
        t_inj = ocl.INJECTION["geocent_time"]
        kw["time_prior"] = ocl.OracleUniform(t_inj - 0.1, t_inj + 0.1)
    return ocl.OracleLikelihood(ifos, waveform_arguments=WA, **kw)


def recon(like, draws, uniforms, idx, with_jitter):
    out = []
    for i in idx:
        p = {k: float(v[i]) for k, v in draws.items() if with_jitter or k != "time_jitter"}
        new = like.generate_posterior_sample_from_marginalized_likelihood(p, uniforms[i])
        out.append([new["geocent_time"], new["luminosity_distance"], new["phase"]])
    return np.array(out)


@pytest.mark.parametrize("mode", ["phase", "distance", "distance_phase", "time", "time_phase"])
def test_reconstruction_matches_reference(mode):
    g, ifos, draws = load()
    like = make_oracle_likelihood(ifos, mode)
    idx = list(range(0, 20, 3)) + [12]
    got = recon(like, draws, g["uniforms_" + mode], idx, "time" in mode)
    ref = g["recon_" + mode][idx]
    assert np.allclose(got[:, 0], ref[:, 0], rtol=0, atol=1e-9)          # seconds
    assert np.allclose(got[:, 1:], ref[:, 1:], rtol=1e-9, atol=1e-9)


def test_reconstruction_time_distance_phase():
    g, ifos, draws = load()
    full = ocl.OracleLikelihood(ifos, waveform_arguments=WA, phase_marginalization=True, distance_marginalization=True,
                                distance_prior=ocl.OraclePowerLaw(2, 100.0, 5000.0),
                                table_processes=min(8, os.cpu_count() or 1))
    like = make_oracle_likelihood(ifos, "time_distance_phase", lookup_table=full._dist_margd_loglikelihood_array)
    idx = [0, 5, 12, 13, 17]
    got = recon(like, draws, g["uniforms_time_distance_phase"], idx, True)
    ref = g["recon_time_distance_phase"][idx]
    assert np.allclose(got[:, 0], ref[:, 0], rtol=0, atol=1e-9)
    assert np.allclose(got[:, 1:], ref[:, 1:], rtol=1e-9, atol=1e-9)


def test_per_detector_snrs():
    g, ifos, draws = load()
    like = ocl.OracleLikelihood(ifos, waveform_arguments=WA)
    for i in (0, 7, 12, 19):
        p = {k: float(v[i]) for k, v in draws.items() if k != "time_jitter"}
        per_det = like.log_likelihood_ratio(p, return_snrs=True)
        for d, (dh, hh) in enumerate(per_det):
            assert abs(dh / hh ** 0.5 - g["matched_filter_snr"][i, d]) < 1e-10 * abs(g["matched_filter_snr"][i, d])
            assert abs(hh ** 0.5 - g["optimal_snr"][i, d]) < 1e-10 * g["optimal_snr"][i, d]


def test_reconstruction_8s_normalisation_matches_reference():
    """duration != 4 s: pins that the reference's 16384 Hz transform carries no 4/T factor (base.py:626)."""
    g = np.load(os.path.join(GOLDEN, "recon_8s_zero_H1L1.npz"))
    ifos = [ocl.OracleInterferometer(str(n), 2048.0, 8.0, float(g["start_time"])) for n in g["detectors"]]
    conv = ocl.convert_to_lal_binary_black_hole_parameters(dict(ocl.INJECTION))
    pols = ocl.lal_binary_black_hole(ifos[0].frequency_array, *[conv[k] for k in ocl.SOURCE_ARGS], **WA)
    for ifo in ifos:
        ifo.frequency_domain_strain = ifo.get_detector_response(pols, conv)
    t_inj = ocl.INJECTION["geocent_time"]
    like = ocl.OracleLikelihood(ifos, waveform_arguments=WA, time_marginalization=True, phase_marginalization=True,
                                time_prior=ocl.OracleUniform(t_inj - 0.1, t_inj + 0.1))
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_")}
    got = recon(like, draws, g["uniforms_time_phase"], range(len(g["recon_time_phase"])), True)
    ref = g["recon_time_phase"]
    assert np.allclose(got[:, 0], ref[:, 0], rtol=0, atol=1e-9)
    assert np.allclose(got[:, 2], ref[:, 2], rtol=1e-9, atol=1e-9)
