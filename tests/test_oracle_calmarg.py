"""Pins the oracle's restatement of the calibration-marginalised likelihood (base.py:333-346, 860-877;
calibration.py:503-591) against golden vectors from the UNMODIFIED reference (oracle/tools/make_golden_calmarg.py)."""
import os

import numpy as np

from oracle import cbc_likelihood as ocl
from test_oracle_recon import load as load_recon, WA

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def setup():
    g = np.load(os.path.join(GOLDEN, "calmarg_4s_H1L1V1.npz"))
    _, ifos, _ = load_recon()
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_")}
    curves = {}
    for ifo in ifos:
        f = ifo.frequency_array[ifo.frequency_mask]
        curves[ifo.name] = ocl.curves_from_spline_nodes(ifo.name, g[f"curve_nodes_{ifo.name}"], f, int(g["n_points"]))
        assert np.allclose(curves[ifo.name][[0, 17, 39]][:, ::16], g[f"curve_samples_{ifo.name}"], rtol=1e-13, atol=1e-14)
    return g, ifos, draws, curves


def test_calibration_marginalised_likelihood_matches_reference():
    g, ifos, draws, curves = setup()
    n = len(draws["chirp_mass"])
    like = ocl.OracleLikelihood(ifos, waveform_arguments=WA, calibration_draws=curves)
    got = np.array([like.log_likelihood_ratio({k: float(v[i]) for k, v in draws.items()}) for i in range(n)])
    assert np.allclose(got, g["lnl_cal"], rtol=1e-10, atol=1e-10)
    like = ocl.OracleLikelihood(ifos, waveform_arguments=WA, calibration_draws=curves, phase_marginalization=True)
    got = np.array([like.log_likelihood_ratio({k: float(v[i]) for k, v in draws.items()}) for i in range(n)])
    assert np.allclose(got, g["lnl_cal_phase"], rtol=1e-10, atol=1e-10)


def test_calibration_distance_phase_matches_reference():
    g, ifos, draws, curves = setup()
    prior = ocl.OraclePowerLaw(2, 100.0, 5000.0)
    like = ocl.OracleLikelihood(ifos, waveform_arguments=WA, calibration_draws=curves, phase_marginalization=True,
                                distance_marginalization=True, distance_prior=prior,
                                table_processes=min(8, os.cpu_count() or 1))
    idx = [0, 3, 12, 15, 19]
    got = np.array([like.log_likelihood_ratio({k: float(v[i]) for k, v in draws.items()}) for i in idx])
    assert np.allclose(got, g["lnl_cal_distance_phase"][idx], rtol=1e-10, atol=1e-10)


def _recon(like, draws, uni, rows):
    out = []
    for i in rows:
        u = [np.nan, uni[i, 1], uni[i, 2], uni[i, 0]]       # oracle order: time, distance, phase, calibration
        new = like.generate_posterior_sample_from_marginalized_likelihood({k: float(v[i]) for k, v in draws.items()}, u)
        out.append([new["recalib_index"], new["luminosity_distance"], new["phase"]])
    return np.array(out)


def test_calibration_reconstruction_matches_reference():
    """recalib_index / distance / phase reconstruction with calibration marginalisation (base.py:502-578, 289-290)."""
    g, ifos, draws, curves = setup()
    n = len(draws["chirp_mass"])
    like = ocl.OracleLikelihood(ifos, waveform_arguments=WA, calibration_draws=curves)
    got = _recon(like, draws, g["uniforms_cal"], range(n))
    assert np.array_equal(got[:, 0], g["recon_cal"][:, 0])
    assert np.allclose(got[:, 1:], g["recon_cal"][:, 1:], rtol=1e-12)
    like = ocl.OracleLikelihood(ifos, waveform_arguments=WA, calibration_draws=curves, phase_marginalization=True)
    got = _recon(like, draws, g["uniforms_cal_phase"], range(n))
    assert np.array_equal(got[:, 0], g["recon_cal_phase"][:, 0])
    assert np.allclose(got, g["recon_cal_phase"], rtol=1e-9)
    prior = ocl.OraclePowerLaw(2, 100.0, 5000.0)
    like = ocl.OracleLikelihood(ifos, waveform_arguments=WA, calibration_draws=curves, phase_marginalization=True,
                                distance_marginalization=True, distance_prior=prior,
                                table_processes=min(8, os.cpu_count() or 1))
    rows = [0, 3, 12, 15, 19]
    got = _recon(like, draws, g["uniforms_cal_distance_phase"], rows)
    assert np.array_equal(got[:, 0], g["recon_cal_distance_phase"][rows, 0])
    assert np.allclose(got, g["recon_cal_distance_phase"][rows], rtol=1e-9)


def test_time_plus_calibration_matches_reference():
    """base.py:305-323, 860-866: one transform of the calibrated integrand per response curve."""
    g, ifos, draws, curves = setup()
    t_inj = ocl.INJECTION["geocent_time"]
    rows = [0, 3, 12, 15, 19]
    for mode, kw in (("cal_time", {}), ("cal_time_phase", dict(phase_marginalization=True))):
        like = ocl.OracleLikelihood(ifos, waveform_arguments=WA, calibration_draws=curves, time_marginalization=True,
                                    jitter_time=True, time_prior=ocl.OracleUniform(t_inj - 0.1, t_inj + 0.1), **kw)
        got = []
        for i in rows:
            p = {k: float(v[i]) for k, v in draws.items()}
            p["geocent_time"] = float(g["start_time"])
            got.append(like.log_likelihood_ratio(p))
        assert np.allclose(got, g["lnl_" + mode][rows], rtol=1e-10, atol=1e-10), mode
