"""Pins oracle/cbc_multiband.py against golden vectors produced by the UNMODIFIED reference class
MBGravitationalWaveTransient (oracle/tools/make_golden_multiband.py)."""
import numpy as np
import pytest

from oracle import cbc_likelihood as ocl

import reduced_common as rc
from test_oracle_reduced import _eval


@pytest.mark.parametrize("name,bns", [("multiband_bbh_8s_H1L1V1", False), ("multiband_bns_32s_H1L1V1", True)])
def test_multiband_vs_reference(name, bns):
    g, draws = rc.load(name)
    n = 6
    like, ifos = rc.multiband_oracle(g, bns)
    # banding (multiband.py:402-478)
    for key in ("durations", "fb_dfb", "Nbs", "Mbs", "Ks_Ke", "banded_frequency_points", "start_end_idxs",
                "unique_to_original_frequencies"):
        assert np.array_equal(np.asarray(getattr(like, key)), g[key]), key
    assert like.time_offset == float(g["time_offset"]) and like.delta_f_end == float(g["delta_f_end"])
    assert like.maximum_banding_frequency == float(g["maximum_banding_frequency"])
    # coefficients (multiband.py:529-611)
    for ifo in ifos:
        ref = g[f"linear_coeffs_{ifo.name}"]
        assert np.allclose(like.linear_coeffs[ifo.name], ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
        ref = g[f"quadratic_coeffs_{ifo.name}"]
        assert np.allclose(like.quadratic_coeffs[ifo.name], ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    for i in (0, 3):
        p = {k: float(v[i]) for k, v in draws.items()}
        for d, (dh, hh) in enumerate(like.log_likelihood_ratio(p, return_snrs=True)):
            assert abs(dh - g["d_inner_h"][i, d]) < 1e-9 * g["optimal_snr_squared"][i, d]
            assert abs(hh - g["optimal_snr_squared"][i, d]) < 1e-9 * g["optimal_snr_squared"][i, d]
    scale = np.maximum(np.abs(g["lnl_none"][:n]), 0.5 * g["optimal_snr_squared"][:n].sum(axis=1))
    assert np.all(np.abs(_eval(like, draws, n) - g["lnl_none"][:n]) < 1e-9 * scale)
    # multi-banding approximates the full-grid likelihood (reference test: test/gw/likelihood_test.py, 1e-3 level)
    assert np.all(np.abs(g["lnl_none"] - g["lnl_full_grid"]) < 2e-2)
    like, _ = rc.multiband_oracle(g, bns, phase_marginalization=True)
    assert np.all(np.abs(_eval(like, draws, n) - g["lnl_phase"][:n]) < 1e-9 * scale)
    dmin, dmax = g["distance_prior"]
    like, _ = rc.multiband_oracle(g, bns, phase_marginalization=True, distance_marginalization=True,
                                  distance_prior=ocl.OraclePowerLaw(2, float(dmin), float(dmax)),
                                  lookup_table=rc.distance_phase_table())
    assert np.all(np.abs(_eval(like, draws, n) - g["lnl_distance_phase"][:n]) < 1e-9 * scale)


@pytest.mark.parametrize("name,bns", [("multiband_bbh_8s_H1L1V1", False), ("multiband_bns_32s_H1L1V1", True)])
def test_multiband_time_marginalisation_vs_reference(name, bns):
    """multiband.py:714-726, 789-797: FFT of the scattered strain * linear_coeffs array, antenna response at the
    beam-pattern reference time, jitter; golden lnl_time / lnl_time_phase from the unmodified reference."""
    g, draws = rc.load(name)
    n = 4
    tmin, tmax = (float(x) for x in g["geocent_time_prior"])
    d = dict(draws)
    d["geocent_time"] = np.full(len(d["chirp_mass"]), float(g["start_time"]))
    scale = np.maximum(1.0, 0.5 * g["optimal_snr_squared"][:n].sum(axis=1))
    for key, kw in (("lnl_time", {}), ("lnl_time_phase", dict(phase_marginalization=True))):
        like, _ = rc.multiband_oracle(g, bns, time_marginalization=True, jitter_time=True,
                                      time_prior=ocl.OracleUniform(tmin, tmax), **kw)
        assert abs(like._delta_tc - float(g["time_marg_delta_tc"])) < 1e-18
        got = _eval(like, d, n, skip=())
        assert np.all(np.abs(got - g[key][:n]) < 1e-9 * scale), (key, got, g[key][:n])


@pytest.mark.parametrize("name,bns", [("multiband_bbh_8s_H1L1V1", False), ("multiband_bns_32s_H1L1V1", True)])
def test_multiband_ifft_fft_form_vs_reference(name, bns):
    """multiband.py:613-646, 766-787 (linear_interpolation=False)."""
    g, draws = rc.load(name)
    n = 4
    like, _ = rc.multiband_oracle(g, bns, linear_interpolation=False)
    for i in range(n):
        p = {k: float(v[i]) for k, v in draws.items() if k != "time_jitter"}
        for d, (_, hh) in enumerate(like.log_likelihood_ratio(p, return_snrs=True)):
            ref = g["optimal_snr_squared_ifft_fft"][i, d]
            assert abs(hh - ref) < 1e-9 * ref
    scale = np.maximum(np.abs(g["lnl_ifft_fft"][:n]), 0.5 * g["optimal_snr_squared"][:n].sum(axis=1))
    assert np.all(np.abs(_eval(like, draws, n) - g["lnl_ifft_fft"][:n]) < 1e-9 * scale)
