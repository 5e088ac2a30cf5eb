"""Pins oracle/cbc_reduced.py (relative binning, ROQ) against golden vectors produced by the UNMODIFIED reference
classes (oracle/tools/make_golden_reduced.py)."""
import numpy as np
import pytest

from oracle import cbc_likelihood as ocl

import reduced_common as rc


def _eval(like, draws, n, skip=("time_jitter",), **fixed):
    out = []
    for i in range(n):
        p = {k: float(v[i]) for k, v in draws.items() if k not in skip}
        p.update(fixed)
        out.append(like.log_likelihood_ratio(p))
    return np.array(out)


def _close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.array_equal(a[~fin], b[~fin], equal_nan=True)
    assert np.all(np.abs(a[fin] - b[fin]) <= tol * np.maximum(1.0, np.abs(b[fin]))), np.abs(a[fin] - b[fin]).max()


@pytest.mark.parametrize("name,bns", [("relbin_bbh_4s_H1L1V1", False), ("relbin_bns_32s_H1L1V1", True)])
def test_relative_binning_vs_reference(name, bns):
    g, draws = rc.load(name)
    n = 12
    like, ifos = rc.relbin_oracle(g, bns)
    assert np.array_equal(like.bin_freqs, g["bin_freqs"])
    assert np.array_equal(like.bin_inds, g["bin_inds"])
    for ifo in ifos:
        ref = g[f"summary_{ifo.name}"]
        got = np.array(like.summary_data[ifo.name])
        assert np.allclose(got, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    # per-detector inner products before the likelihood's cancellation
    for i in (0, 3):
        p = {k: float(v[i]) for k, v in draws.items() if k != "time_jitter"}
        for d, (dh, hh) in enumerate(like.log_likelihood_ratio(p, return_snrs=True)):
            assert abs(dh - g["d_inner_h"][i, d]) < 1e-9 * g["optimal_snr_squared"][i, d]
            assert abs(hh - g["optimal_snr_squared"][i, d]) < 1e-9 * g["optimal_snr_squared"][i, d]
    scale = np.maximum(np.abs(g["lnl_none"][:n]), 0.5 * g["optimal_snr_squared"][:n].sum(axis=1))
    assert np.all(np.abs(_eval(like, draws, n) - g["lnl_none"][:n]) < 1e-9 * scale)
    like, _ = rc.relbin_oracle(g, bns, phase_marginalization=True)
    assert np.all(np.abs(_eval(like, draws, n) - g["lnl_phase"][:n]) < 1e-9 * scale)
    dmin, dmax = g["distance_prior"]
    like, _ = rc.relbin_oracle(g, bns, phase_marginalization=True, distance_marginalization=True,
                               distance_prior=ocl.OraclePowerLaw(2, float(dmin), float(dmax)),
                               lookup_table=rc.distance_phase_table())
    assert np.all(np.abs(_eval(like, draws, n) - g["lnl_distance_phase"][:n]) < 1e-9 * scale)


def test_relative_binning_time_marginalised_vs_reference():
    g, draws = rc.load("relbin_bbh_4s_H1L1V1")
    n = 8
    t_inj = ocl.INJECTION["geocent_time"]
    like, _ = rc.relbin_oracle(g, False, phase_marginalization=True, time_marginalization=True,
                               time_prior=ocl.OracleUniform(t_inj - 0.1, t_inj + 0.1))
    assert np.array_equal(like.bin_freqs, g["bin_freqs_time"])
    got = _eval(like, draws, n, skip=(), geocent_time=float(g["start_time"]))
    _close(got, g["lnl_time_phase"][:n], 1e-9)


def test_roq_vs_reference():
    g, draws = rc.load("roq_bbh_4s_H1L1V1")
    n = len(draws["chirp_mass"])
    like, ifos = rc.roq_oracle(g)
    assert np.allclose(like.weights["time_samples"], g["time_samples"], rtol=0, atol=1e-12)
    assert np.allclose(like.weights["H1_linear"][0], g["weights_H1_linear_row0"], rtol=1e-9,
                       atol=1e-9 * np.abs(g["weights_H1_linear_row0"]).max())
    assert np.allclose(like.weights["H1_quadratic"], g["weights_H1_quadratic"], rtol=1e-10)
    with np.errstate(divide="ignore", invalid="ignore"):
        got = _eval(like, draws, n)
    _close(got, g["lnl_none"], 1e-8)
    assert np.isneginf(g["lnl_none"][-1]) and np.isneginf(g["lnl_none"][-2])
    like, _ = rc.roq_oracle(g, phase_marginalization=True, distance_marginalization=True,
                            distance_prior=ocl.OraclePowerLaw(2, 100.0, 5000.0), lookup_table=rc.distance_phase_table())
    with np.errstate(divide="ignore", invalid="ignore"):
        got = _eval(like, draws, n)
    _close(got, g["lnl_distance_phase"], 1e-8)
    like, _ = rc.roq_oracle(g, phase_marginalization=True, time_marginalization=True)
    assert abs(like._delta_tc - float(g["delta_tc"])) < 1e-15
    got = _eval(like, draws, 12, skip=(), geocent_time=float(g["time_marg_geocent_time"]))
    _close(got, g["lnl_time_phase"][:12], 1e-8)
