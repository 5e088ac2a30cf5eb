"""GPU parity of the CubicSpline calibration inside the reduced-order kernels (CAL variants of K5 relative binning,
K5 multi-banding, K6 ROQ and K7 ROQ time marginalisation) against the oracle, which applies
OracleCubicSpline.get_calibration_factor at the bin edges / banded points / ROQ nodes exactly as the reference's
get_detector_response (interferometer.py:364) and ROQ calculate_snrs (roq.py:486-497) do.  The calibration-free
paths of the same kernels are pinned against the reference's golden vectors in test_gpu_reduced.py / test_gpu_multiband.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cbc_likelihood as ocl  # noqa: E402

import reduced_common as rc  # noqa: E402
import test_gpu_reduced as tgr  # noqa: E402
import test_gpu_multiband as tgm  # noqa: E402

N_POINTS = 10
RTOL = 1e-8


def _cal_draws(n, seed=99):
    rng = np.random.default_rng(seed)
    out = {}
    for name in rc.NAMES:
        for i in range(N_POINTS):
            out[f"recalib_{name}_amplitude_{i}"] = rng.normal(0, 0.05, n)
            out[f"recalib_{name}_phase_{i}"] = rng.normal(0, 0.05, n)
    return out


@pytest.fixture
def with_calibration(monkeypatch):
    """Product interferometers get a CubicSpline model, oracle interferometers the restated one."""
    from bilby_b200.gw.detector.calibration import CubicSpline
    plain_product, plain_oracle = tgr._product_ifos, rc.oracle_ifos

    def product(oifos, maximum_frequency=None):
        ifos = plain_product(oifos, maximum_frequency=maximum_frequency)
        for ifo in ifos:
            ifo.calibration_model = CubicSpline(f"recalib_{ifo.name}_", ifo.minimum_frequency, ifo.maximum_frequency,
                                                N_POINTS)
        return ifos

    def oracle(*args, **kw):
        ifos = plain_oracle(*args, **kw)
        for o in ifos:
            o.calibration = ocl.OracleCubicSpline(f"recalib_{o.name}_", o.minimum_frequency, o.maximum_frequency, N_POINTS)
        return ifos
    monkeypatch.setattr(tgr, "_product_ifos", product)
    monkeypatch.setattr(tgm, "_product_ifos", product)
    monkeypatch.setattr(rc, "oracle_ifos", oracle)


class _Golden(dict):
    """The golden file plus recalib_* = 0 parameter columns (the fiducial point of relative binning is built from
    row 0 of the parameter columns and needs the calibration keys, like the reference's fiducial_parameters)."""
    @property
    def files(self):
        return list(self.keys())


def _golden_with_zero_calibration(name):
    g, _ = rc.load(name)
    out = _Golden({k: g[k] for k in g.files})
    n = len(g["param_chirp_mass"])
    for k in _cal_draws(1):
        out["param_" + k] = np.zeros(n)
    return out


def _oracle_lnl(like, draws, rows, **fixed):
    out = []
    for i in rows:
        p = {k: float(v[i]) for k, v in draws.items()}
        p.update(fixed)
        out.append(like.log_likelihood_ratio(p))
    return np.array(out)


def _check(got, ref, scale):
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.max(np.abs(got[fin] - ref[fin]) / scale[fin]) < RTOL


def test_relative_binning_with_calibration(with_calibration):
    # the injection was made without a calibration error: the fiducial point carries recalib_* = 0
    g = _golden_with_zero_calibration("relbin_bbh_4s_H1L1V1")
    n = 8
    like, draws = tgr._relbin_product(g, False)
    draws = {k: v[:n] for k, v in draws.items() if k != "time_jitter"}
    draws.update(_cal_draws(n))
    got = like.log_likelihood_ratio_batch(draws)
    olike, _ = rc.relbin_oracle(g, False)
    ref = _oracle_lnl(olike, draws, range(n))
    scale = np.maximum(np.abs(ref), 0.5 * g["optimal_snr_squared"][:n].sum(axis=1))
    _check(got, ref, scale)
    # a calibration error changes the likelihood (the CAL branch really ran)
    plain = like.log_likelihood_ratio_batch({k: (np.zeros(n) if k.startswith("recalib_") else v) for k, v in draws.items()})
    assert np.max(np.abs(plain - got)) > 1e-3


def test_multiband_with_calibration(with_calibration):
    g, _ = rc.load("multiband_bbh_8s_H1L1V1")
    n = 8
    like, draws = tgm._mb_product(g, False)
    draws = {k: v[:n] for k, v in draws.items()}
    draws.update(_cal_draws(n))
    got = like.log_likelihood_ratio_batch(draws)
    olike, _ = rc.multiband_oracle(g, False)
    ref = _oracle_lnl(olike, draws, range(n))
    scale = np.maximum(np.abs(ref), 0.5 * g["optimal_snr_squared"][:n].sum(axis=1))
    _check(got, ref, scale)


def test_roq_with_calibration(with_calibration):
    from bilby_b200.core.prior import Uniform
    g, _ = rc.load("roq_bbh_4s_H1L1V1")
    like, draws = tgr._roq_product(g)
    n = len(draws["chirp_mass"])
    rows = [0, 1, 2, 5, 9, n - 2, n - 1]             # the last two lie outside the ROQ time window (-inf)
    d = {k: v[rows] for k, v in draws.items() if k != "time_jitter"}
    d.update(_cal_draws(len(rows)))
    got = like.log_likelihood_ratio_batch(d)
    olike, _ = rc.roq_oracle(g)
    with np.errstate(divide="ignore", invalid="ignore"):
        ref = _oracle_lnl(olike, d, range(len(rows)))
    scale = np.maximum(1.0, 0.5 * g["optimal_snr_squared"][rows].sum(axis=1))
    _check(got, ref, scale)
    # time + phase marginalisation: K7 (h_linear with calibration -> ZGEMM -> interpolation -> logsumexp)
    like, _ = tgr._roq_product(g, phase_marginalization=True, time_marginalization=True, jitter_time=True,
                               extra_priors=dict(phase=Uniform(0, 2 * np.pi, "phase")))
    m = 5
    dt = {k: v[:m] for k, v in d.items()}
    dt["geocent_time"] = np.full(m, float(g["time_marg_geocent_time"]))
    dt["time_jitter"] = g["param_time_jitter"][:m]
    got = like.log_likelihood_ratio_batch(dt)
    olike, _ = rc.roq_oracle(g, phase_marginalization=True, time_marginalization=True)
    ref = _oracle_lnl(olike, dt, range(m))
    _check(got, ref, scale[:m])
