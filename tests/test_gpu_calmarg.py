"""GPU parity of the calibration-marginalised likelihood (bb_set_calibration_marginalization + the ordinary entry
points) against the UNMODIFIED reference's outputs (tests/golden/calmarg_4s_H1L1V1.npz, made by
oracle/tools/make_golden_calmarg.py: base.py:333-346, 860-877, 1037-1051)."""
import os

import numpy as np
import pytest

from test_gpu_parity import _build, _priors, RTOL

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _likelihood(**on):
    g = np.load(os.path.join(GOLDEN, "calmarg_4s_H1L1V1.npz"))
    names = [str(x) for x in g["detectors"]]
    npts = int(g["n_points"])
    table = {}
    for name in names:
        nodes = g[f"curve_nodes_{name}"]
        table[name] = {f"recalib_{name}_{kind}_{i}": nodes[:, k, i]
                       for k, kind in enumerate(("amplitude", "phase")) for i in range(npts)}
    kw = dict(phase_marginalization=on.get("phase", False), distance_marginalization=on.get("luminosity_distance", False),
              time_marginalization=on.get("geocent_time", False), calibration_marginalization=True, calibration_lookup_table=table,
              number_of_response_curves=int(g["n_curves"]), priors=_priors(**on))
    _, like, _ = _build("noise_H1L1V1", **kw)
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_")}
    return g, like, draws


def test_response_curves_match_reference():
    g, like, _ = _likelihood()
    for ifo in like.interferometers:
        got = like.calibration_draws[ifo.name][[0, 17, 39]][:, ::16]
        assert np.allclose(got, g[f"curve_samples_{ifo.name}"], rtol=1e-12, atol=1e-13)
        assert ifo.calibration_model.__class__.__name__ == "Recalibrate"       # calibration.py:552
    assert like._marginalized_parameters == ["recalib_index"]


@pytest.mark.parametrize("mode,on", [("cal", {}), ("cal_phase", dict(phase=True)),
                                     ("cal_distance_phase", dict(phase=True, luminosity_distance=True))])
def test_calibration_marginalised_likelihood_vs_reference(mode, on):
    g, like, draws = _likelihood(**on)
    lnl = like.log_likelihood_ratio_batch(draws)
    ref = g["lnl_" + mode]
    g0 = np.load(os.path.join(GOLDEN, "bbh_4s_noise_H1L1V1.npz"))
    # scale as in test_gpu_parity: max(|lnL|, rho_opt^2 / 2); the optimal SNRs come from this likelihood itself
    snr = like.compute_snrs_batch(draws)
    rho2 = sum(snr[f"{ifo.name}_optimal_snr"] ** 2 for ifo in like.interferometers)
    err = np.abs(lnl - ref) / np.maximum(np.abs(ref), 0.5 * rho2)
    assert err.max() < RTOL, err.max()
    # scalar API = batch of one
    one = like.log_likelihood_ratio({k: float(v[12]) for k, v in draws.items()})
    assert abs(one - lnl[12]) < 1e-9 * max(1.0, abs(lnl[12]))


def test_single_identity_curve_reduces_to_the_plain_likelihood():
    """Property: one response curve equal to 1 everywhere gives back the unmarginalised likelihood."""
    g, like, draws = _likelihood()
    _, plain, _ = _build("noise_H1L1V1")
    n_mask = int(like.interferometers[0].frequency_mask.sum())
    table = {ifo.name: np.ones((1, n_mask), dtype=complex) for ifo in like.interferometers}
    _, one_curve, _ = _build("noise_H1L1V1", calibration_marginalization=True, calibration_lookup_table=table,
                             number_of_response_curves=1, priors=_priors())
    a = one_curve.log_likelihood_ratio_batch(draws)
    b = plain.log_likelihood_ratio_batch(draws)
    assert np.abs(a - b).max() < 1e-9 * np.abs(b).max()


def test_time_plus_calibration_is_refused():
    with pytest.raises(NotImplementedError):
        _likelihood(geocent_time=True)
