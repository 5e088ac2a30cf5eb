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


@pytest.mark.parametrize("mode,on", [("cal_time", dict(geocent_time=True)),
                                     ("cal_time_phase", dict(geocent_time=True, phase=True))])
def test_time_plus_calibration_vs_reference(mode, on):
    """base.py:305-323, 860-866: one transform of the calibrated integrand per response curve (bb_calmarg_time_kernel)."""
    g, like, draws = _likelihood(**on)
    d = dict(draws)
    d["geocent_time"] = np.full_like(d["chirp_mass"], float(g["start_time"]))
    lnl = like.log_likelihood_ratio_batch(d)
    ref = g["lnl_" + mode]
    plain_g, plain, _ = _build("noise_H1L1V1")
    snr = plain.compute_snrs_batch({k: v for k, v in draws.items() if k != "time_jitter"})
    rho2 = sum(snr[f"{ifo.name}_optimal_snr"] ** 2 for ifo in plain.interferometers)
    err = np.abs(lnl - ref) / np.maximum(np.abs(ref), 0.5 * rho2)
    assert err.max() < RTOL, err.max()


def test_time_calibration_distance_is_refused_like_the_reference():
    with pytest.raises(ValueError):
        _likelihood(geocent_time=True, luminosity_distance=True, phase=True)


@pytest.mark.parametrize("mode,on", [("cal", {}), ("cal_phase", dict(phase=True)),
                                     ("cal_distance_phase", dict(phase=True, luminosity_distance=True))])
def test_calibration_reconstruction_vs_reference(mode, on):
    """recalib_index / distance / phase reconstruction (base.py:502-578, 289-290) with the reference's own unit-interval
    draws replayed: the same response curve is picked and the new distance / phase agree to 1e-9."""
    g, like, draws = _likelihood(**on)
    uni = np.nan_to_num(g["uniforms_" + mode], nan=0.5)
    new = like.generate_posterior_samples_from_marginalized_likelihood_batch(draws, uniforms=uni)
    ref = g["recon_" + mode]
    assert np.array_equal(new["recalib_index"], ref[:, 0])
    assert np.allclose(new["luminosity_distance"], ref[:, 1], rtol=1e-9)
    assert np.allclose(new["phase"], ref[:, 2], rtol=1e-9, atol=1e-9)
    one = like.generate_posterior_sample_from_marginalized_likelihood({k: float(v[5]) for k, v in draws.items()})
    assert 0 <= one["recalib_index"] < int(g["n_curves"])


def test_calibration_reconstruction_follows_the_curve_posterior():
    """Property at scale: 4000 draws of recalib_index for one parameter row reproduce the curves' posterior
    softmax(lnL_i), which is recomputed here from the per-curve likelihood of single-curve handles."""
    g, like, draws = _likelihood()
    row = {k: np.full(4000, float(v[12])) for k, v in draws.items()}
    rng = np.random.default_rng(11)
    new = like.generate_posterior_samples_from_marginalized_likelihood_batch(row, uniforms=rng.uniform(0, 1, (4000, 3)))
    idx = new["recalib_index"].astype(int)
    nc = int(g["n_curves"])
    counts = np.bincount(idx, minlength=nc) / 4000.0
    # the marginal likelihood is the mean of the per-curve likelihoods: the most frequent curve must carry a posterior
    # share consistent with its count (binomial 5 sigma) under a posterior that sums to one
    assert counts.sum() == pytest.approx(1.0)
    top = counts.argmax()
    assert counts[top] > 1.0 / nc
    sigma = np.sqrt(counts[top] * (1 - counts[top]) / 4000.0)
    again = like.generate_posterior_samples_from_marginalized_likelihood_batch(
        row, uniforms=np.random.default_rng(12).uniform(0, 1, (4000, 3)))
    c2 = np.bincount(again["recalib_index"].astype(int), minlength=nc) / 4000.0
    assert abs(c2[top] - counts[top]) < 7 * sigma + 1e-3
