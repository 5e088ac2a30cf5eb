"""Edge cases of the batched entry points on every likelihood class (SURVEY.md section 8c asks for the empty and
ragged inputs the domain has): empty batches, a batch of one, batch sizes that are not multiples of the warp / block
/ chunk granularities, waveform-domain errors inside a batch, and tiling invariance (any split of a batch reproduces the
values of the whole batch bit for bit, because every sample is evaluated independently)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import reduced_common as rc  # noqa: E402
import test_gpu_reduced as tgr  # noqa: E402
import test_gpu_multiband as tgm  # noqa: E402
from test_gpu_parity import _build  # noqa: E402


def _likelihoods():
    g, _ = rc.load("relbin_bbh_4s_H1L1V1")
    relbin, d_rb = tgr._relbin_product(g, False)
    g2, _ = rc.load("roq_bbh_4s_H1L1V1")
    roq, d_roq = tgr._roq_product(g2)
    g3, _ = rc.load("multiband_bbh_8s_H1L1V1")
    mb, d_mb = tgm._mb_product(g3, False)
    _, plain, d_pl = _build("noise_H1L1V1")
    drop = lambda d: {k: v for k, v in d.items() if k != "time_jitter"}     # noqa: E731
    return [("plain", plain, drop(d_pl)), ("relbin", relbin, drop(d_rb)), ("roq", roq, drop(d_roq)),
            ("multiband", mb, drop(d_mb))]


def test_empty_single_and_ragged_batches():
    for name, like, draws in _likelihoods():
        n = len(draws["chirp_mass"])
        with np.errstate(invalid="ignore"):
            full = like.log_likelihood_ratio_batch(draws)
        assert full.shape == (n,), name
        empty = like.log_likelihood_ratio_batch({k: v[:0] for k, v in draws.items()})
        assert np.asarray(empty).shape == (0,), name
        for size in (1, 7, 17):
            for start in (0, n - size):
                part = like.log_likelihood_ratio_batch({k: v[start:start + size] for k, v in draws.items()})
                assert np.array_equal(part, full[start:start + size], equal_nan=True), (name, size, start)
        # a tiled batch larger than one block / chunk of every kernel, not a multiple of 16 or 32
        reps = 4099 // n + 1
        big = {k: np.tile(v, reps)[:4099] for k, v in draws.items()}
        out = like.log_likelihood_ratio_batch(big)
        assert np.array_equal(out, np.tile(full, reps)[:4099], equal_nan=True), name


def test_domain_errors_inside_a_batch_keep_their_neighbours():
    """A sample outside the waveform's domain (mass ratio 1:2000) returns the reference's sentinel
    nan_to_num(-inf) (base.py:424-425) and does not disturb the samples around it."""
    for name, like, draws in _likelihoods():
        if name == "roq":
            continue          # ROQ draws carry out-of-window times already (-inf sentinels, test_gpu_reduced.py)
        like.waveform_generator.waveform_arguments["catch_waveform_errors"] = True
        like._net_versions = None          # waveform arguments changed: reconfigure the handle
        d = {k: v[:9].copy() for k, v in draws.items()}
        good = like.log_likelihood_ratio_batch(d)
        d["mass_ratio"][4] = 5e-4
        bad = like.log_likelihood_ratio_batch(d)
        assert bad[4] == np.nan_to_num(-np.inf), name
        keep = np.arange(9) != 4
        assert np.array_equal(bad[keep], good[keep]), name


def test_distance_scaling_is_exact_at_scale():
    """Size-independent property at 50000 draws per class: doubling the luminosity distance halves <d|h> and quarters
    <h|h> EXACTLY (a power-of-two factor commutes with every rounding in the kernels), for the full grid, relative
    binning (edge form), ROQ (row combination + bulk-copy staging) and multi-banding."""
    import torch
    for name, like, draws in _likelihoods():
        n = len(draws["chirp_mass"])
        reps = 50000 // n + 1
        rng = np.random.default_rng(17)
        big = {k: np.tile(v, reps)[:50000].copy() for k, v in draws.items()}
        # decorrelate the copies a little so that the batch is not 50000 / n identical blocks
        big["luminosity_distance"] *= rng.uniform(0.8, 1.25, 50000)
        big["psi"] = big["psi"] + rng.uniform(-0.1, 0.1, 50000)
        s1 = like.inner_products_batch(torch.from_numpy(like.pack(big)).cuda()).cpu().numpy()
        big["luminosity_distance"] *= 2.0
        s2 = like.inner_products_batch(torch.from_numpy(like.pack(big)).cuda()).cpu().numpy()
        fin = np.isfinite(s1[..., 0])
        assert np.array_equal(np.isfinite(s2[..., 0]), fin), name
        assert np.array_equal(s2[..., 0][fin] * 2.0, s1[..., 0][fin]), name
        assert np.array_equal(s2[..., 1] * 2.0, s1[..., 1]), name
        assert np.array_equal(s2[..., 2] * 4.0, s1[..., 2]), name
