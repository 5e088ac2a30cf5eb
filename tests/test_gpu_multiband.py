"""GPU parity for the multi-banded likelihood through the C ABI (bb_set_multiband ->
K5 edge form) against golden vectors of the UNMODIFIED reference class MBGravitationalWaveTransient
(tests/golden/multiband_*.npz, oracle/tools/make_golden_multiband.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cbc_likelihood as ocl  # noqa: E402

import reduced_common as rc  # noqa: E402
from test_gpu_reduced import _product_ifos, _scale, _priors, RTOL  # noqa: E402


def _mb_product(g, bns, **kw):
    import bilby_b200 as bb
    from bilby_b200.core.prior import Uniform
    from bilby_b200.gw import conversion, source
    inj = rc.injection_of(g)
    approx = str(g["approximant"])
    wa_full = dict(waveform_approximant=approx, reference_frequency=50.0, minimum_frequency=20.0)
    oifos = rc.oracle_ifos(g, inj, ocl.lal_binary_neutron_star if bns else ocl.lal_binary_black_hole, wa_full, lambdas=bns)
    ifos = _product_ifos(oifos)
    model = source.binary_neutron_star_frequency_sequence if bns else source.binary_black_hole_frequency_sequence
    conv = conversion.convert_to_lal_binary_neutron_star_parameters if bns \
        else conversion.convert_to_lal_binary_black_hole_parameters
    wfg = bb.gw.WaveformGenerator(duration=float(g["duration"]), sampling_frequency=float(g["sampling_frequency"]),
                                  start_time=float(g["start_time"]), frequency_domain_source_model=model,
                                  parameter_conversion=conv,
                                  waveform_arguments=dict(waveform_approximant=approx, reference_frequency=50.0))
    tmin, tmax = (float(x) for x in g["geocent_time_prior"])
    priors = kw.pop("priors", None) or _priors()
    priors["geocent_time"] = Uniform(tmin, tmax, "geocent_time")
    like = bb.gw.likelihood.MBGravitationalWaveTransient(ifos, wfg, reference_chirp_mass=float(g["reference_chirp_mass"]),
                                                         priors=priors, **kw)
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_")}
    return like, draws


@pytest.mark.parametrize("name,bns", [("multiband_bbh_8s_H1L1V1", False), ("multiband_bns_32s_H1L1V1", True)])
def test_multiband_vs_reference(name, bns):
    from bilby_b200.core.prior import Uniform, PowerLaw
    import torch
    g, _ = rc.load(name)
    like, draws = _mb_product(g, bns)
    # host set-up == reference set-up
    for key in ("durations", "fb_dfb", "Nbs", "Mbs", "Ks_Ke", "banded_frequency_points", "start_end_idxs",
                "unique_to_original_frequencies"):
        assert np.array_equal(np.asarray(getattr(like, key)), g[key]), key
    for ifo in like.interferometers:
        for kind in ("linear_coeffs", "quadratic_coeffs"):
            ref = g[f"{kind}_{ifo.name}"]
            assert np.allclose(getattr(like, kind)[ifo.name], ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max()), kind
    lnl = like.log_likelihood_ratio_batch(draws)
    assert np.max(np.abs(lnl - g["lnl_none"]) / _scale(g, g["lnl_none"])) < RTOL
    snr = like.inner_products_batch(torch.from_numpy(like.pack(draws)).cuda()).cpu().numpy()
    hh = g["optimal_snr_squared"]
    assert np.max(np.abs(snr[..., 0] + 1j * snr[..., 1] - g["d_inner_h"]) / hh) < RTOL
    assert np.max(np.abs(snr[..., 2] - hh) / hh) < RTOL
    one = like.log_likelihood_ratio({k: float(v[3]) for k, v in draws.items()})
    assert one == lnl[3]
    # the source model at the unique banded points, as the reference's waveform generator returns it
    p = {k: float(v[2]) for k, v in draws.items()}
    pols = like.waveform_generator.frequency_domain_strain(p)
    assert len(pols["plus"]) == len(np.unique(g["banded_frequency_points"]))
    like, _ = _mb_product(g, bns, phase_marginalization=True, priors=_priors(phase=Uniform(0, 2 * np.pi, "phase")))
    lnl = like.log_likelihood_ratio_batch(draws)
    assert np.max(np.abs(lnl - g["lnl_phase"]) / _scale(g, g["lnl_phase"])) < RTOL
    dmin, dmax = (float(x) for x in g["distance_prior"])
    like, _ = _mb_product(g, bns, phase_marginalization=True, distance_marginalization=True,
                          priors=_priors(phase=Uniform(0, 2 * np.pi, "phase"),
                                         luminosity_distance=PowerLaw(2, dmin, dmax, "luminosity_distance")))
    lnl = like.log_likelihood_ratio_batch(draws)
    assert np.max(np.abs(lnl - g["lnl_distance_phase"]) / _scale(g, g["lnl_distance_phase"])) < RTOL
    # weights round trip (multiband.py:647-712, dict form)
    like2, _ = _mb_product(g, bns, weights=like.weights, phase_marginalization=True, distance_marginalization=True,
                           priors=_priors(phase=Uniform(0, 2 * np.pi, "phase"),
                                          luminosity_distance=PowerLaw(2, dmin, dmax, "luminosity_distance")))
    assert np.array_equal(like2.log_likelihood_ratio_batch(draws), lnl)


def test_multiband_tracks_full_grid_at_scale():
    """20000 draws near the injection: the multi-banded likelihood follows the full-grid one (K1) to the level the
    reference's own accuracy test asks for (test/gw/likelihood_test.py MB tests: 1e-3 relative on ln L)."""
    import bilby_b200 as bb
    from bilby_b200.gw import conversion, source
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle", "tools"))
    g, _ = rc.load("multiband_bbh_8s_H1L1V1")
    like, _ = _mb_product(g, False)
    inj = rc.injection_of(g)
    rng = np.random.default_rng(3)
    n = 20000
    m1, m2 = inj["mass_1"], inj["mass_2"]
    mc = (m1 * m2) ** 0.6 / (m1 + m2) ** 0.2
    d = dict(chirp_mass=mc * (1 + rng.uniform(-2e-3, 2e-3, n)), mass_ratio=np.clip(m2 / m1 + rng.uniform(-0.05, 0.05, n), 0.2, 1),
             chi_1=inj["chi_1"] + rng.uniform(-0.02, 0.02, n), chi_2=inj["chi_2"] + rng.uniform(-0.02, 0.02, n),
             luminosity_distance=inj["luminosity_distance"] * rng.uniform(0.7, 1.5, n),
             theta_jn=inj["theta_jn"] + rng.uniform(-0.2, 0.2, n), psi=inj["psi"] + rng.uniform(-0.2, 0.2, n),
             phase=rng.uniform(0, 2 * np.pi, n), ra=inj["ra"] + rng.uniform(-0.05, 0.05, n),
             dec=inj["dec"] + rng.uniform(-0.05, 0.05, n), geocent_time=inj["geocent_time"] + rng.uniform(-2e-3, 2e-3, n))
    lnl_mb = like.log_likelihood_ratio_batch(d)
    wfg = bb.gw.WaveformGenerator(duration=float(g["duration"]), sampling_frequency=float(g["sampling_frequency"]),
                                  start_time=float(g["start_time"]),
                                  frequency_domain_source_model=source.lal_binary_black_hole,
                                  parameter_conversion=conversion.convert_to_lal_binary_black_hole_parameters,
                                  waveform_arguments=dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0,
                                                          minimum_frequency=20.0))
    full = bb.gw.GravitationalWaveTransient(like.interferometers, wfg)
    lnl_full = full.log_likelihood_ratio_batch(d)
    assert np.all(np.isfinite(lnl_mb))
    assert np.max(np.abs(lnl_mb - lnl_full)) < 2e-2


@pytest.mark.parametrize("name,bns", [("multiband_bbh_8s_H1L1V1", False), ("multiband_bns_32s_H1L1V1", True)])
def test_multiband_time_marginalisation_vs_reference(name, bns):
    """multiband.py:714-726, 789-797 on the device: the reference's FFT of the scattered strain * linear_coeffs array as a
    dense contraction over the banded points for the times inside the prior (csrc/bb_reduced.cuh bb_mb_* +
    bb_gemm.cuh), vs lnl_time / lnl_time_phase of the unmodified reference."""
    from bilby_b200.core.prior import Uniform
    g, _ = rc.load(name)
    for key, kw in (("lnl_time", {}), ("lnl_time_phase", dict(phase_marginalization=True))):
        pri = _priors(phase=Uniform(0, 2 * np.pi, "phase")) if kw else _priors()
        like, draws = _mb_product(g, bns, time_marginalization=True, jitter_time=True, priors=pri, **kw)
        assert abs(like._delta_tc - float(g["time_marg_delta_tc"])) < 1e-18
        if kw:
            assert pri["geocent_time"] == float(g["start_time"])           # base.py:187: the sampler's prior is pinned
        d = dict(draws)
        d["geocent_time"] = np.full(len(d["chirp_mass"]), float(g["start_time"]))
        lnl = like.log_likelihood_ratio_batch(d)
        scale = np.maximum(1.0, 0.5 * g["optimal_snr_squared"].sum(axis=1))
        assert np.max(np.abs(lnl - g[key]) / scale) < RTOL, (key, lnl[:4], g[key][:4])
        one = like.log_likelihood_ratio({k: float(v[2]) for k, v in d.items()})
        assert one == lnl[2]


@pytest.mark.parametrize("log2n,batch", [(8, 5), (9, 3), (13, 4), (14, 6), (15, 2), (18, 1)])
def test_batched_fft_vs_numpy(log2n, batch):
    """csrc/bb_fft.cuh (four-step FFT over global memory) vs numpy.fft.fft, the transform of multiband.py:766-797."""
    import torch
    from bilby_b200 import _lib
    h = _lib.Handle()
    rng = np.random.default_rng(log2n)
    n = 1 << log2n
    x = rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))
    xd = torch.from_numpy(x).cuda()
    out = torch.empty_like(xd)
    _lib.check(h.lib.bb_fft_device(h.ptr, xd.data_ptr(), out.data_ptr(), batch, log2n, None))
    ref = np.fft.fft(x, axis=1)
    assert np.max(np.abs(out.cpu().numpy() - ref)) < 1e-13 * np.abs(ref).max() * log2n


@pytest.mark.parametrize("name,bns", [("multiband_bbh_8s_H1L1V1", False), ("multiband_bns_32s_H1L1V1", True)])
def test_multiband_ifft_fft_form_vs_reference(name, bns):
    """linear_interpolation=False (multiband.py:613-646, 766-787): per-detector <h|h> and lnL vs the unmodified
    reference, plus the combination with time marginalisation vs the oracle."""
    import torch
    g, _ = rc.load(name)
    like, draws = _mb_product(g, bns, linear_interpolation=False)
    d = {k: v for k, v in draws.items() if k != "time_jitter"}
    snr = like.inner_products_batch(torch.from_numpy(like.pack(d)).cuda()).cpu().numpy()
    hh = g["optimal_snr_squared_ifft_fft"]
    assert np.max(np.abs(snr[..., 2] - hh) / hh) < RTOL
    assert np.max(np.abs(snr[..., 0] + 1j * snr[..., 1] - g["d_inner_h"]) / hh) < RTOL
    lnl = like.log_likelihood_ratio_batch(d)
    assert np.max(np.abs(lnl - g["lnl_ifft_fft"]) / _scale(g, g["lnl_ifft_fft"])) < RTOL
    # weights dict round trip keeps the form
    like2, _ = _mb_product(g, bns, weights=like.weights)
    assert like2.linear_interpolation is False
    assert np.array_equal(like2.log_likelihood_ratio_batch(d), lnl)
