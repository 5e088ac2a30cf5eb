"""GPU parity for the reduced-order likelihoods through the C ABI:
relative binning (K5, K5t) and ROQ (K6, K7) against golden vectors of the UNMODIFIED reference classes
(tests/golden/{relbin,roq}_*.npz, oracle/tools/make_golden_reduced.py) and against the oracle at sizes the
golden files do not cover."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cbc_likelihood as ocl  # noqa: E402
from oracle import cbc_reduced as ocr  # noqa: E402

import reduced_common as rc  # noqa: E402

RTOL = 1e-8
T_INJ = ocl.INJECTION["geocent_time"]


def _product_ifos(oifos, maximum_frequency=None):
    from bilby_b200.gw.detector import InterferometerList
    ifos = InterferometerList([o.name for o in oifos])
    for ifo, o in zip(ifos, oifos):
        ifo.minimum_frequency = 20.0
        ifo.maximum_frequency = o.sampling_frequency / 2 if maximum_frequency is None else maximum_frequency
        ifo.set_strain_data_from_frequency_domain_strain(o.frequency_domain_strain, sampling_frequency=o.sampling_frequency,
                                                         duration=o.duration, start_time=o.start_time)
    return ifos


def _relbin_product(g, bns, **kw):
    import bilby_b200 as bb
    from bilby_b200.gw import conversion, source
    inj = rc.injection_of(g)
    approx = str(g["approximant"])
    wa = dict(waveform_approximant=approx, reference_frequency=50.0, minimum_frequency=20.0)
    oifos = rc.oracle_ifos(g, inj, ocl.lal_binary_neutron_star if bns else ocl.lal_binary_black_hole, wa, lambdas=bns)
    oifos = oifos[:kw.pop("n_det", len(oifos))]
    ifos = _product_ifos(oifos)
    model = source.lal_binary_neutron_star_relative_binning if bns else source.lal_binary_black_hole_relative_binning
    conv = conversion.convert_to_lal_binary_neutron_star_parameters if bns \
        else conversion.convert_to_lal_binary_black_hole_parameters
    wfg = bb.gw.WaveformGenerator(duration=float(g["duration"]), sampling_frequency=float(g["sampling_frequency"]),
                                  start_time=float(g["start_time"]), frequency_domain_source_model=model,
                                  parameter_conversion=conv, waveform_arguments=wa)
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_")}
    fid = {k: float(v[0]) for k, v in draws.items() if k != "time_jitter"}
    like = bb.gw.likelihood.RelativeBinningGravitationalWaveTransient(ifos, wfg, fiducial_parameters=fid, **kw)
    return like, draws


def _scale(g, lnl):
    return np.maximum(np.abs(lnl), 0.5 * g["optimal_snr_squared"].sum(axis=1))


def _priors(**kw):
    from bilby_b200.core.prior import PriorDict
    return PriorDict(kw)


@pytest.mark.parametrize("approx", ["IMRPhenomD", "TaylorF2"])
def test_frequency_sequence_waveform_vs_oracle(approx):
    """source.py:1068-1140 semantics on the device: every node evaluated, no f_min / f_max masking."""
    import bilby_b200 as bb
    from bilby_b200.gw import source
    bns = approx == "TaylorF2"
    model = source.binary_neutron_star_roq if bns else source.binary_black_hole_roq
    freqs = np.array([20.0, 23.7, 64.125, 100.0, 250.5, 400.0, 511.9, 900.0])
    p = dict(mass_1=1.5 if bns else 36.0, mass_2=1.3 if bns else 29.0, luminosity_distance=400.0, a_1=0.3, tilt_1=0.0,
             phi_12=0.0, a_2=0.2, tilt_2=np.pi, phi_jl=0.0, theta_jn=0.7, phase=1.1)
    if bns:
        p.update(lambda_1=300.0, lambda_2=500.0)
    got = model(None, **p, waveform_approximant=approx, reference_frequency=20.0,
                frequency_nodes_linear=freqs, frequency_nodes_quadratic=freqs[::2])
    ref = ocr._sequence_polarizations(freqs, p["mass_1"], p["mass_2"], p["luminosity_distance"], p["a_1"], p["tilt_1"],
                                      p["a_2"], p["tilt_2"], p["theta_jn"], p["phase"], p.get("lambda_1", 0.0),
                                      p.get("lambda_2", 0.0), approx, 20.0)
    for mode in ("plus", "cross"):
        s = np.abs(ref[mode]).max()
        assert np.max(np.abs(got["linear"][mode] - ref[mode])) < 1e-10 * s
        assert np.max(np.abs(got["quadratic"][mode] - ref[mode][::2])) < 1e-10 * s


@pytest.mark.parametrize("name,bns", [("relbin_bbh_4s_H1L1V1", False), ("relbin_bns_32s_H1L1V1", True)])
def test_relative_binning_vs_reference(name, bns):
    from bilby_b200.core.prior import Uniform, PowerLaw
    import torch
    g, _ = rc.load(name)
    like, draws = _relbin_product(g, bns)
    assert np.array_equal(like.bin_freqs, g["bin_freqs"])
    assert np.array_equal(like.bin_inds, g["bin_inds"])
    for ifo in like.interferometers:
        ref = g[f"summary_{ifo.name}"]
        assert np.allclose(np.array(like.summary_data[ifo.name]), ref, rtol=1e-8, atol=1e-8 * np.abs(ref).max())
    d = {k: v for k, v in draws.items() if k != "time_jitter"}
    lnl = like.log_likelihood_ratio_batch(d)
    assert np.max(np.abs(lnl - g["lnl_none"]) / _scale(g, g["lnl_none"])) < RTOL
    snr = like.inner_products_batch(torch.from_numpy(like.pack(d)).cuda()).cpu().numpy()
    hh = g["optimal_snr_squared"]
    assert np.max(np.abs(snr[..., 0] + 1j * snr[..., 1] - g["d_inner_h"]) / hh) < RTOL
    assert np.max(np.abs(snr[..., 2] - hh) / hh) < RTOL
    # scalar API == batch of one
    one = like.log_likelihood_ratio({k: float(v[3]) for k, v in d.items()})
    assert one == lnl[3]
    like, _ = _relbin_product(g, bns, phase_marginalization=True, priors=_priors(phase=Uniform(0, 2 * np.pi, "phase")))
    lnl = like.log_likelihood_ratio_batch(d)
    assert np.max(np.abs(lnl - g["lnl_phase"]) / _scale(g, g["lnl_phase"])) < RTOL
    dmin, dmax = (float(x) for x in g["distance_prior"])
    like, _ = _relbin_product(g, bns, phase_marginalization=True, distance_marginalization=True,
                              priors=_priors(phase=Uniform(0, 2 * np.pi, "phase"),
                                             luminosity_distance=PowerLaw(2, dmin, dmax, "luminosity_distance")))
    lnl = like.log_likelihood_ratio_batch(d)
    assert np.max(np.abs(lnl - g["lnl_distance_phase"]) / _scale(g, g["lnl_distance_phase"])) < RTOL


@pytest.mark.parametrize("n_det", [1, 2])
def test_relative_binning_fewer_detectors_vs_oracle(n_det):
    """K5 with one and two detectors (row-blocked edge tables are laid out per detector count; plain reductions instead of
    the nine-sum butterfly) vs the oracle restatement of relative.py:365-430 on the 32 s BNS."""
    g, _ = rc.load("relbin_bns_32s_H1L1V1")
    like, draws = _relbin_product(g, True, n_det=n_det)
    o3, oifos = rc.relbin_oracle(g, True)
    draws0 = {k: float(v[0]) for k, v in draws.items() if k != "time_jitter"}
    o = ocr.OracleRelativeBinning(oifos[:n_det], draws0, source_model=ocr.lal_binary_neutron_star_relative_binning,
                                  waveform_arguments=dict(waveform_approximant=str(g["approximant"]),
                                                          reference_frequency=50.0, minimum_frequency=20.0))
    d = {k: v for k, v in draws.items() if k != "time_jitter"}
    got = like.log_likelihood_ratio_batch(d)
    ref = np.array([o.log_likelihood_ratio({k: float(v[i]) for k, v in d.items()}) for i in range(len(got))])
    scale = np.maximum(np.abs(ref), 0.5 * g["optimal_snr_squared"][:, :n_det].sum(axis=1))
    assert np.max(np.abs(got - ref) / scale) < RTOL


def test_relative_binning_time_marginalised_vs_reference():
    from bilby_b200.core.prior import Uniform
    g, _ = rc.load("relbin_bbh_4s_H1L1V1")
    like, draws = _relbin_product(g, False, phase_marginalization=True, time_marginalization=True, jitter_time=True,
                                  priors=_priors(phase=Uniform(0, 2 * np.pi, "phase"),
                                                 geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time")))
    assert np.array_equal(like.bin_freqs, g["bin_freqs_time"])
    d = dict(draws)
    d["geocent_time"] = np.full_like(d["chirp_mass"], float(g["start_time"]))
    lnl = like.log_likelihood_ratio_batch(d)
    assert np.max(np.abs(lnl - g["lnl_time_phase"]) / _scale(g, g["lnl_time_phase"])) < RTOL


def _roq_product(g, **kw):
    import bilby_b200 as bb
    from bilby_b200.gw import conversion, source
    from bilby_b200.core.prior import Uniform
    inj = dict(ocl.INJECTION)
    fmax = float(g["maximum_frequency"])
    oifos = rc.oracle_ifos(g, inj, ocl.lal_binary_black_hole,
                           dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0, minimum_frequency=20.0),
                           maximum_frequency=fmax)
    oifos = oifos[:kw.pop("n_det", len(oifos))]
    ifos = _product_ifos(oifos, maximum_frequency=fmax)
    for ifo, snr in zip(ifos, g["optimal_snrs"]):
        ifo.meta_data["optimal_SNR"] = float(snr)      # what inject_signal records (interferometer.py:513)
    wfg = bb.gw.WaveformGenerator(
        duration=float(g["duration"]), sampling_frequency=float(g["sampling_frequency"]),
        start_time=float(g["start_time"]), frequency_domain_source_model=source.binary_black_hole_roq,
        parameter_conversion=conversion.convert_to_lal_binary_black_hole_parameters,
        waveform_arguments=dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0,
                                frequency_nodes_linear=g["frequency_nodes_linear"],
                                frequency_nodes_quadratic=g["frequency_nodes_quadratic"]))
    pri = dict(geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time"))
    pri.update(kw.pop("extra_priors", {}))
    like = bb.gw.likelihood.ROQGravitationalWaveTransient(
        ifos, wfg, _priors(**pri), linear_matrix=g["linear_matrix"].astype(complex),
        quadratic_matrix=g["quadratic_matrix"].astype(complex), **kw)
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_")}
    return like, draws


def _close(a, b, scale):
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.array_equal(a[~fin], b[~fin], equal_nan=True)
    assert np.max(np.abs(a[fin] - b[fin]) / scale[fin]) < RTOL


def test_roq_vs_reference():
    import torch
    from bilby_b200.core.prior import Uniform, PowerLaw
    g, _ = rc.load("roq_bbh_4s_H1L1V1")
    like, draws = _roq_product(g)
    assert np.allclose(like.weights["time_samples"], g["time_samples"], rtol=0, atol=1e-12)
    ref0 = g["weights_H1_linear_row0"]
    assert np.allclose(like.weights["H1_linear"][0][0], ref0, rtol=1e-9, atol=1e-9 * np.abs(ref0).max())
    assert np.allclose(like.weights["H1_quadratic"][0], g["weights_H1_quadratic"], rtol=1e-10)
    d = {k: v for k, v in draws.items() if k != "time_jitter"}
    hh = g["optimal_snr_squared"]
    scale = np.maximum(1.0, 0.5 * hh.sum(axis=1))
    lnl = like.log_likelihood_ratio_batch(d)
    _close(lnl, g["lnl_none"], scale)
    assert np.isneginf(lnl[-1]) and np.isneginf(lnl[-2])            # outside the ROQ time window (roq.py:532-533)
    snr = like.inner_products_batch(torch.from_numpy(like.pack(d)).cuda()).cpu().numpy()
    ok = np.isfinite(g["d_inner_h"].real).all(axis=1)
    assert np.max(np.abs(snr[ok, :, 0] + 1j * snr[ok, :, 1] - g["d_inner_h"][ok]) / hh[ok]) < RTOL
    assert np.max(np.abs(snr[..., 2] - hh) / hh) < RTOL
    like, _ = _roq_product(g, phase_marginalization=True, distance_marginalization=True,
                           extra_priors=dict(phase=Uniform(0, 2 * np.pi, "phase"),
                                             luminosity_distance=PowerLaw(2, 100.0, 5000.0, "luminosity_distance")))
    _close(like.log_likelihood_ratio_batch(d), g["lnl_distance_phase"], scale)


@pytest.mark.parametrize("n_det", [1, 2])
def test_roq_fewer_detectors_vs_oracle(n_det):
    """K6 with one and two detectors (lane d works out detector d's window: fewer detectors than three take the plain
    reductions and the last-detector clamp), rows outside the ROQ time window included, vs the oracle (roq.py:467-549)."""
    import torch
    g, _ = rc.load("roq_bbh_4s_H1L1V1")
    like, draws = _roq_product(g, n_det=n_det)
    _, oifos = rc.roq_oracle(g)
    o = ocr.OracleROQ(oifos[:n_det], g["linear_matrix"].astype(complex), g["quadratic_matrix"].astype(complex),
                      g["frequency_nodes_linear"], g["frequency_nodes_quadratic"],
                      time_prior=ocl.OracleUniform(T_INJ - 0.1, T_INJ + 0.1),
                      waveform_arguments=dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0),
                      optimal_snrs=list(g["optimal_snrs"])[:n_det])
    d = {k: v for k, v in draws.items() if k != "time_jitter"}
    got = like.log_likelihood_ratio_batch(d)
    n = len(got)
    ref = np.array([o.log_likelihood_ratio({k: float(v[i]) for k, v in d.items()}) for i in range(n)])
    scale = np.maximum(1.0, 0.5 * g["optimal_snr_squared"][:, :n_det].sum(axis=1))
    _close(got, ref, scale)
    assert np.isneginf(got[-1]) and np.isneginf(got[-2])
    snr = like.inner_products_batch(torch.from_numpy(like.pack(d)).cuda()).cpu().numpy()
    assert snr.shape == (n, n_det, 3)
    hh = g["optimal_snr_squared"][:, :n_det]
    assert np.max(np.abs(snr[..., 2] - hh) / hh) < RTOL


def test_roq_time_marginalised_vs_reference():
    """The dense all-times contraction (ZGEMM) + five-sample interpolation + logsumexp (roq.py:604-651)."""
    from bilby_b200.core.prior import Uniform
    g, _ = rc.load("roq_bbh_4s_H1L1V1")
    like, draws = _roq_product(g, phase_marginalization=True, time_marginalization=True, jitter_time=True,
                               extra_priors=dict(phase=Uniform(0, 2 * np.pi, "phase")))
    assert abs(like._delta_tc - float(g["delta_tc"])) < 1e-15
    d = dict(draws)
    d["geocent_time"] = np.full_like(d["chirp_mass"], float(g["time_marg_geocent_time"]))
    lnl = like.log_likelihood_ratio_batch(d)
    scale = np.maximum(1.0, 0.5 * g["optimal_snr_squared"].sum(axis=1))
    assert np.max(np.abs(lnl - g["lnl_time_phase"]) / scale) < RTOL


def test_relative_binning_tracks_full_likelihood_at_scale():
    """relative_binning_test.py:128-135 at batch scale: 1e5 draws around the fiducial point, relative binning vs
    the full-grid CUDA likelihood (approximation error, not a parity gate) + exact distance scaling."""
    import torch
    import bilby_b200 as bb
    from bilby_b200.gw import conversion, source
    g, _ = rc.load("relbin_bbh_4s_H1L1V1")
    like, draws = _relbin_product(g, False)
    rng = np.random.default_rng(3)
    n = 100_000
    base = {k: float(v[0]) for k, v in draws.items() if k != "time_jitter"}
    d = {k: np.full(n, v) for k, v in base.items()}
    d["chirp_mass"] = base["chirp_mass"] * (1 + rng.uniform(-2e-4, 2e-4, n))
    d["phase"] = rng.uniform(0, 2 * np.pi, n)
    d["luminosity_distance"] = base["luminosity_distance"] * rng.uniform(0.8, 1.3, n)
    d["geocent_time"] = base["geocent_time"] + rng.uniform(-3e-4, 3e-4, n)
    rows = torch.from_numpy(like.pack(d)).cuda()
    rb = like.inner_products_batch(rows).clone()
    wfg = bb.gw.WaveformGenerator(duration=4.0, sampling_frequency=2048.0, start_time=float(g["start_time"]),
                                  frequency_domain_source_model=source.lal_binary_black_hole,
                                  parameter_conversion=conversion.convert_to_lal_binary_black_hole_parameters,
                                  waveform_arguments=dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0,
                                                          minimum_frequency=20.0))
    full_like = bb.gw.GravitationalWaveTransient(like.interferometers, wfg)
    full = full_like.inner_products_batch(rows)
    hh = full[..., 2]
    assert ((rb[..., 2] - hh).abs() / hh).max().item() < 2e-3
    assert ((rb[..., :2] - full[..., :2]).norm(dim=-1) / hh).max().item() < 2e-3
    rows2 = rows.clone()
    rows2[:, 4] *= 2.0
    rb2 = like.inner_products_batch(rows2)
    assert torch.equal(rb2[..., 2] * 4.0, rb[..., 2]) or ((rb2[..., 2] * 4.0 - rb[..., 2]).abs() / rb[..., 2]).max() < 1e-14


def test_roq_multiple_bases_selected_per_sample_vs_oracle(tmp_path):
    """roq.py:368-439: several linear bases with chirp-mass ranges - every sample must be evaluated with the FIRST basis
    whose range contains it; compared with the single-basis oracle of that basis.  Also the weight-file round trip
    with several bases (npz with indexed keys)."""
    import bilby_b200 as bb
    from bilby_b200.core.prior import Uniform
    g, _ = rc.load("roq_bbh_4s_H1L1V1")
    lin, quad = g["linear_matrix"].astype(complex), g["quadratic_matrix"].astype(complex)     # [n_freq, n_basis]
    nodes_l, nodes_q = g["frequency_nodes_linear"], g["frequency_nodes_quadratic"]
    half = lin.shape[1] // 2
    # basis 0: the golden basis (chirp mass 20-30); basis 1: its first half of elements (chirp mass 29-40); overlap at
    # 29-30 -> basis 0 wins there (the reference takes the first match)
    linear = dict(basis_linear={"0": dict(basis=lin.T, frequency_nodes=nodes_l),
                                "1": dict(basis=lin[:, :half].T, frequency_nodes=nodes_l[:half])},
                  prior_range_linear=dict(chirp_mass=np.array([[20.0, 30.0], [29.0, 40.0]])))
    quadratic = dict(basis_quadratic={"0": dict(basis=quad.T, frequency_nodes=nodes_q)})
    like, draws = _roq_product(g, extra_priors=dict(chirp_mass=Uniform(22.0, 38.0, "chirp_mass")))
    base_kw = dict(interferometers=like.interferometers, waveform_generator=like.waveform_generator)
    pri = _priors(geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time"), chirp_mass=Uniform(22.0, 38.0, "chirp_mass"))
    multi = bb.gw.likelihood.ROQGravitationalWaveTransient(priors=pri, linear_matrix=linear, quadratic_matrix=quadratic,
                                                           **base_kw)
    assert multi.number_of_bases_linear == 2 and multi.number_of_bases_quadratic == 1
    d = {k: v[:-2].copy() for k, v in draws.items() if k != "time_jitter"}          # drop the out-of-window rows
    n = len(d["chirp_mass"])
    d["chirp_mass"] = np.linspace(22.5, 37.5, n)
    got = multi.log_likelihood_ratio_batch(d)
    o0, _ = rc.roq_oracle(g)
    o1 = ocr.OracleROQ(o0.ifos, lin[:, :half], quad, nodes_l[:half], nodes_q,
                       time_prior=ocl.OracleUniform(T_INJ - 0.1, T_INJ + 0.1),
                       waveform_arguments=dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0),
                       optimal_snrs=list(g["optimal_snrs"]))
    use1 = d["chirp_mass"] > 30.0
    assert use1.any() and (~use1).any()
    for i in range(n):
        p = {k: float(v[i]) for k, v in d.items()}
        ref = (o1 if use1[i] else o0).log_likelihood_ratio(p)
        assert abs(got[i] - ref) < RTOL * max(1.0, abs(ref), 0.5 * g["optimal_snr_squared"][i].sum()), (i, got[i], ref)
        if i in (0, n - 1):
            assert multi.log_likelihood_ratio(p) == got[i]
            assert multi.basis_number_linear == int(use1[i])
    # weight files with several bases
    path = str(tmp_path / "roq_weights_multi.npz")
    multi.save_weights(path)
    again = bb.gw.likelihood.ROQGravitationalWaveTransient(priors=_priors(
        geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time"), chirp_mass=Uniform(22.0, 38.0, "chirp_mass")),
        weights=path, **base_kw)
    assert again.number_of_bases_linear == 2
    assert np.array_equal(again.log_likelihood_ratio_batch(d), got)
    # a prior inside one basis' range keeps only that basis
    one = bb.gw.likelihood.ROQGravitationalWaveTransient(priors=_priors(
        geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time"), chirp_mass=Uniform(31.0, 38.0, "chirp_mass")),
        weights=path, **base_kw)
    assert one.number_of_bases_linear == 1 and len(one.weights["frequency_nodes_linear"][0]) == half


def test_roq_multibanded_basis_vs_reference():
    """roq.py:920-974, 1006-1053: ROQ weights from a MULTIBANDED basis (three bands of 4 / 2 / 1 s) and the likelihood
    evaluated with them, against the unmodified reference (oracle/tools/make_golden_roq_multiband.py); the multibanded
    ROQ likelihood also tracks the reference's full-grid likelihood of the same draws."""
    import bilby_b200 as bb
    from bilby_b200.gw import conversion, source
    from bilby_b200.core.prior import Uniform
    g, draws = rc.load("roq_multiband_bbh_4s_H1L1V1")
    fmax = float(g["maximum_frequency"])
    oifos = rc.oracle_ifos(g, dict(ocl.INJECTION), ocl.lal_binary_black_hole,
                           dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0, minimum_frequency=20.0),
                           maximum_frequency=fmax)
    ifos = _product_ifos(oifos, maximum_frequency=fmax)
    for ifo, snr in zip(ifos, g["optimal_snrs"]):
        ifo.meta_data["optimal_SNR"] = float(snr)      # what inject_signal records; sets the ROQ time resolution
    wfg = bb.gw.WaveformGenerator(
        duration=float(g["duration"]), sampling_frequency=float(g["sampling_frequency"]),
        start_time=float(g["start_time"]), frequency_domain_source_model=source.binary_black_hole_roq,
        parameter_conversion=conversion.convert_to_lal_binary_black_hole_parameters,
        waveform_arguments=dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0))
    linear = dict(multiband_linear=np.array(True), durations_s_linear=g["durations_s"],
                  start_end_frequency_bins_linear=g["start_end_frequency_bins"],
                  basis_linear={"0": dict(basis=g["basis_linear"].astype(complex),
                                          frequency_nodes=g["frequency_nodes_linear"])})
    quadratic = dict(multiband_quadratic=np.array(True), durations_s_quadratic=g["durations_s"],
                     start_end_frequency_bins_quadratic=g["start_end_frequency_bins"],
                     basis_quadratic={"0": dict(basis=g["basis_quadratic"].astype(complex),
                                                frequency_nodes=g["frequency_nodes_quadratic"])})
    like = bb.gw.likelihood.ROQGravitationalWaveTransient(
        ifos, wfg, _priors(geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time")),
        linear_matrix=linear, quadratic_matrix=quadratic)
    assert np.allclose(like.weights["time_samples"], g["time_samples"], rtol=0, atol=1e-12)
    for name in ("H1", "L1", "V1"):
        ref = g[f"weights_{name}_linear"]
        got = like.weights[f"{name}_linear"][0][::37]
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) < 1e-9 * np.abs(ref).max(), name
        assert np.allclose(like.weights[f"{name}_quadratic"][0], g[f"weights_{name}_quadratic"], rtol=1e-10), name
    import torch
    lnl = like.log_likelihood_ratio_batch(draws)
    hh = like.inner_products_batch(torch.from_numpy(like.pack(draws)).cuda()).cpu().numpy()[..., 2]
    scale = np.maximum(1.0, 0.5 * hh.sum(axis=1))
    _close(lnl, g["lnl_none"], scale)
    assert np.max(np.abs(lnl - g["lnl_full_grid"])) < 5.0            # the approximation itself (20-element basis)
