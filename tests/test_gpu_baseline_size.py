"""GPU parity at the sizes BASELINE.json states ("Next round" item 1b): every configuration of
bench.py / bench_configs.py is evaluated through the C ABI on fresh seeded prior draws and compared with the oracle
(CPU restatement of the reference path, fanned out over the host cores) on the SAME data:

  configs[0]  1e4 prior draws, BBH 4 s H1+L1 zero-noise injection, no marginalisation
  configs[1]  1e4 draws of the 1e6-draw workload, H1L1V1, distance + phase marginalisation (the bench.py likelihood)
  configs[2]  4096 draws, BBH 8 s H1L1V1, time marginalisation + CubicSpline(10) (the two-kernel K4a || K4b path)
  configs[3]  256 draws, BNS TaylorF2 + tides 128 s @ 4096 Hz H1L1V1 (259585 masked bins per detector)
  configs[4]  relative binning and ROQ (N_l = 256, N_q = 96) for the 128 s BNS
plus a directed K1 test that puts every region boundary of IMRPhenomD (and the band edges) on a row edge, just
beside it, mid-row and outside the band / the frequency shard.

Gate (north star): |d lnL| <= 1e-8 * max(|lnL|, 1/2 sum rho_opt^2), float64.
"""
import numpy as np
import pytest

import baseline_common as bc
from oracle import cbc_likelihood as ocl
from oracle import cbc_reduced as ocr

pytestmark = pytest.mark.gpu

RTOL = 1e-8
T_INJ = 1126259642.413
WA_BBH = dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0)
WA_BNS = dict(waveform_approximant="TaylorF2", reference_frequency=50.0, minimum_frequency=20.0)


def _check(got, ref, hh, what):
    got, ref = np.asarray(got), np.asarray(ref)
    assert np.all(np.isfinite(ref)), what
    err = np.abs(got - ref) / bc.scale_of(ref, hh)
    worst = int(np.argmax(err))
    assert err[worst] < RTOL, (what, worst, got[worst], ref[worst], err[worst])
    return float(err[worst])


def test_configs0_1e4_prior_draws_vs_oracle():
    """BASELINE.json configs[0] as written: fast_tutorial-style zero-noise injection, H1+L1, 1e4 prior draws;
    the draws go in as (chirp_mass, mass_ratio, chi_i, ...) so the host conversion (conversion.py:182-283) is on the
    compared path."""
    import bilby_b200 as bb
    from bilby_b200.gw.detector import InterferometerList
    from bilby_b200.gw.source import lal_binary_black_hole
    from bilby_b200.workloads import INJECTION, draw_bbh_prior
    n = 10000
    inj = dict(INJECTION)
    start = inj["geocent_time"] - 2.0
    wfg = bb.gw.WaveformGenerator(duration=4.0, sampling_frequency=2048.0, start_time=start,
                                  frequency_domain_source_model=lal_binary_black_hole, waveform_arguments=dict(WA_BBH))
    ifos = InterferometerList(["H1", "L1"])
    ifos.set_strain_data_from_zero_noise(2048.0, 4.0, start)
    ifos.inject_signal(parameters=inj, waveform_generator=wfg)
    like = bb.gw.GravitationalWaveTransient(ifos, wfg)
    draws = draw_bbh_prior(n, np.random.default_rng(20261017))
    got = like.log_likelihood_ratio_batch(draws)
    olike = ocl.OracleLikelihood(bc.oracle_ifos_like(ifos), waveform_arguments=dict(WA_BBH))
    ref = bc.oracle_map(olike, draws, n)
    err = _check(got, ref, bc.total_optimal_snr_squared(like, draws), "configs[0]")
    print(f"configs[0]: {n} draws, max scaled |dlnL| = {err:.2e}")


def test_configs1_1e4_draws_distance_phase_vs_oracle():
    """The likelihood bench.py times (configs[1]): Gaussian noise + injection, H1L1V1, distance + phase
    marginalisation; the oracle builds its OWN 400 x 800 lookup table (base.py:994-1018) on the host cores."""
    import os
    import bench as hb
    from bilby_b200.workloads import draw_bbh_prior
    n = 10000
    like = hb.build_likelihood()
    draws = draw_bbh_prior(n, np.random.default_rng(hb.DRAW_SEED))
    got = like.log_likelihood_ratio_batch(draws)
    olike = ocl.OracleLikelihood(bc.oracle_ifos_like(like.interferometers), waveform_arguments=dict(WA_BBH),
                                 phase_marginalization=True, distance_marginalization=True,
                                 distance_prior=ocl.OraclePowerLaw(2, 100.0, 5000.0),
                                 table_processes=min(os.cpu_count() or 1, 32))
    ref = bc.oracle_map(olike, draws, n)
    err = _check(got, ref, bc.total_optimal_snr_squared(like, draws), "configs[1]")
    print(f"configs[1]: {n} draws, max scaled |dlnL| = {err:.2e}")


def test_configs2_4096_draws_time_marginalisation_and_calibration_vs_oracle():
    import bench_configs as cfg
    n = 4096        # >= 2048: the two-kernel K4a || K4b path the bench runs
    like, rows, cal, _, _ = cfg.build("cfg2", n)
    draws = like._bench_draws
    got = like.log_likelihood_ratio_rows_host(rows, cal)
    olike = ocl.OracleLikelihood(bc.oracle_ifos_like(like.interferometers, calibration_points=10),
                                 waveform_arguments=dict(WA_BBH), time_marginalization=True, jitter_time=True,
                                 time_prior=ocl.OracleUniform(T_INJ - 0.1, T_INJ + 0.1))
    ref = bc.oracle_map(olike, draws, n)
    err = _check(got, ref, bc.total_optimal_snr_squared(like, draws, cal), "configs[2]")
    print(f"configs[2]: {n} draws, max scaled |dlnL| = {err:.2e}")


def test_configs3_256_taylorf2_128s_draws_vs_oracle():
    import bench_configs as cfg
    n = 256
    like, rows, _, _, _ = cfg.build("cfg3", n)
    draws = like._bench_draws
    got = like.log_likelihood_ratio_rows_host(rows)
    olike = ocl.OracleLikelihood(bc.oracle_ifos_like(like.interferometers), source_model=ocl.lal_binary_neutron_star,
                                 waveform_arguments=dict(WA_BNS))
    per_det = bc.oracle_map(olike, draws, n, snrs=True)
    ref = np.array([sum(d.real for d, _ in s) - 0.5 * sum(h for _, h in s) for s in per_det])
    hh = np.array([sum(h for _, h in s) for s in per_det])
    err = _check(got, ref, hh, "configs[3]")
    import torch
    snr = like.inner_products_batch(torch.from_numpy(rows).cuda()).cpu().numpy()
    for i, s in enumerate(per_det):
        for d, (dh, h) in enumerate(s):
            assert abs(complex(snr[i, d, 0], snr[i, d, 1]) - dh) < RTOL * h, (i, d)
            assert abs(snr[i, d, 2] - h) < RTOL * h, (i, d)
    print(f"configs[3]: {n} draws, max scaled |dlnL| = {err:.2e}")


def test_configs4_relative_binning_128s_bns_vs_oracle():
    import bench_configs as cfg
    n = 2048
    like, rows, _, _, _ = cfg.build("cfg4_relbin", n)
    draws = like._bench_draws
    got = like.log_likelihood_ratio_rows_host(rows)
    olike = ocr.OracleRelativeBinning(bc.oracle_ifos_like(like.interferometers), like._bench_fiducial,
                                      source_model=ocr.lal_binary_neutron_star_relative_binning,
                                      waveform_arguments=dict(WA_BNS), chi=1, epsilon=0.5)
    assert np.array_equal(olike.bin_freqs, like.bin_freqs)
    ref = bc.oracle_map(olike, draws, n)
    err = _check(got, ref, bc.total_optimal_snr_squared(like, draws), "configs[4] relative binning")
    print(f"configs[4] relative binning ({len(like.bin_freqs) - 1} bins): {n} draws, max scaled |dlnL| = {err:.2e}")


def test_configs4_roq_128s_bns_vs_oracle():
    """ROQ with the bench's synthetic empirical-interpolation basis (N_l = 256, N_q = 96): the oracle builds its own
    weights with one numpy inverse FFT per basis element and detector (roq.py:849-916), the product on the device."""
    import bench_configs as cfg
    n = 1024
    like, rows, _, _, _ = cfg.build("cfg4_roq", n)
    draws = like._bench_draws
    basis = like._bench_basis
    got = like.log_likelihood_ratio_rows_host(rows)
    ifos = like.interferometers
    olike = ocr.OracleROQ(bc.oracle_ifos_like(ifos), basis["linear_matrix"], basis["quadratic_matrix"],
                          basis["frequency_nodes_linear"], basis["frequency_nodes_quadratic"],
                          time_prior=ocl.OracleUniform(T_INJ - 0.05, T_INJ + 0.05),
                          source_model=ocr.binary_neutron_star_roq,
                          waveform_arguments=dict(waveform_approximant="TaylorF2", reference_frequency=20.0),
                          optimal_snrs=[ifo.meta_data.get("optimal_SNR", 30) for ifo in ifos])
    assert np.array_equal(olike.weights["time_samples"], like.weights["time_samples"])
    for ifo in ifos:
        w_ref = olike.weights[ifo.name + "_linear"]
        assert np.max(np.abs(like.weights[ifo.name + "_linear"][0] - w_ref)) < 1e-9 * np.abs(w_ref).max(), ifo.name
    ref = bc.oracle_map(olike, draws, n)
    err = _check(got, ref, bc.total_optimal_snr_squared(like, draws), "configs[4] ROQ")
    print(f"configs[4] ROQ (N_l={basis['linear_matrix'].shape[1]}): {n} draws, max scaled |dlnL| = {err:.2e}")


# ---------------------------------------------------------------------------------------------------------------------
# K1 region boundaries, directed (bb_k1.cuh: rows are split at the rows that contain ka1 / ka2 / kp1 / kp2; kmin / kmax)
# ---------------------------------------------------------------------------------------------------------------------
def _boundary_cases(f_min=20.0):
    """Total masses that put each IMRPhenomD region boundary at chosen bins of the 4 s grid (df = 0.25 Hz, rows of 32
    bins).  With (eta, chi_1, chi_2) fixed every boundary frequency scales as 1 / M."""
    from oracle import phenomd as opd
    q, chi1, chi2 = 0.8, 0.35, -0.2
    c = opd.PhenomDCoefficients(30.0, 30.0 * q, chi1, chi2)
    dimensionless = dict(ka1=opd.AMP_FJOIN_INS, kp1=opd.PHI_FJOIN_INS, ka2=float(c.fmaxCalc), kp2=0.5 * float(c.fRD),
                         kmax=opd.F_CUT)
    df = 0.25
    cases = []
    for name, mf in dimensionless.items():
        for k, tag in ((32 * 9, "row edge"), (32 * 9 + 16, "mid row"), (32 * 2 + 16 + 4, "just above kmin"),
                       (32 * 2 + 8, "below the band (k < kmin)")):
            for eps, side in ((0.0, "on the bin"), (-1e-9, "just below"), (1e-9, "just above")):
                f_target = k * df * (1.0 + eps)
                if name == "kmax" and f_target <= f_min + 0.5:
                    continue        # f_cut <= f_min is the waveform-domain error (its own test)
                total = mf / (f_target * opd.MTSUN_SI)
                cases.append((f"{name} {tag} {side}", total / (1 + q), total * q / (1 + q), chi1, chi2))
    # boundaries above the band: light systems (f_cut = 0.2 / Ms > 1024 Hz, kmax clipped at the row edge k = 4096)
    cases.append(("all boundaries in band, kmax at Nyquist", 12.0, 9.0, chi1, chi2))
    cases.append(("ringdown above Nyquist", 6.0, 5.0, 0.9, 0.9))
    return cases


@pytest.mark.parametrize("f_min", [20.0, 24.0])          # kmin = 80 (mid row) and 96 (row edge)
def test_k1_region_boundaries_on_row_edges_mid_row_and_out_of_band(f_min):
    import torch
    import bilby_b200 as bb
    from bilby_b200 import _lib
    from bilby_b200.gw.detector import InterferometerList
    from bilby_b200.gw.source import lal_binary_black_hole
    from bilby_b200.workloads import INJECTION
    cases = _boundary_cases(f_min)
    n = len(cases)
    rng = np.random.default_rng(7)
    draws = dict(mass_1=np.array([c[1] for c in cases]), mass_2=np.array([c[2] for c in cases]),
                 chi_1=np.array([c[3] for c in cases]), chi_2=np.array([c[4] for c in cases]),
                 luminosity_distance=rng.uniform(500, 3000, n), theta_jn=np.arccos(rng.uniform(-1, 1, n)),
                 psi=rng.uniform(0, np.pi, n), phase=rng.uniform(0, 2 * np.pi, n), ra=rng.uniform(0, 2 * np.pi, n),
                 dec=np.arcsin(rng.uniform(-1, 1, n)), geocent_time=rng.uniform(T_INJ - 0.1, T_INJ + 0.1, n))
    inj = dict(INJECTION)
    start = inj["geocent_time"] - 2.0
    wa = dict(WA_BBH, minimum_frequency=f_min)
    wfg = bb.gw.WaveformGenerator(duration=4.0, sampling_frequency=2048.0, start_time=start,
                                  frequency_domain_source_model=lal_binary_black_hole, waveform_arguments=dict(wa))
    ifos = InterferometerList(["H1", "L1", "V1"])
    for ifo in ifos:
        ifo.minimum_frequency = f_min
    ifos.set_strain_data_from_power_spectral_densities(2048.0, 4.0, start, rng=np.random.default_rng(5))
    ifos.inject_signal(parameters=inj, waveform_generator=wfg)
    like = bb.gw.GravitationalWaveTransient(ifos, wfg)
    olike = ocl.OracleLikelihood(bc.oracle_ifos_like(ifos), waveform_arguments=dict(wa))
    per_det = bc.oracle_map(olike, draws, n, snrs=True)
    rows = torch.from_numpy(np.ascontiguousarray(like.pack(draws))).cuda()

    def compare(snr, what):
        for i, s in enumerate(per_det):
            for d, (dh, h) in enumerate(s):
                assert abs(complex(snr[i, d, 0], snr[i, d, 1]) - dh) < RTOL * h, (what, cases[i][0], d)
                assert abs(snr[i, d, 2] - h) < RTOL * h, (what, cases[i][0], d)

    compare(like.inner_products_batch(rows).cpu().numpy(), "full band")
    # the same samples alone in their block (no neighbours to hide behind) ...
    for i in (0, n // 2, n - 1):
        one = like.inner_products_batch(rows[i:i + 1].contiguous()).cpu().numpy()
        for d, (dh, h) in enumerate(per_det[i]):
            assert abs(complex(one[0, d, 0], one[0, d, 1]) - dh) < RTOL * h, (cases[i][0], d)
    # ... and with the frequency axis cut into shards whose edges fall on a row edge, mid row and ON the boundaries'
    # bins: every shard sees some boundaries inside, some exactly at its edge and some outside (out of shard)
    net = like.device_network
    for edges in ((0, 288, net.n_freq), (0, 304, 305, 1000, net.n_freq), (0, 84, 96, 4096, net.n_freq)):
        total = torch.zeros((n, net.n_det, 3), dtype=torch.float64, device="cuda")
        for b, e in zip(edges[:-1], edges[1:]):
            _lib.check(net.lib.bb_set_frequency_shard(net.ptr, b, e))
            total += like.inner_products_batch(rows)
        _lib.check(net.lib.bb_set_frequency_shard(net.ptr, 0, net.n_freq))
        compare(total.cpu().numpy(), f"shards {edges}")
