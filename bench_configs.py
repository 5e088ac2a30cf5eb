#!/usr/bin/env python
"""Measurement of the BASELINE.json configurations that are NOT the headline bench line (bench.py measures
configs[1]).  One JSON line per configuration: device-resident throughput, end-to-end throughput through the
host entry point, and the roofline of the dominant kernel (CUDA events around its launches, algorithmic work per
SURVEY.md section 8d / DESIGN.md section 4).  Used for the ncu captures under profiles/ as well.

    python bench_configs.py --config cfg0|cfg2|cfg3|cfg4_relbin|cfg4_roq|cfg4_roq_time [--batch N] [--steps K]

  cfg0          configs[0]: BBH 4 s H1+L1, no marginalisation                      (K0 + K1 + K3)
  cfg2          configs[2]: BBH 8 s H1L1V1, time marginalisation + CubicSpline     (K0 + K4)
  cfg3          configs[3]: BNS TaylorF2+tides 128 s @ 4096 Hz H1L1V1             (K0 + K1<TaylorF2>)
  cfg4_relbin   configs[4]: relative binning for the 128 s BNS                     (K0 + K5)
  cfg4_roq      configs[4]: ROQ for the 128 s BNS, synthetic basis                 (K0 + K6)
  cfg4_roq_time configs[4]: ROQ with time marginalisation (dense contraction)      (K0 + K7: hlinear + DMMA contraction + epilogue)
  calmarg       SURVEY 8f rank 4: calibration (1000 curves) + phase marginalisation, 4 s H1L1V1  (K0 + series + DMMA contractions)
  recon         SURVEY 8f rank 2: marginalised-parameter reconstruction, 4 s H1L1V1, time + distance + phase
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench as hb  # noqa: E402  (helpers: ClockSampler, MTSUN)

T_INJ = 1126259642.413
BNS_INJ = dict(mass_1=1.5, mass_2=1.3, chi_1=0.02, chi_2=0.01, luminosity_distance=100.0, theta_jn=0.4, psi=2.659,
               phase=1.3, geocent_time=T_INJ, ra=1.375, dec=-1.2108, lambda_1=400.0, lambda_2=600.0)
NOISE_SEED = 88170235
_ROQ_BASIS = {}


def bns_draws(n, rng, narrow=False):
    if narrow:     # the box a relative-binning / ROQ analysis of this event would use
        mc0 = (1.5 * 1.3) ** 0.6 / 2.8 ** 0.2
        mc = mc0 * (1 + rng.uniform(-1e-4, 1e-4, n))
        q = rng.uniform(0.8, 0.95, n)
        t = rng.uniform(T_INJ - 2e-3, T_INJ + 2e-3, n)
    else:
        mc, q, t = rng.uniform(1.15, 1.25, n), rng.uniform(0.5, 1.0, n), rng.uniform(T_INJ - 0.1, T_INJ + 0.1, n)
    return dict(chirp_mass=mc, mass_ratio=q, chi_1=rng.uniform(-0.05, 0.05, n), chi_2=rng.uniform(-0.05, 0.05, n),
                luminosity_distance=(10.0 ** 3 + rng.uniform(0, 1, n) * (500.0 ** 3 - 10.0 ** 3)) ** (1 / 3),
                theta_jn=np.arccos(rng.uniform(-1, 1, n)), psi=rng.uniform(0, np.pi, n),
                phase=rng.uniform(0, 2 * np.pi, n), ra=rng.uniform(0, 2 * np.pi, n),
                dec=np.arcsin(rng.uniform(-1, 1, n)), geocent_time=t,
                lambda_1=rng.uniform(0, 5000, n), lambda_2=rng.uniform(0, 5000, n))


def make_ifos(names, fs, duration, start, wfg, inj):
    from bilby_b200.gw.detector import InterferometerList
    ifos = InterferometerList(names)
    ifos.set_strain_data_from_power_spectral_densities(fs, duration, start, rng=np.random.default_rng(NOISE_SEED))
    ifos.inject_signal(parameters=dict(inj), waveform_generator=wfg)
    return ifos


def build(config, n):
    """-> (likelihood, rows [n,16], cal or None, flop_per_eval(rows) callable, description dict)."""
    import bilby_b200 as bb
    from bilby_b200.core.prior import PriorDict, Uniform, PowerLaw
    from bilby_b200.gw import conversion, source
    from bilby_b200.workloads import INJECTION, draw_bbh_prior
    rng = np.random.default_rng(hb.DRAW_SEED)
    if config == "cfg1":        # the headline configuration of bench.py, here for profiling captures
        like = hb.build_likelihood()
        rows = hb.draw_rows(like, n, hb.DRAW_SEED)
        df = 0.25

        def flop1(rows_):
            bins = hb.active_bins(rows_, df, 4097)
            return bins * hb.FLOP_PER_BIN + len(rows_) * hb.EPILOGUE_FLOP, bins / len(rows_)
        return like, rows, None, flop1, dict(workload=hb.WORKLOAD, kernel="bb_inner_product_kernel<3,IMRPhenomD>")
    if config in ("cfg0", "cfg2", "calmarg", "calmarg_time"):
        duration = 8.0 if config == "cfg2" else 4.0
        names = ["H1", "L1"] if config == "cfg0" else ["H1", "L1", "V1"]
        fs = 2048.0
        inj = dict(INJECTION)
        start = inj["geocent_time"] - duration + 2
        wfg = bb.gw.WaveformGenerator(duration=duration, sampling_frequency=fs, start_time=start,
                                      frequency_domain_source_model=source.lal_binary_black_hole,
                                      waveform_arguments=dict(waveform_approximant="IMRPhenomD",
                                                              reference_frequency=50.0, minimum_frequency=20.0))
        ifos = make_ifos(names, fs, duration, start, wfg, inj)
        draws = draw_bbh_prior(n, rng)
        df = 1.0 / duration
        msec = None
        if config in ("calmarg", "calmarg_time"):
            from bilby_b200.core.prior import Gaussian
            from bilby_b200.core.utils import random as bb_random
            bb_random.seed(hb.NOISE_SEED)
            from bilby_b200.gw.detector.calibration import CubicSpline
            n_curves = 1000
            pri = dict(phase=Uniform(0, 2 * np.pi, "phase"))
            for ifo in ifos:
                ifo.calibration_model = CubicSpline(f"recalib_{ifo.name}_", ifo.minimum_frequency,
                                                    ifo.maximum_frequency, 10)
                for i in range(10):
                    for kind in ("amplitude", "phase"):
                        key = f"recalib_{ifo.name}_{kind}_{i}"
                        pri[key] = Gaussian(0.0, 0.05, key)
            tm = config == "calmarg_time"
            if tm:
                pri["geocent_time"] = Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time")
            like = bb.gw.GravitationalWaveTransient(ifos, wfg, phase_marginalization=True, calibration_marginalization=True,
                                                    time_marginalization=tm, jitter_time=True,
                                                    number_of_response_curves=n_curves, priors=PriorDict(pri))
            if tm:
                draws["geocent_time"] = np.full(n, float(start))
                draws["time_jitter"] = rng.uniform(-1 / fs, 1 / fs, n)
            rows = like.pack(draws)
            ldk = (len(ifos[0].frequency_array) + 7) // 8 * 8

            def flop_cm(rows_):
                # algorithmic: (8 + 2) flop per (ACTIVE bin, detector, curve) - bins above the waveform's cut-off
                # 0.2 / (M t_sun) contribute nothing (the contraction is trimmed to each chunk's active window)
                msec_ = (rows_[:, 0] + rows_[:, 1]) * hb.MTSUN
                fmp_ = np.minimum(fs / 2, 0.2 / msec_)
                k1_ = np.minimum(np.floor(fmp_ / df), np.floor(fs / 2 / df) + 1)
                bins_ = np.maximum(k1_ - np.ceil(20.0 / df), 0)
                if tm:
                    # per (sample, curve): series = sum_det X C on the active bins (8 flop each), one 4096-point
                    # transform (5 N log2 N), 120 flop per time inside the prior; <h|h> per curve as above
                    nfft = int(round(duration * fs / 2))
                    per_pair = 5.0 * nfft * np.log2(nfft) + 120.0 * int(0.2 * fs / 2)
                    return (float(np.sum(bins_)) * (8 + 2) * 3 * n_curves + len(rows_) * n_curves * per_pair,
                            float(np.mean(bins_) * 3))
                return float(np.sum(bins_)) * (8 + 2) * 3 * n_curves, float(np.mean(bins_) * 3)
            if tm:
                return like, rows, None, flop_cm, dict(
                    workload=f"SURVEY 8f rank 4: time + calibration (+ phase) marginalisation over {n_curves} CubicSpline(10) "
                             "response curves, BBH 4s@2048Hz H1L1V1 IMRPhenomD: one 4096-point transform per (sample, curve)",
                    kernel="bb_calmarg_series_kernel + bb_gemm_nt_kernel<real> (DMMA) + bb_calmarg_time_kernel + bb_calmarg_lse_kernel",
                    bound="shared memory / FP64 (in-shared-memory FFT per response curve)", n_curves=n_curves)
            return like, rows, None, flop_cm, dict(
                workload=f"SURVEY 8f rank 4: calibration marginalisation over {n_curves} CubicSpline(10) response curves + "
                         "phase marginalisation, BBH 4s@2048Hz H1L1V1 IMRPhenomD: [batch x 3*4104] x [3*4104 x 1000] "
                         "complex + real contraction per chunk", kernel="bb_calmarg_series_kernel + bb_gemm_nt_kernel<complex>, <real> (DMMA) + epilogue",
                bound="fp64 tensor (DMMA)", n_curves=n_curves)
        if config == "cfg0":
            like = bb.gw.GravitationalWaveTransient(ifos, wfg)
            rows = like.pack(draws)
            like._bench_draws = draws
            cal = None
            per_bin, extra = 240 + 30 * 2, 10
            work = "configs[0]: BBH 4s@2048Hz H1+L1 IMRPhenomD, no marginalisation"
            kernel = "bb_inner_product_kernel<2,IMRPhenomD>"
        else:
            from bilby_b200.gw.detector.calibration import CubicSpline
            for ifo in ifos:
                ifo.calibration_model = CubicSpline(f"recalib_{ifo.name}_", ifo.minimum_frequency,
                                                    ifo.maximum_frequency, 10)
            pri = PriorDict(dict(geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time")))
            like = bb.gw.GravitationalWaveTransient(ifos, wfg, time_marginalization=True, jitter_time=True, priors=pri)
            draws["geocent_time"] = np.full(n, float(start))
            draws["time_jitter"] = rng.uniform(-1 / fs, 1 / fs, n)
            crng = np.random.default_rng(99)
            for name in names:
                for i in range(10):
                    draws[f"recalib_{name}_amplitude_{i}"] = crng.normal(0, 0.05, n)
                    draws[f"recalib_{name}_phase_{i}"] = crng.normal(0, 0.05, n)
            rows = like.pack(draws)
            like._bench_draws = draws
            cal = like._cal_from_parameters(draws, n, np)
            per_bin = 240 + 70 * 3
            n_prior = int(0.2 * fs / 2)
            extra = 5 * 8192 * 13 + 120 * n_prior
            work = "configs[2]: BBH 8s@2048Hz H1L1V1 IMRPhenomD, time marginalisation (8192-pt FFT) + CubicSpline(10)"
            kernel = "bb_series_fill_kernel<3,IMRPhenomD,CAL> || bb_series_fft_kernel (two streams)"

        def flop(rows_):
            msec = (rows_[:, 0] + rows_[:, 1]) * hb.MTSUN
            fmp = np.minimum(fs / 2, 0.2 / msec)
            k1 = np.minimum(np.floor(fmp / df), np.floor(fs / 2 / df) + 1)
            bins = np.maximum(k1 - np.ceil(20.0 / df), 0)
            return float(np.sum(bins) * per_bin + len(rows_) * extra), float(np.mean(bins))
        return like, rows, cal, flop, dict(workload=work, kernel=kernel)

    # ---- 128 s BNS family
    duration, fs = 128.0, 4096.0
    names = ["H1", "L1", "V1"]
    inj = dict(BNS_INJ)
    start = T_INJ - duration + 2
    conv = conversion.convert_to_lal_binary_neutron_star_parameters
    wa = dict(waveform_approximant="TaylorF2", reference_frequency=50.0, minimum_frequency=20.0)
    wfg_full = bb.gw.WaveformGenerator(duration=duration, sampling_frequency=fs, start_time=start,
                                       frequency_domain_source_model=source.lal_binary_neutron_star,
                                       parameter_conversion=conv, waveform_arguments=dict(wa))
    ifos = make_ifos(names, fs, duration, start, wfg_full, inj)
    n_masked = int(ifos[0].frequency_mask.sum())
    if config == "cfg3":
        like = bb.gw.GravitationalWaveTransient(ifos, wfg_full)
        like._bench_draws = bns_draws(n, rng)
        rows = like.pack(like._bench_draws)
        per = (170 + 30 * 3) * n_masked
        return like, rows, None, (lambda r: (float(len(r)) * per, float(n_masked))), dict(
            workload="configs[3]: BNS TaylorF2+tides 128s@4096Hz H1L1V1 (259585 masked bins/detector), no "
                     "marginalisation, one GPU holds the whole frequency axis", kernel="bb_inner_product_kernel<3,TaylorF2>")
    mc0 = (1.5 * 1.3) ** 0.6 / 2.8 ** 0.2
    fid = dict(inj)
    fid.pop("mass_1"), fid.pop("mass_2")
    fid.update(chirp_mass=mc0, mass_ratio=1.3 / 1.5)
    if config == "cfg4_relbin":
        wfg = bb.gw.WaveformGenerator(duration=duration, sampling_frequency=fs, start_time=start,
                                      frequency_domain_source_model=source.lal_binary_neutron_star_relative_binning,
                                      parameter_conversion=conv, waveform_arguments=dict(wa))
        like = bb.gw.likelihood.RelativeBinningGravitationalWaveTransient(ifos, wfg, fiducial_parameters=fid,
                                                                          epsilon=0.5, chi=1)
        like._bench_draws = bns_draws(n, rng, narrow=True)
        like._bench_fiducial = fid
        rows = like.pack(like._bench_draws)
        ne = len(like.bin_freqs)
        per = (170 + 105 * 3) * ne
        return like, rows, None, (lambda r: (float(len(r)) * per, float(ne))), dict(
            workload=f"configs[4]: relative binning (epsilon=0.5, chi=1, {ne - 1} bins) for the 128s BNS, H1L1V1",
            kernel="bb_relbin_kernel<3,TaylorF2>", n_edges=ne)
    if config in ("mb", "mb_time"):
        wfg = bb.gw.WaveformGenerator(duration=duration, sampling_frequency=fs, start_time=start,
                                      frequency_domain_source_model=source.binary_neutron_star_frequency_sequence,
                                      parameter_conversion=conv,
                                      waveform_arguments=dict(waveform_approximant="TaylorF2", reference_frequency=50.0))
        pri = PriorDict(dict(geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time")))
        t0 = time.time()
        tm = config == "mb_time"
        if tm:
            pri["phase"] = Uniform(0, 2 * np.pi, "phase")
        like = bb.gw.likelihood.MBGravitationalWaveTransient(ifos, wfg, reference_chirp_mass=1.15, priors=pri,
                                                             time_marginalization=tm, phase_marginalization=tm)
        npts = len(like.banded_frequency_points)
        sys.stderr.write(f"multi-banding: {like.number_of_bands} bands, {npts} points, set up in {time.time() - t0:.1f} s\n")
        draws = bns_draws(n, rng, narrow=True)
        if tm:
            draws["geocent_time"] = np.full(n, float(start))
            draws["time_jitter"] = rng.uniform(-1 / fs, 1 / fs, n)
            n_times = int(np.ceil(0.2 / like._delta_tc)) + 8
            rows = like.pack(draws)
            per = (170 + 78 * 3) * npts + 8.0 * npts * n_times
            return like, rows, None, (lambda r: (float(len(r)) * per, float(npts))), dict(
                workload=f"SURVEY 8f rank 4: multi-banded likelihood + time & phase marginalisation ({npts} banded points x "
                         f"{n_times} times of the {int(like.Nbs[-1]) // 2}-point transform) for the 128s BNS, H1L1V1",
                kernel="bb_mb_series_kernel + bb_gemm_nt_kernel<complex> (DMMA) + bb_mb_time_marg_kernel",
                bound="fp64 tensor (DMMA)", n_points=npts, n_times=n_times)
        rows = like.pack(draws)
        per = (170 + 78 * 3) * npts          # per (point, detector): sincospi 60, K h 6, <d|h> 8, <h|h> 4
        return like, rows, None, (lambda r: (float(len(r)) * per, float(npts))), dict(
            workload=f"SURVEY 8f rank 4: multi-banded likelihood ({like.number_of_bands} bands, {npts} banded points vs "
                     f"{n_masked} grid bins) for the 128s BNS, H1L1V1", kernel="bb_relbin_kernel<3,TaylorF2,edge form>",
            n_points=npts, n_bands=int(like.number_of_bands))
    # ---- ROQ with a synthetic empirical-interpolation basis built from device waveforms (set-up, untimed)
    import torch
    tm = config == "cfg4_roq_time"
    n_lin, n_quad, n_train = (256, 96, 384)
    freqs = ifos[0].frequency_array[ifos[0].frequency_mask]
    train = bns_draws(n_train, np.random.default_rng(5), narrow=True)
    h = wfg_full._get_handle()
    from bilby_b200 import _lib
    from bilby_b200.gw import _params
    approx, f_ref, f_min, f_max = wfg_full.approximant_config()
    _lib.check(h.lib.bb_set_waveform(h.ptr, approx, 20.0, f_min, f_max))
    conv_tr, _ = conv(dict(train))
    conv_tr["luminosity_distance"] = np.ones(n_train)
    conv_tr["theta_jn"] = np.zeros(n_train)
    conv_tr["phase"] = np.zeros(n_train)
    rows_tr = torch.from_numpy(_params.pack_rows(conv_tr, n_train, np)).cuda()
    fr = torch.from_numpy(np.ascontiguousarray(freqs)).cuda()
    out = torch.empty((n_train, 2, len(freqs), 2), dtype=torch.float64, device="cuda")
    _lib.check(h.lib.bb_frequency_sequence_strain_device(h.ptr, rows_tr.data_ptr(), n_train, fr.data_ptr(), len(freqs),
                                                         float(freqs[0]), out.data_ptr(), None))
    hp = torch.view_as_complex(out[:, 0].contiguous())           # [n_train, n_freq]
    del out

    def interpolant(tr, nb):
        tr = tr / torch.linalg.norm(tr, dim=1, keepdim=True)
        _, _, vh = torch.linalg.svd(tr, full_matrices=False)
        v = vh[:nb].T.contiguous()                                 # [n_freq, nb]
        nodes = [int(torch.argmax(v[:, 0].abs()))]
        for j in range(1, nb):
            idx = torch.tensor(nodes, device=v.device)
            c = torch.linalg.solve(v[idx, :j], v[idx, j])
            r = v[:, j] - v[:, :j] @ c
            r[idx] = 0
            nodes.append(int(torch.argmax(r.abs())))
        nodes = np.array(sorted(nodes))
        idx = torch.tensor(nodes, device=v.device)
        b = v @ torch.linalg.inv(v[idx, :])
        return b.cpu().numpy(), nodes
    key = (n_lin, n_quad, n_train, len(freqs))
    if key not in _ROQ_BASIS:           # the two ROQ configurations of one process share the synthetic basis
        _ROQ_BASIS.clear()
        _ROQ_BASIS[key] = interpolant(hp, n_lin) + interpolant((hp.abs() ** 2).to(torch.complex128), n_quad)
    bl, nl, bq, nq = _ROQ_BASIS[key]
    del hp
    torch.cuda.empty_cache()
    wfg = bb.gw.WaveformGenerator(duration=duration, sampling_frequency=fs, start_time=start,
                                  frequency_domain_source_model=source.binary_neutron_star_roq,
                                  parameter_conversion=conv,
                                  waveform_arguments=dict(waveform_approximant="TaylorF2", reference_frequency=20.0,
                                                          frequency_nodes_linear=freqs[nl],
                                                          frequency_nodes_quadratic=freqs[nq]))
    pri = dict(geocent_time=Uniform(T_INJ - 0.05, T_INJ + 0.05, "geocent_time"))
    kw = {}
    if tm:
        pri["phase"] = Uniform(0, 2 * np.pi, "phase")
        kw = dict(time_marginalization=True, phase_marginalization=True, jitter_time=True)
    t0 = time.time()
    like = bb.gw.likelihood.ROQGravitationalWaveTransient(ifos, wfg, PriorDict(pri), linear_matrix=bl,
                                                          quadratic_matrix=bq, **kw)
    n_time = len(like.weights["time_samples"])
    sys.stderr.write(f"ROQ weights: {n_time} time samples x {n_lin} linear, {n_quad} quadratic, built in "
                     f"{time.time() - t0:.1f} s\n")
    draws = bns_draws(n, rng, narrow=True)
    if tm:
        draws["geocent_time"] = np.full(n, float(start))
        draws["time_jitter"] = rng.uniform(-like._delta_tc / 2, like._delta_tc / 2, n)
        per = (170 + 20) * n_lin + 100 * n_quad + 8 * n_time * n_lin * 3 + len(like._times) * (3 * 40 + 100)
        work = (f"configs[4]: ROQ (synthetic basis N_l={n_lin}, N_q={n_quad}, {n_time} ROQ times) + time&phase "
                f"marginalisation ({len(like._times)} times) for the 128s BNS: dense W conj(h) contraction")
        kernel = "bb_roq_hlinear_kernel + bb_gemm_nt_kernel<complex> (DMMA) || bb_roq_time_marg_kernel (aux stream)"
        # rows of W the device contracts: only those a time inside the prior (+- light travel time) can touch
        dmax = max(float(np.linalg.norm(ifo.vertex)) for ifo in ifos) / 299792458.0
        ts = like.weights["time_samples"]
        step = ts[1] - ts[0]
        r_lo = max(0, int(np.floor((T_INJ - 0.05 - start - dmax - ts[0]) / step)) - 3)
        r_hi = min(n_time - 1, int(np.floor((T_INJ + 0.05 - start + dmax - ts[0]) / step)) + 3)
        n_rows = r_hi - (r_lo // 64) * 64 + 1
    else:
        per = (170 + 60 + 120) * n_lin + (100 + 6) * n_quad
        work = (f"configs[4]: ROQ (synthetic basis N_l={n_lin}, N_q={n_quad}, {n_time} ROQ times) for the 128s BNS, "
                "H1L1V1")
        kernel = "bb_roq_kernel<3,TaylorF2>"
    rows = like.pack(draws)
    like._bench_draws = draws
    like._bench_basis = dict(linear_matrix=bl, quadratic_matrix=bq, frequency_nodes_linear=freqs[nl],
                             frequency_nodes_quadratic=freqs[nq])
    extra = dict(bound="fp64 tensor (DMMA)", n_time_contracted=n_rows,
                 executed_gemm_flop_per_eval=8.0 * n_rows * n_lin * 3) if tm else {}
    return like, rows, None, (lambda r: (float(len(r)) * per, float(n_lin))), dict(
        workload=work, kernel=kernel, n_linear=n_lin, n_quadratic=n_quad, n_time=n_time, **extra)


DEFAULT_BATCH = dict(calmarg=75776, cfg0=1_000_000, cfg1=1_000_000, cfg2=100_000, cfg3=8192, cfg4_relbin=1_000_000, cfg4_roq=1_000_000,
                     cfg4_roq_time=65536, mb=65536, mb_time=16384, calmarg_time=1024)


def recon_bench(n, steps):
    """Rows per second of generate_posterior_samples_from_marginalized_likelihood_batch (host arrays in, host arrays
    out) next to the oracle's restatement of the reference's per-row loop on one host core."""
    import torch
    import bilby_b200 as bb
    from bilby_b200.core.prior import PriorDict, Uniform, PowerLaw
    from bilby_b200.gw import source
    from bilby_b200.workloads import INJECTION
    duration, fs, names = 4.0, 2048.0, ["H1", "L1", "V1"]
    inj = dict(INJECTION)
    start = inj["geocent_time"] - duration + 2
    wa = dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0)
    wfg = bb.gw.WaveformGenerator(duration=duration, sampling_frequency=fs, start_time=start,
                                  frequency_domain_source_model=source.lal_binary_black_hole, waveform_arguments=wa)
    ifos = make_ifos(names, fs, duration, start, wfg, inj)
    pri = PriorDict(dict(geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time"),
                         phase=Uniform(0, 2 * np.pi, "phase"),
                         luminosity_distance=PowerLaw(2, 100.0, 5000.0, "luminosity_distance")))
    like = bb.gw.GravitationalWaveTransient(ifos, wfg, time_marginalization=True, distance_marginalization=True,
                                            phase_marginalization=True, jitter_time=True, priors=pri)
    rng = np.random.default_rng(hb.DRAW_SEED)
    rows = dict(chirp_mass=28.0956 + rng.normal(0, 0.05, n), mass_ratio=np.clip(29 / 36 + rng.normal(0, 0.02, n), 0.2, 1),
                chi_1=0.4 + rng.normal(0, 0.02, n), chi_2=0.3 + rng.normal(0, 0.02, n),
                luminosity_distance=np.full(n, float(like._ref_dist)), theta_jn=0.4 + rng.normal(0, 0.1, n),
                psi=np.full(n, 2.659), phase=np.zeros(n), ra=1.375 + rng.normal(0, 0.02, n),
                dec=-1.2108 + rng.normal(0, 0.02, n), geocent_time=T_INJ + rng.normal(0, 1e-3, n),
                time_jitter=rng.uniform(-1 / fs, 1 / fs, n))
    uni = rng.uniform(0, 1, (n, 3))
    for _ in range(2):
        new = like.generate_posterior_samples_from_marginalized_likelihood_batch(rows, uniforms=uni)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        new = like.generate_posterior_samples_from_marginalized_likelihood_batch(rows, uniforms=uni)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    # CPU: the oracle's restatement of base.py:502-773, one row at a time (what the reference's pool workers run)
    from oracle import cbc_likelihood as ocl
    oifos = [ocl.OracleInterferometer(nm, fs, duration, start) for nm in names]
    for o, ifo in zip(oifos, ifos):
        o.frequency_domain_strain = np.asarray(ifo.frequency_domain_strain)
    table = np.asarray(like._dist_margd_loglikelihood_array)
    olike = ocl.OracleLikelihood(oifos, waveform_arguments=wa, time_marginalization=True, distance_marginalization=True,
                                 phase_marginalization=True, distance_prior=ocl.OraclePowerLaw(2, 100.0, 5000.0),
                                 time_prior=ocl.OracleUniform(T_INJ - 0.1, T_INJ + 0.1), lookup_table=table)
    m = 8
    t0 = time.perf_counter()
    ref = np.array([[v for v in (lambda q: (q["geocent_time"], q["luminosity_distance"], q["phase"]))(
        olike.generate_posterior_sample_from_marginalized_likelihood({k: float(v[i]) for k, v in rows.items()}, uni[i]))]
        for i in range(m)])
    cpu_dt = (time.perf_counter() - t0) / m
    got = np.stack([new["geocent_time"][:m], new["luminosity_distance"][:m], new["phase"][:m]], axis=1)
    err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)))
    print(json.dumps(dict(metric="reconstructed posterior rows/sec", value=n / dt, unit="rows/s", n_gpus=1, steps=steps,
                          config=dict(workload="SURVEY 8f rank 2: time + distance + phase reconstruction, BBH 4s H1L1V1, "
                                               "16384 Hz time posterior (16 modulated 4096-pt transforms per row)",
                                      batch=n),
                          ms_per_step=dt * 1e3, dtype="f64", data="synthetic",
                          cpu_baseline=dict(value=1.0 / cpu_dt, unit="rows/s", cores=1, kind="port",
                                            sample=f"{m} rows, oracle restatement of base.py:502-773"),
                          max_rel_diff_vs_oracle=err)))


def run_config(config, n, steps, warmup, world=1, rank=0, local=0, dist=None, clocks=True):
    """Times one configuration on this rank's GPU (process group, if any, already initialised by the caller).
    Returns the JSON-able line on rank 0, None elsewhere.  N > 1: samples are independent units -> every rank evaluates
    its own batch of draws (weak scaling, no data-path collective); value = all ranks' evaluations / max-over-ranks
    device time."""
    import torch
    from bilby_b200 import _lib
    seed0 = hb.DRAW_SEED
    hb.DRAW_SEED = seed0 + rank
    try:
        like, rows_np, cal_np, flop, desc = build(config, n)
    finally:
        hb.DRAW_SEED = seed0
    net = like.device_network
    lib = net.lib
    rows_np = np.ascontiguousarray(rows_np)
    rows_dev = torch.from_numpy(rows_np).cuda()
    cal_dev = torch.from_numpy(np.ascontiguousarray(cal_np)).cuda() if cal_np is not None else None
    stream = torch.cuda.current_stream()
    peak = ctypes.c_double(0.0)
    tensor_bound = "DMMA" in desc.get("bound", "")
    _lib.check((lib.bb_fp64_tensor_peak if tensor_bound else lib.bb_fp64_peak)(net.ptr, ctypes.byref(peak)))
    warmup = max(3, warmup)

    def step_device():
        return like._evaluate_device(rows_dev, cal_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    for _ in range(warmup):
        out = step_device()
    barrier()
    _lib.check(lib.bb_profile_enable(net.ptr, 1))
    torch.cuda.cudart().cudaProfilerStart()        # `ncu --profile-from-start off` skips the set-up kernels
    launches0 = lib.bb_launch_count(net.ptr)
    sampler = hb.ClockSampler(local) if clocks else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        out = step_device()
    e1.record(stream)
    barrier()
    torch.cuda.cudart().cudaProfilerStop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.bb_launch_count(net.ptr) - launches0
    k_ms, k_n = ctypes.c_double(0.0), ctypes.c_long(0)
    _lib.check(lib.bb_profile_read(net.ptr, ctypes.byref(k_ms), ctypes.byref(k_n)))
    _lib.check(lib.bb_profile_enable(net.ptr, 0))
    # end to end through the host entry point, from / to page-locked host buffers (the bench contract's e2e)
    from bilby_b200.core.utils import pinned_empty
    rows_pin = pinned_empty(rows_np.shape)
    rows_pin[...] = rows_np
    cal_pin = None
    if cal_np is not None:
        cal_pin = pinned_empty(np.shape(cal_np))
        cal_pin[...] = cal_np
    res = pinned_empty(n)
    like.log_likelihood_ratio_rows_host(rows_pin, cal_pin, out=res)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        like.log_likelihood_ratio_rows_host(rows_pin, cal_pin, out=res)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    clock_info = sampler.stop() if sampler else None
    front = None
    if config in ("cfg4_relbin", "cfg4_roq"):
        front = front_end_rate(like, n, steps, stream, torch)
    if world > 1:
        dist.barrier()
    del like, net, rows_dev, cal_dev, out
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    n_all = n * world
    total_flop, units = flop(rows_np)
    k_avg = k_ms.value / max(1, k_n.value) * (k_n.value / steps)       # dominant-kernel time per step
    achieved = total_flop / (k_avg * 1e-3) / 1e12 if k_avg > 0 else 0.0
    fin = np.isfinite(res)
    return dict(metric="log-likelihood evals/sec", config=dict(desc, batch=n, partition=f"samples x{world}"),
                value=n_all * steps / (ms_total * 1e-3),
                unit="evals/s", n_gpus=world, scaling="weak", steps=steps, warmup=warmup, ms_per_step=ms_total / steps,
                dtype="f64", data="synthetic", clocks=clock_info,
                e2e=dict(value=n_all * steps / (e2e_ms * 1e-3), unit="evals/s", ms_per_step=e2e_ms / steps,
                         h2d_bytes_per_step=int(rows_np.nbytes + (cal_np.nbytes if cal_np is not None else 0)),
                         d2h_bytes_per_step=n * 8),
                gpu_launches=int(launches),
                roofline=dict(bound="fp64 tensor" if tensor_bound else "fp64", achieved=achieved, peak=peak.value, unit="TFLOP/s",
                              frac=achieved / peak.value if peak.value else None, kernel=desc["kernel"],
                              kernel_ms_per_step=k_avg, kernel_share_of_step=k_avg / (ms_total / steps),
                              algorithmic_flop_per_step=total_flop, units_per_eval=units,
                              peak_source="in-run DMMA stream kernel (bb_fp64_tensor_peak)" if tensor_bound
                              else "in-run DFMA stream kernel (bb_fp64_peak)"),
                checksum_lnl=float(np.sum(res[fin])), finite_fraction=float(fin.mean()),
                **(dict(device_front_end=front) if front else {}))


def front_end_rate(like, n, steps, stream, torch):
    """The device-resident sampling front end (csrc/bb_sampling.cuh) in front of the same likelihood: unit-cube points
    that already live on the device -> PriorDict.rescale + parameter conversion + rows (one kernel) -> lnL.  The
    priors are the box bns_draws(narrow=True) samples from; nothing crosses PCIe."""
    from bilby_b200.core.prior import PriorDict, Uniform, PowerLaw, Sine, Cosine
    from bilby_b200.core.sampler import BatchedLikelihood
    mc0 = (1.5 * 1.3) ** 0.6 / 2.8 ** 0.2
    pri = PriorDict(dict(
        chirp_mass=Uniform(mc0 * (1 - 1e-4), mc0 * (1 + 1e-4), "chirp_mass"), mass_ratio=Uniform(0.8, 0.95, "mass_ratio"),
        chi_1=Uniform(-0.05, 0.05, "chi_1"), chi_2=Uniform(-0.05, 0.05, "chi_2"),
        luminosity_distance=PowerLaw(2, 10.0, 500.0, "luminosity_distance"), theta_jn=Sine(name="theta_jn"),
        psi=Uniform(0, np.pi, "psi"), phase=Uniform(0, 2 * np.pi, "phase"), ra=Uniform(0, 2 * np.pi, "ra"),
        dec=Cosine(name="dec"), geocent_time=Uniform(T_INJ - 2e-3, T_INJ + 2e-3, "geocent_time"),
        lambda_1=Uniform(0, 5000, "lambda_1"), lambda_2=Uniform(0, 5000, "lambda_2")))
    batched = BatchedLikelihood(like, pri)
    u = torch.rand((n, batched.ndim), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(7))
    for _ in range(3):
        lnl = batched.log_likelihood_from_unit_cube(u)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        lnl = batched.log_likelihood_from_unit_cube(u)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    fin = torch.isfinite(lnl)
    return dict(value=n / (ms * 1e-3), unit="evals/s", ms_per_step=ms, finite_fraction=float(fin.double().mean().item()),
                what="unit-cube points resident on the device -> bb_rows_from_unit_cube_device (prior rescale + "
                     "conversion + rows) -> lnL; no host<->device traffic")


def run_frequency_sharded(n, steps, warmup, world, rank, dist, exchange="fused"):
    """configs[3] with the frequency axis in `world` contiguous shards (strong scaling: every rank evaluates ALL n
    samples on its bin range; fused peer-memory exchange or one NCCL all-reduce; replicated epilogue)."""
    import torch
    from bilby_b200.parallel import FrequencyShardedLikelihood
    like, rows_np, _, flop, desc = build("cfg3", n)
    rows = torch.from_numpy(np.ascontiguousarray(rows_np)).cuda()
    check = like.log_likelihood_ratio_batch(rows[:256]).cpu().numpy()          # unsharded, before the shard is set
    sharded = FrequencyShardedLikelihood(like, rank, world, fused_max_rows=n if exchange == "fused" else 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, warmup)):
        out = sharded.log_likelihood_ratio_rows(rows)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = sharded.log_likelihood_ratio_rows(rows)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    snrs = like.inner_products_batch(rows)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(steps):
        if world > 1:
            dist.all_reduce(snrs)
    a1.record()
    barrier()
    ar = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ar, op=dist.ReduceOp.MAX)
    sharded.check_exchange()
    err = float(np.max(np.abs(out[:256].cpu().numpy() - check)) / np.max(np.abs(check)))
    fused = bool(sharded.fused)
    k_begin, k_end = sharded.k_begin, sharded.k_end
    del sharded, like, rows, out, snrs
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    total_flop, _ = flop(rows_np)
    t = float(ms.item()) * 1e-3 / steps
    return dict(metric="log-likelihood evals/sec (TaylorF2+tides 128 s H1L1V1, frequency-sharded)", value=n / t,
                unit="evals/s", n_gpus=world, steps=steps, warmup=max(3, warmup), ms_per_step=t * 1e3, scaling="strong",
                config=dict(workload=desc["workload"], batch=n,
                            partition=f"frequency axis in {world} contiguous shards, bins [{k_begin}, {k_end}) on rank 0; "
                                      f"exchange of {n * 3 * 3 * 8} bytes per rank per step",
                            exchange=("fused into K1 (peer-memory stores over NVLink + flag round)" if fused
                                      else "NCCL all-reduce")),
                exchange_us_per_step_nccl_allreduce_alone=1e3 * float(ar.item()) / steps,
                algorithmic_tflops=total_flop / t / 1e12, max_rel_diff_vs_unsharded=err)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=sorted(DEFAULT_BATCH) + ["recon"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench_configs.py needs a CUDA device (bilby_b200 has no CPU path)")
    if args.config == "recon":
        return recon_bench(args.batch or 20000, args.steps)
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = run_config(args.config, args.batch or DEFAULT_BATCH[args.config], args.steps, args.warmup, world, rank, local, dist)
    if world > 1:
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
