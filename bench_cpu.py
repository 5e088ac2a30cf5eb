#!/usr/bin/env python
"""CPU baseline arm for every BASELINE.json configuration (bench.py's `cpu_baseline` leg and `--impl reference`): the
oracle - the CPU restatement of the reference's path under oracle/ - evaluated one parameter dict at a time over a
fork pool of all host cores, exactly how bilby's samplers fan the likelihood out
(bilby/core/sampler/base_sampler.py:772-800).  Never touches CUDA; `kind` is "port" because lalsimulation is absent
(profiles/r2/lal_probe.txt) and /root/reference does not travel to the GPU box.

    python bench_cpu.py --config cfg0|cfg1|cfg2|cfg3|cfg4_relbin|cfg4_roq|cfg4_roq_time [--evals N] [--steps K]

Prints one JSON line: {"config", "value" (evals/s), "cores", "kind", "sample", "ms_per_eval_per_core"}.
The data are synthetic draws of the same shape as the GPU arm's (same detectors, duration, band, marginalisation);
for the two ROQ configurations the weights have the bench's shape (N_l = 256, N_q = 96, 3045 ROQ times) but random
content - building them on the CPU takes minutes and the evaluation time does not depend on their values.
"""
import argparse
import json
import multiprocessing
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_INJ = 1126259642.413
NOISE_SEED = 88170235
DRAW_SEED = 20261017
_STATE = {}
# evaluations per step that keep one step at a few seconds on 16 cores
DEFAULT_EVALS = dict(cfg0=16000, cfg1=4000, cfg2=4000, cfg3=192, cfg4_relbin=24000, cfg4_roq=16000, cfg4_roq_time=480)


def _eval(p):
    return _STATE["like"].log_likelihood_ratio(p)


def _bbh_ifos(ocl, names, duration, fs, start, calibration_points=0):
    from bilby_b200.workloads import INJECTION
    wa = dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0)
    ifos = [ocl.OracleInterferometer(n, fs, duration, start) for n in names]
    rng = np.random.default_rng(NOISE_SEED)
    conv = ocl.convert_to_lal_binary_black_hole_parameters(dict(INJECTION))
    for ifo in ifos:
        ifo.set_gaussian_noise(rng)
    pols = ocl.lal_binary_black_hole(ifos[0].frequency_array, *[conv[k] for k in ocl.SOURCE_ARGS], **wa)
    for ifo in ifos:
        ifo.frequency_domain_strain = ifo.frequency_domain_strain + ifo.get_detector_response(pols, conv)
        if calibration_points:
            ifo.calibration = ocl.OracleCubicSpline(f"recalib_{ifo.name}_", ifo.minimum_frequency, ifo.maximum_frequency,
                                                    calibration_points)
    return ifos, wa


def _bns_ifos(ocl):
    import bench_configs as bc
    duration, fs = 128.0, 4096.0
    start = T_INJ - duration + 2
    wa = dict(waveform_approximant="TaylorF2", reference_frequency=50.0, minimum_frequency=20.0)
    inj = dict(bc.BNS_INJ)
    ifos = [ocl.OracleInterferometer(n, fs, duration, start) for n in ("H1", "L1", "V1")]
    rng = np.random.default_rng(NOISE_SEED)
    conv = ocl.convert_to_lal_binary_black_hole_parameters(inj)
    pols = ocl.lal_binary_neutron_star(ifos[0].frequency_array, *[conv[k] for k in ocl.SOURCE_ARGS], inj["lambda_1"],
                                       inj["lambda_2"], **wa)
    for ifo in ifos:
        ifo.set_gaussian_noise(rng)
        ifo.frequency_domain_strain = ifo.frequency_domain_strain + ifo.get_detector_response(pols, conv)
    return ifos, wa, inj, start


def setup(config, cores):
    """-> (oracle likelihood, draws: dict of arrays generator n -> dict, description)."""
    from oracle import cbc_likelihood as ocl
    from oracle import cbc_reduced as ocr
    from bilby_b200.workloads import INJECTION, draw_bbh_prior
    import bench_configs as bc
    rng = np.random.default_rng(DRAW_SEED)
    if config in ("cfg0", "cfg1", "cfg2"):
        duration = 8.0 if config == "cfg2" else 4.0
        names = ["H1", "L1"] if config == "cfg0" else ["H1", "L1", "V1"]
        start = INJECTION["geocent_time"] - duration + 2
        ifos, wa = _bbh_ifos(ocl, names, duration, 2048.0, start, calibration_points=10 if config == "cfg2" else 0)
        if config == "cfg0":
            like = ocl.OracleLikelihood(ifos, waveform_arguments=wa)
            return like, (lambda n: draw_bbh_prior(n, rng)), "configs[0]: BBH 4s H1+L1, no marginalisation"
        if config == "cfg1":
            like = ocl.OracleLikelihood(ifos, waveform_arguments=wa, phase_marginalization=True,
                                        distance_marginalization=True, distance_prior=ocl.OraclePowerLaw(2, 100.0, 5000.0),
                                        table_processes=cores)
            return like, (lambda n: draw_bbh_prior(n, rng)), "configs[1]: BBH 4s H1L1V1, distance + phase marginalisation"
        like = ocl.OracleLikelihood(ifos, waveform_arguments=wa, time_marginalization=True, jitter_time=True,
                                    time_prior=ocl.OracleUniform(T_INJ - 0.1, T_INJ + 0.1))

        def draws2(n):
            d = draw_bbh_prior(n, rng)
            d["geocent_time"] = np.full(n, float(start))
            d["time_jitter"] = rng.uniform(-1 / 2048.0, 1 / 2048.0, n)
            for name in names:
                for i in range(10):
                    d[f"recalib_{name}_amplitude_{i}"] = rng.normal(0, 0.05, n)
                    d[f"recalib_{name}_phase_{i}"] = rng.normal(0, 0.05, n)
            return d
        return like, draws2, "configs[2]: BBH 8s H1L1V1, time marginalisation (8192-pt FFT x3) + CubicSpline(10)"
    ifos, wa, inj, start = _bns_ifos(ocl)
    if config == "cfg3":
        like = ocl.OracleLikelihood(ifos, source_model=ocl.lal_binary_neutron_star, waveform_arguments=wa)
        return like, (lambda n: bc.bns_draws(n, rng)), "configs[3]: BNS TaylorF2+tides 128s@4096Hz H1L1V1, full grid"
    mc0 = (1.5 * 1.3) ** 0.6 / 2.8 ** 0.2
    fid = dict(inj)
    fid.pop("mass_1"), fid.pop("mass_2")
    fid.update(chirp_mass=mc0, mass_ratio=1.3 / 1.5)
    if config == "cfg4_relbin":
        like = ocr.OracleRelativeBinning(ifos, fid, source_model=ocr.lal_binary_neutron_star_relative_binning,
                                         waveform_arguments=wa, chi=1, epsilon=0.5)
        return like, (lambda n: bc.bns_draws(n, rng, narrow=True)), \
            f"configs[4]: relative binning ({like.number_of_bins} bins) for the 128s BNS"
    # ROQ: weights of the bench's shape, random content (see the module docstring)
    n_lin, n_quad, n_time, step = 256, 96, 3045, 128.0 / 2 ** 21
    tm = config == "cfg4_roq_time"
    freqs = ifos[0].frequency_array[ifos[0].frequency_mask]
    nodes_l = freqs[np.unique(np.geomspace(1, len(freqs) - 1, 4 * n_lin).astype(int))[:n_lin]]
    nodes_q = freqs[np.unique(np.geomspace(1, len(freqs) - 1, 4 * n_quad).astype(int))[:n_quad]]
    wrng = np.random.default_rng(3)
    first = int(np.floor((T_INJ - 0.05 - 2 * ocr.RADIUS_OF_EARTH / ocr.SPEED_OF_LIGHT - 5 * step - start) / step))
    weights = dict(time_samples=np.arange(first, first + n_time) * step)
    for ifo in ifos:
        weights[ifo.name + "_linear"] = (wrng.standard_normal((n_time, n_lin)) + 1j * wrng.standard_normal((n_time, n_lin)))
        weights[ifo.name + "_quadratic"] = np.abs(wrng.standard_normal(n_quad))
    kw = dict(time_marginalization=True, phase_marginalization=True, jitter_time=True, delta_tc=step) if tm else {}
    like = ocr.OracleROQ(ifos, None, None, nodes_l, nodes_q, time_prior=ocl.OracleUniform(T_INJ - 0.05, T_INJ + 0.05),
                         source_model=ocr.binary_neutron_star_roq,
                         waveform_arguments=dict(waveform_approximant="TaylorF2", reference_frequency=20.0),
                         weights=weights, time_space=step, **kw)

    def draws_roq(n):
        d = bc.bns_draws(n, rng, narrow=True)
        if tm:
            d["geocent_time"] = np.full(n, float(start))
            d["time_jitter"] = rng.uniform(-step / 2, step / 2, n)
        return d
    what = "+ time & phase marginalisation (all-times contraction W conj(h))" if tm else "(five-time interpolation)"
    return like, draws_roq, f"configs[4]: ROQ N_l={n_lin}, N_q={n_quad}, {n_time} ROQ times {what}, 128s BNS"


def run(config, n_eval, steps, warmup, cores):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    like, draws_fn, what = setup(config, cores)
    draws = draws_fn(n_eval)
    plist = [{k: float(v[i]) for k, v in draws.items()} for i in range(n_eval)]
    _STATE["like"] = like
    times = []
    with multiprocessing.get_context("fork").Pool(cores) as pool:
        chunk = max(1, n_eval // (cores * 4))
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            out = pool.map(_eval, plist, chunksize=chunk)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    total = sum(times)
    value = n_eval * len(times) / total
    return dict(config=config, value=value, unit="evals/s", cores=cores, kind="port",
                ms_per_eval_per_core=1e3 * cores / value, ms_per_step=1e3 * total / len(times),
                finite_fraction=float(np.mean(np.isfinite(out))),
                sample=f"{what}; {n_eval} draws per step x {len(times)} steps (rng {DRAW_SEED}), fork pool of {cores} "
                       "processes, OMP_NUM_THREADS=1; oracle restatement incl. IMRPhenomD / TaylorF2 (lalsimulation absent)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=sorted(DEFAULT_EVALS))
    ap.add_argument("--evals", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    args = ap.parse_args()
    cores = os.cpu_count() or 1
    n_eval = args.evals or max(cores, DEFAULT_EVALS[args.config] * cores // 16)
    print(json.dumps(run(args.config, n_eval, max(1, args.steps), args.warmup, cores)))


if __name__ == "__main__":
    main()
