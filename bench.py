#!/usr/bin/env python
"""Benchmark of the B200-native CBC likelihood hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): BBH, IMRPhenomD, H1L1V1, 4 s @ 2048 Hz, f_min 20 Hz, Gaussian noise +
fast_tutorial injection, distance + phase marginalisation, 1e6 prior draws per step per GPU.
A "step" is one pass of the hot path over one batch.  Prints ONE JSON line (contract in the task
description): value = device-resident whole-job evaluations/s; e2e = the same through the C ABI with
HOST buffers (pinned H2D of the parameter rows + D2H of lnL inside the timed region); roofline for the
dominant kernel (K1) from CUDA events bracketing its launches; cpu_baseline = the oracle port of the
reference's CPU path timed on the box's host cores on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DURATION = 4.0
FS = 2048.0
DETECTORS = ["H1", "L1", "V1"]
NOISE_SEED = 88170235
DRAW_SEED = 20261017
FLOP_PER_BIN = 240 + 30 * len(DETECTORS)      # SURVEY.md section 8d convention (config 2: 330 flop / active bin)
EPILOGUE_FLOP = 310
MTSUN = 6.6743e-11 * 1.988409870698050731911960804878414216e30 / 299792458.0 ** 3
METRIC = "log-likelihood evals/sec (IMRPhenomD H1L1V1, batched)"
WORKLOAD = "configs[1]: BBH 4s@2048Hz H1L1V1 IMRPhenomD, distance+phase marginalisation, 1e6-sample batches"


# ------------------------------------------------------------------------------------------------
# CPU side (oracle port of the reference path) - never touches CUDA
# ------------------------------------------------------------------------------------------------
_ORACLE = {}


def _oracle_setup(table_processes):
    from oracle import cbc_likelihood as ocl
    from bilby_b200.workloads import INJECTION
    inj = dict(INJECTION)
    start = inj["geocent_time"] - DURATION + 2
    ifos = [ocl.OracleInterferometer(n, FS, DURATION, start) for n in DETECTORS]
    rng = np.random.default_rng(NOISE_SEED)
    wa = dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0)
    conv = ocl.convert_to_lal_binary_black_hole_parameters(inj)
    for ifo in ifos:
        ifo.set_gaussian_noise(rng)
    pols = ocl.lal_binary_black_hole(ifos[0].frequency_array, *[conv[k] for k in ocl.SOURCE_ARGS], **wa)
    for ifo in ifos:
        ifo.frequency_domain_strain = ifo.frequency_domain_strain + ifo.get_detector_response(pols, conv)
    like = ocl.OracleLikelihood(ifos, waveform_arguments=wa, phase_marginalization=True,
                                distance_marginalization=True, distance_prior=ocl.OraclePowerLaw(2, 100.0, 5000.0),
                                table_processes=table_processes)
    return like


def _oracle_eval(p):
    return _ORACLE["like"].log_likelihood_ratio(p)


def cpu_reference_run(n_eval, steps, warmup, cores):
    """Times `steps` passes of n_eval evaluations over a fork pool (the reference's own fan-out,
    bilby/core/sampler/base_sampler.py:772-800).  Returns evals/s and per-step ms."""
    import multiprocessing
    from oracle import cbc_likelihood as ocl
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    _ORACLE["like"] = _oracle_setup(table_processes=cores)
    from bilby_b200.workloads import draw_bbh_prior
    draws = draw_bbh_prior(n_eval, np.random.default_rng(DRAW_SEED))
    plist = [{k: float(v[i]) for k, v in draws.items()} for i in range(n_eval)]
    ctx = multiprocessing.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        chunk = max(1, n_eval // (cores * 4))
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_oracle_eval, plist, chunksize=chunk)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    total = sum(times)
    return n_eval * len(times) / total, 1e3 * total / len(times)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def build_likelihood():
    import bilby_b200 as bb
    from bilby_b200.core.prior import PriorDict, Uniform, PowerLaw
    from bilby_b200.gw.detector import InterferometerList
    from bilby_b200.gw.source import lal_binary_black_hole
    from bilby_b200.workloads import INJECTION
    inj = dict(INJECTION)
    start = inj["geocent_time"] - DURATION + 2
    wfg = bb.gw.WaveformGenerator(
        duration=DURATION, sampling_frequency=FS, start_time=start,
        frequency_domain_source_model=lal_binary_black_hole,
        waveform_arguments=dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0,
                                minimum_frequency=20.0))
    ifos = InterferometerList(DETECTORS)
    ifos.set_strain_data_from_power_spectral_densities(FS, DURATION, start, rng=np.random.default_rng(NOISE_SEED))
    ifos.inject_signal(parameters=inj, waveform_generator=wfg)
    priors = PriorDict(dict(phase=Uniform(0, 2 * np.pi, "phase"),
                            luminosity_distance=PowerLaw(2, 100.0, 5000.0, "luminosity_distance")))
    tmp = os.path.join(tempfile.gettempdir(), f"bb200_lookup_{os.getpid()}.npz")
    like = bb.gw.GravitationalWaveTransient(ifos, wfg, phase_marginalization=True, distance_marginalization=True,
                                            priors=priors, distance_marginalization_lookup_table=tmp)
    return like


def draw_rows(like, n, seed):
    from bilby_b200.workloads import draw_bbh_prior
    draws = draw_bbh_prior(n, np.random.default_rng(seed))
    return np.ascontiguousarray(like.pack(draws))


def active_bins(rows, df, n_freq, f_min=20.0, f_max=1024.0):
    """Active (sample, bin) count per SURVEY.md section 8d: masked bins with f < min(f_max, 0.2/(M t_sun))."""
    msec = (rows[:, 0] + rows[:, 1]) * MTSUN
    fmp = np.minimum(f_max, 0.2 / msec)
    k1 = np.minimum(np.floor(fmp / df), np.floor(f_max / df) + 1)
    k0 = np.ceil(f_min / df)
    return float(np.sum(np.maximum(k1 - k0, 0)))


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.path = os.path.join(tempfile.gettempdir(), f"bb200_clocks_{os.getpid()}.csv")
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            # under-load samples: the top half of the observed clocks
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def gpu_run(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from bilby_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - bilby_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.batch
    like = build_likelihood()
    net = like.device_network
    lib = net.lib
    rows_np = draw_rows(like, n, DRAW_SEED + rank)          # every rank its own draws (weak scaling)
    rows_pinned = torch.from_numpy(rows_np).pin_memory()
    out_pinned = torch.empty(n, dtype=torch.float64).pin_memory()
    rows_dev = rows_pinned.cuda()
    out_dev = torch.empty(n, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()

    def step_device():
        _lib.check(lib.bb_log_likelihood_ratio_device(net.ptr, rows_dev.data_ptr(), n, out_dev.data_ptr(),
                                                      ctypes.c_void_p(stream.cuda_stream)))

    def step_host():
        _lib.check(lib.bb_log_likelihood_ratio_host(net.ptr, rows_pinned.data_ptr(), n, out_pinned.data_ptr()))

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

    # ---- FP64 roofline denominator (measured here: MEASURED_PEAKS.json has no FP64 figure)
    peak = ctypes.c_double(0.0)
    _lib.check(lib.bb_fp64_peak(net.ptr, ctypes.byref(peak)))

    # ---- device-resident throughput
    for _ in range(args.warmup):
        step_device()
    barrier()
    _lib.check(lib.bb_profile_enable(net.ptr, 1))
    launches0 = lib.bb_launch_count(net.ptr)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.bb_launch_count(net.ptr) - launches0
    k1_ms, k1_n = ctypes.c_double(0.0), ctypes.c_long(0)
    _lib.check(lib.bb_profile_read(net.ptr, ctypes.byref(k1_ms), ctypes.byref(k1_n)))
    _lib.check(lib.bb_profile_enable(net.ptr, 0))

    # ---- end to end through the C ABI with host buffers
    for _ in range(max(1, min(args.warmup, 2))):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    clock_info = clocks.stop() if clocks else None
    lnl_sum = float(out_pinned.sum())

    value = world * n * args.steps / (ms_total * 1e-3)
    e2e_value = world * n * args.steps / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (K1)
    df = net.ifos[0].frequency_array[1]
    bins = active_bins(rows_np, df, net.n_freq)
    k1_avg_ms = k1_ms.value / max(1, k1_n.value)
    achieved = bins * FLOP_PER_BIN / (k1_avg_ms * 1e-3) / 1e12 if k1_avg_ms > 0 else 0.0
    traffic, capture = None, {}
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        try:
            capture = json.load(open(tpath))
            traffic = capture.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roofline = dict(bound="fp64", achieved=achieved, peak=peak.value, unit="TFLOP/s",
                    frac=achieved / peak.value if peak.value else None, traffic=traffic,
                    traffic_source="profiles/k1_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one "
                                   "`ncu --set full` capture of this kernel, scaled to this batch (not measured in this run)",
                    kernel="bb_inner_product_kernel<3>", kernel_ms=k1_avg_ms,
                    kernel_share_of_step=k1_ms.value / (e0.elapsed_time(e1)) if k1_n.value else None,
                    algorithmic_flop_per_launch=bins * FLOP_PER_BIN, active_bins_per_eval=bins / n,
                    peak_source="in-run DFMA stream kernel (bb_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                    hbm_peak_gbs=peaks.get("hbm_gbs"),
                    algorithmic_hbm_bytes_per_launch=n * (16 + 1) * 8,
                    # `achieved` counts SURVEY section 8d's convention (330 flop per bin: sincos = 60, ...); the kernel
                    # executes about half as many FP64 operations (recurrences, tile columns, a shared sincospi), so
                    # frac can exceed 1 while the hardware counter of the same capture shows the pipe half idle
                    executed=dict(sm__pipe_fp64_cycles_active_pct=capture.get("sm__pipe_fp64_cycles_active_pct"),
                                  smsp__issue_active_pct=capture.get("smsp__issue_active_pct"),
                                  source=capture.get("source")))

    # ---- what a stock bilby sampler sees: one parameter dict per call (bilby/core/sampler/base_sampler.py:538-563)
    scalar_us = None
    if rank == 0:
        from bilby_b200.workloads import draw_bbh_prior
        d1 = draw_bbh_prior(256, np.random.default_rng(DRAW_SEED))
        dicts = [{k: float(v[i]) for k, v in d1.items()} for i in range(256)]
        for p in dicts[:16]:
            like.log_likelihood_ratio(p)
        t0 = time.perf_counter()
        for p in dicts:
            like.log_likelihood_ratio(p)
        scalar_us = (time.perf_counter() - t0) / len(dicts) * 1e6

    # ---- the other BASELINE.json configurations (and, for N > 1, the frequency-sharded long signal) in the same run
    del like, net, rows_dev, out_dev
    torch.cuda.empty_cache()
    extra = None
    if not args.no_extra:
        extra = run_extra(args, world, rank, local_rank, dist if world > 1 else None)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on the host cores (subprocess: fork pools and CUDA do not mix)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_eval = args.cpu_evals if args.cpu_evals else 250 * cores
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
               "--cpu-evals", str(n_eval), "--quiet-json"]
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
            line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
            ref = json.loads(line)
            cpu = ref["cpu_baseline"]
        except Exception as exc:   # pragma: no cover
            cpu = dict(value=None, unit="evals/s", cores=cores, kind="port", sample=f"failed: {exc}")

    line = dict(metric=METRIC, value=value, unit="evals/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD, batch_per_gpu=n, detectors=DETECTORS, duration_s=DURATION,
                            sampling_frequency_hz=FS, marginalisation="distance+phase",
                            partition=f"samples x{world} (no data-path collective)",
                            l2="inputs+scratch per step (128 MB rows + 576 MB coefficient records) exceed the 126 MB L2"),
                clocks=clock_info,
                e2e=dict(value=e2e_value, unit="evals/s", h2d_bytes_per_step=n * 16 * 8, d2h_bytes_per_step=n * 8,
                         ms_per_step=e2e_ms / args.steps, api="bb_log_likelihood_ratio_host (C ABI, pinned host buffers)"),
                gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu, checksum_lnl=lnl_sum,
                scalar_call_us=scalar_us, extra=extra)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


EXTRA_CONFIGS = ("cfg0", "cfg2", "cfg3", "cfg4_relbin", "cfg4_roq", "cfg4_roq_time")


def run_extra(args, world, rank, local_rank, dist):
    """every BASELINE.json configuration under the driver's eyes.  Same JSON line, key `extra`:
    per configuration the device-resident value, the end-to-end value through the host entry point, the roofline of
    its dominant kernel and (N = 1) the CPU baseline from the oracle port on a bounded sample; for N > 1 also
    configs[3] with the frequency axis sharded over the N GPUs (strong scaling), through the fused peer-memory exchange
    and through one NCCL all-reduce."""
    import bench_configs as bc
    out = dict(note="steps=5, warmup=3 per configuration; value = device-resident, e2e = host buffers through the C ABI; "
                    "weak scaling over samples (every rank its own batch) unless stated", configs={})
    for cfg in EXTRA_CONFIGS:
        try:
            line = bc.run_config(cfg, bc.DEFAULT_BATCH[cfg], 5, 3, world, rank, local_rank, dist, clocks=False)
        except Exception as exc:      # pragma: no cover - reported, never hidden
            line = dict(error=repr(exc)) if rank == 0 else None
        if rank != 0:
            continue
        if "error" not in line:
            r = line["roofline"]
            line = dict(workload=line["config"]["workload"], batch_per_gpu=line["config"]["batch"], value=line["value"],
                        unit=line["unit"], ms_per_step=line["ms_per_step"],
                        e2e=dict(value=line["e2e"]["value"], h2d_bytes_per_step=line["e2e"]["h2d_bytes_per_step"],
                                 d2h_bytes_per_step=line["e2e"]["d2h_bytes_per_step"]),
                        gpu_launches=line["gpu_launches"], checksum_lnl=line["checksum_lnl"],
                        roofline=dict(bound=r["bound"], achieved=r["achieved"], peak=r["peak"], unit=r["unit"], frac=r["frac"],
                                      kernel=r["kernel"], kernel_share_of_step=r["kernel_share_of_step"],
                                      peak_source=r["peak_source"]),
                        **({"device_front_end": line["device_front_end"]} if "device_front_end" in line else {}),
                        **{k: line["config"][k] for k in ("n_time_contracted", "executed_gemm_flop_per_eval") if k in line["config"]})
            if world == 1 and not args.no_cpu_baseline:
                cmd = [sys.executable, os.path.join(ROOT, "bench_cpu.py"), "--config", cfg]
                try:
                    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
                    cpu = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
                    line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
                except Exception as exc:      # pragma: no cover
                    line["cpu_baseline"] = dict(value=None, kind="port", sample=f"failed: {exc}")
        out["configs"][cfg] = line
    if world > 1:
        shard = {}
        for exchange in ("fused", "nccl"):
            try:
                shard[exchange] = bc.run_frequency_sharded(8192, 3, 3, world, rank, dist, exchange=exchange)
            except Exception as exc:      # pragma: no cover
                shard[exchange] = dict(error=repr(exc))
        out["frequency_sharded_configs3"] = shard
    return out if rank == 0 else None


def reference_run(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_eval = args.cpu_evals if args.cpu_evals else 250 * cores
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    value, ms = cpu_reference_run(n_eval, max(1, args.steps), args.warmup, cores)
    sample = (f"{n_eval} prior draws per step (rng {DRAW_SEED}) of the 1e6-draw workload, fork pool of {cores} "
              f"processes, OMP_NUM_THREADS=1; waveform = oracle restatement of IMRPhenomD (lalsimulation absent)")
    cpu = dict(value=value, unit="evals/s", cores=cores, kind="port", sample=sample)
    line = dict(impl="reference", metric=METRIC, value=value, unit="evals/s", n_gpus=world, steps=max(1, args.steps),
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD, batch_per_gpu=n_eval, detectors=DETECTORS, duration_s=DURATION,
                            sampling_frequency_hz=FS, marginalisation="distance+phase"),
                cpu_baseline=cpu,
                e2e=dict(value=value, unit="evals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1_000_000)
    ap.add_argument("--cpu-evals", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quiet-json", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other configurations (headline line only)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_run(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        gpu_run(args)


if __name__ == "__main__":
    main()
