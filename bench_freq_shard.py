#!/usr/bin/env python
"""configs[3] with the frequency axis sharded over N GPUs: BNS TaylorF2 + tides, 128 s at
4096 Hz, H1L1V1.  Every rank evaluates ALL samples on its contiguous bin range, one NCCL all-reduce of the partial
inner products ([n, n_det, 3] float64), replicated epilogue.  Strong scaling: the total work is fixed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench_freq_shard.py [--batch 8192] [--steps 5]
    python bench_freq_shard.py            # N = 1 (no collective)

Prints ONE JSON line on rank 0: whole-job evaluations/s (max over ranks of the CUDA-event time), the time of the
all-reduce alone, and the agreement with the unsharded evaluation of the same rows.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="fused: K1 stores partials into peer memory + flag round (csrc/bb_exchange.cuh); "
                         "nccl: one NCCL all-reduce")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench_configs as bc
    from bilby_b200.parallel import FrequencyShardedLikelihood
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    like, rows_np, _, flop, desc = bc.build("cfg3", args.batch)
    rows = torch.from_numpy(np.ascontiguousarray(rows_np)).cuda()
    check = like.log_likelihood_ratio_batch(rows[:256]).cpu().numpy()          # unsharded, before the shard is set
    sharded = FrequencyShardedLikelihood(like, rank, world, fused_max_rows=args.batch if args.exchange == "fused" else 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        out = sharded.log_likelihood_ratio_rows(rows)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = sharded.log_likelihood_ratio_rows(rows)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    # the exchange alone
    snrs = like.inner_products_batch(rows)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(args.steps):
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
    if rank == 0:
        total_flop, _ = flop(rows_np)
        t = float(ms.item()) * 1e-3 / args.steps
        print(json.dumps(dict(
            metric="log-likelihood evals/sec (TaylorF2+tides 128 s H1L1V1, frequency-sharded)", value=args.batch / t,
            unit="evals/s", n_gpus=world, steps=args.steps, warmup=max(3, args.warmup), ms_per_step=t * 1e3,
            scaling="strong", dtype="f64", data="synthetic",
            config=dict(workload=desc["workload"], batch=args.batch,
                        partition=f"frequency axis in {world} contiguous shards, bins [{sharded.k_begin}, {sharded.k_end}) "
                                  f"on rank 0; exchange of {args.batch * 3 * 3 * 8} bytes per rank per step",
                        exchange=("fused into K1 (peer-memory stores over NVLink + flag round)" if sharded.fused
                                  else "NCCL all-reduce")),
            allreduce_ms_per_step=float(ar.item()) / args.steps,
            algorithmic_tflops=total_flop / t / 1e12, max_rel_diff_vs_unsharded=err)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
