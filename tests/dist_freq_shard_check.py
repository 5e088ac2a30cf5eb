"""Run under torchrun with N >= 2 ranks (one GPU each): the frequency-sharded TaylorF2 likelihood with one
NCCL all-reduce of partial inner products equals the unsharded evaluation (tests/test_gpu_bns.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.dirname(__file__))


def main():
    import bilby_b200 as bb
    from bilby_b200.parallel import FrequencyShardedLikelihood
    from test_gpu_bns import build_pair, bns_draws
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    _, ifos, wfg = build_pair()
    like = bb.gw.GravitationalWaveTransient(ifos, wfg)
    draws = bns_draws(256, np.random.default_rng(5))
    rows = torch.from_numpy(like.pack(draws)).cuda()
    full = like.log_likelihood_ratio_batch(rows).cpu().numpy()
    sharded = FrequencyShardedLikelihood(like, rank, world)
    got = sharded.log_likelihood_ratio_rows(rows).cpu().numpy()
    err = np.max(np.abs(got - full)) / np.max(np.abs(full))
    assert err < 1e-10, err
    dist.barrier()
    if rank == 0:
        print("FREQ_SHARD_OK", world, err)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
