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
    # the same with the exchange fused into the kernels (peer-memory stores + flag round, csrc/bb_exchange.cuh);
    # several back-to-back calls exercise both buffer parities
    fused = FrequencyShardedLikelihood(like, rank, world, fused_max_rows=512)
    assert fused.fused
    for it in range(5):
        sub = rows[: 256 - 16 * it]
        got_f = fused.log_likelihood_ratio_rows(sub).cpu().numpy()
        err_f = np.max(np.abs(got_f - full[: 256 - 16 * it])) / np.max(np.abs(full))
        assert err_f < 1e-10, (it, err_f)
    fused.check_exchange()
    phase = bb.gw.GravitationalWaveTransient(
        ifos, wfg, phase_marginalization=True,
        priors=bb.core.prior.PriorDict(dict(phase=bb.core.prior.Uniform(0, 2 * np.pi, "phase"))))
    full_p = phase.log_likelihood_ratio_batch(rows).cpu().numpy()
    fused_p = FrequencyShardedLikelihood(phase, rank, world, fused_max_rows=256)
    got_p = fused_p.log_likelihood_ratio_rows(rows).cpu().numpy()
    assert np.max(np.abs(got_p - full_p)) / np.max(np.abs(full_p)) < 1e-10
    fused_p.check_exchange()
    dist.barrier()
    if rank == 0:
        print("FREQ_SHARD_OK", world, err, "fused", err_f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
