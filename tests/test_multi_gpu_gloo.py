"""N > 1 host logic on CPU (gloo, world_size 2): sample sharding needs no collective; frequency sharding
sums rank-local partial inner products with one all-reduce.  The partial sums come
from the oracle here (no GPU in the build container); the GPU path uses the same partition helpers."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bilby_b200.parallel import frequency_shards, shard_range


def test_partition_helpers():
    for n, w in ((10, 3), (1_000_000, 8), (7, 8), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
    sh = frequency_shards(2560, 262144, 8, 262145)
    assert sh[0][0] == 0 and sh[-1][1] == 262145
    assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cbc_likelihood as ocl
    from bilby_b200.parallel import allreduce_inner_products
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bbh_4s_zero_H1L1.npz"))
    st = float(g["start_time"])
    ifos = [ocl.OracleInterferometer(n, 2048.0, 4.0, st) for n in ("H1", "L1")]
    for ifo in ifos:
        ifo.frequency_domain_strain = g[f"strain_{ifo.name}"]
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_") and k != "param_time_jitter"}
    wa = dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0)
    like = ocl.OracleLikelihood(ifos, waveform_arguments=wa)
    # ---- frequency sharding: each rank owns a contiguous bin range of the mask
    shards = frequency_shards(80, 4096, world, 4097)
    b, e = shards[rank]
    for ifo in ifos:
        m = np.zeros(4097, dtype=bool)
        m[b:e] = True
        ifo.frequency_mask = ifo.frequency_mask & m
    n = 6
    part = np.zeros((n, 2, 3))
    for i in range(n):
        per_det = like.log_likelihood_ratio({k: float(v[i]) for k, v in draws.items()}, return_snrs=True)
        for d, (dh, hh) in enumerate(per_det):
            part[i, d] = [dh.real, dh.imag, hh]
    t = torch.from_numpy(part)
    allreduce_inner_products(t)
    full = t.numpy()
    lnl = full[..., 0].sum(axis=1) - full[..., 2].sum(axis=1) / 2
    assert np.allclose(lnl, g["lnl_none"][:n], rtol=1e-10, atol=1e-10)
    # ---- sample sharding: no collective, gather only to compare
    lo, hi = shard_range(8, rank, world)
    mine = torch.tensor(g["lnl_none"][lo:hi])
    out = [torch.empty(shard_range(8, r, world)[1] - shard_range(8, r, world)[0], dtype=torch.float64)
           for r in range(world)]
    dist.all_gather(out, mine)
    assert np.array_equal(torch.cat(out).numpy(), g["lnl_none"][:8])
    dist.destroy_process_group()


def test_world_size_two_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
