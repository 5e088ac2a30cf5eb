"""Multi-GPU partitioning of the hot path.  One process per GPU.

* samples  : independent units -> contiguous slices per rank, NO collective on the data path.
* frequency: (long BNS signals) contiguous bin ranges per rank; every rank evaluates all samples on its
             slice, one all-reduce(SUM, float64, [n, n_det, 3]) of the partial inner products, then the
             replicated epilogue.
The reference has no equivalent (it only fans single evaluations out over a multiprocessing.Pool,
bilby/core/sampler/base_sampler.py:772-800).
"""
import numpy as np


def shard_range(n, rank, world_size):
    """Contiguous, balanced [begin, end) of n units for this rank."""
    base, rem = divmod(int(n), int(world_size))
    begin = rank * base + min(rank, rem)
    end = begin + base + (1 if rank < rem else 0)
    return begin, end


def frequency_shards(k_lo, k_hi, world_size, n_freq):
    """Split the masked bin range [k_lo, k_hi] into world_size contiguous shards [begin, end) covering
    [0, n_freq): the first shard starts at 0 and the last ends at n_freq so every bin has one owner."""
    edges = [shard_range(k_hi + 1 - k_lo, r, world_size) for r in range(world_size)]
    out = []
    for r, (b, e) in enumerate(edges):
        begin = 0 if r == 0 else k_lo + b
        end = n_freq if r == world_size - 1 else k_lo + e
        out.append((begin, end))
    return out


def allreduce_inner_products(snrs):
    """In-place SUM all-reduce of rank-local partial inner products (torch tensor, NCCL on GPUs, gloo in
    the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(snrs, op=dist.ReduceOp.SUM)
    return snrs


class FrequencyShardedLikelihood:
    """Wraps a GravitationalWaveTransient so that each rank owns one contiguous bin range.

    fused_max_rows > 0 (GPUs of one node, world_size <= 8): the exchange is fused into the kernels - K1 stores its
    partial sums into every rank's buffer over NVLink peer memory, a flag round replaces the collective, and the
    epilogue sums the partials (csrc/bb_exchange.cuh).  torch.distributed is used once, to pass the 64-byte CUDA IPC
    handles around.  Otherwise: one NCCL / gloo all-reduce of the partial inner products."""

    def __init__(self, likelihood, rank, world_size, fused_max_rows=0):
        from . import _lib
        self.likelihood = likelihood
        self.rank, self.world_size = rank, world_size
        net = likelihood.device_network
        masks = np.array([ifo.frequency_mask for ifo in likelihood.interferometers])
        idx = np.where(masks.any(axis=0))[0]
        shards = frequency_shards(int(idx[0]), int(idx[-1]), world_size, net.n_freq)
        self.k_begin, self.k_end = shards[rank]
        _lib.check(net.lib.bb_set_frequency_shard(net.ptr, self.k_begin, self.k_end))
        self.fused = False
        if fused_max_rows > 0 and world_size > 1:
            self._connect(net, int(fused_max_rows))

    def _connect(self, net, max_rows):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _lib
        mine = (ctypes.c_ubyte * 64)()
        _lib.check(net.lib.bb_exchange_create(net.ptr, self.world_size, self.rank, max_rows, mine))
        local = torch.tensor(list(mine), dtype=torch.uint8, device=net.device)
        gathered = [torch.empty_like(local) for _ in range(self.world_size)]
        dist.all_gather(gathered, local)
        handles = torch.stack(gathered).cpu().numpy().tobytes()
        _lib.check(net.lib.bb_exchange_connect(net.ptr, handles))
        dist.barrier()
        self.fused = True

    def log_likelihood_ratio_rows(self, rows):
        if self.fused:
            import ctypes
            from . import _lib
            net = self.likelihood.device_network
            torch = net.torch
            out = torch.empty(rows.shape[0], dtype=torch.float64, device=rows.device)
            _lib.check(net.lib.bb_log_likelihood_ratio_sharded_device(net.ptr, rows.data_ptr(), rows.shape[0],
                                                                      out.data_ptr(), net._stream()))
            return out
        snrs = self.likelihood.inner_products_batch(rows)
        allreduce_inner_products(snrs)
        return self.likelihood.likelihood_from_inner_products(rows, snrs)

    def check_exchange(self):
        """Raises if a peer's arrival flag was ever missed (the device-side wait gives up after ~10 s)."""
        if self.fused:
            import ctypes
            from . import _lib
            net = self.likelihood.device_network
            status = ctypes.c_int(0)
            _lib.check(net.lib.bb_exchange_status(net.ptr, ctypes.byref(status)))
            if status.value:
                raise _lib.BilbyB200Error(f"frequency-shard exchange: rank {status.value - 1} never arrived")
