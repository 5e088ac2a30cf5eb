// Frequency-sharded likelihood with the exchange of partial inner products fused into the kernels
// (included by bb_kernels.cu).  One process per GPU; every rank owns one contiguous bin range
// (bb_set_frequency_shard) and evaluates ALL samples on it.
//
//   K1 (bb_inner_product_kernel, BBPush)   each warp stores its sample's 3 x n_det partial sums straight into this
//                                          rank's slot of EVERY rank's exchange buffer: peer memory over
//                                          NVLink / NVSwitch, sample by sample, under the arithmetic of the other warps
//   bb_exchange_signal_kernel              one warp: system-scope release of "epoch" into every peer's flag word,
//                                          then acquire-spin on the own flags until every peer has arrived
//   bb_exchange_epilogue_kernel            sums the `world` partials per sample and runs the usual point likelihood
//
// No NCCL call, no host synchronisation and no separate reduction pass over the data: the all-reduce of
// bilby_b200/parallel.py::allreduce_inner_products becomes `world` posted stores per result and one flag round.
// Buffers are plain cudaMalloc memory shared between the processes with CUDA IPC handles (exchanged by the host
// code over torch.distributed); two parities of the data area make back-to-back calls safe (a rank can run at most
// one call ahead of its slowest peer because it needs that peer's flag of the previous call).
#pragma once

static void bb_exchange_release(bb_handle* h) {
    if (!h->xc_local) return;
    for (int r = 0; r < h->xc_world; ++r)
        if (r != h->xc_rank && h->xc_peer[r]) cudaIpcCloseMemHandle(h->xc_peer[r]);
    cudaFree(h->xc_local);
    cudaFree(h->xc_error);
    h->xc_local = nullptr;
    h->xc_error = nullptr;
    for (int r = 0; r < 8; ++r) h->xc_peer[r] = nullptr;
    h->xc_world = 0;
    h->xc_cap = 0;
    h->xc_epoch = 0;
}

extern "C" int bb_exchange_create(bb_handle* h, int world, int rank, long max_rows, void* ipc_handle_out) {
    if (!h || !h->have_network) return bb_fail("bb_exchange_create: network not set");
    if (world < 1 || world > BB_MAX_RANKS || rank < 0 || rank >= world || max_rows < 1 || !ipc_handle_out)
        return bb_fail("bb_exchange_create: bad arguments (1 <= world <= 8)");
    BB_CUDA(cudaSetDevice(h->device));
    bb_exchange_release(h);
    const size_t bytes = BB_XC_HEADER + (size_t)2 * world * max_rows * h->net.n_det * 3 * sizeof(double);
    BB_CUDA(cudaMalloc(&h->xc_local, bytes));
    BB_CUDA(cudaMemset(h->xc_local, 0, bytes));
    BB_CUDA(cudaMalloc(&h->xc_error, sizeof(int)));
    BB_CUDA(cudaMemset(h->xc_error, 0, sizeof(int)));
    BB_CUDA(cudaDeviceSynchronize());
    h->xc_world = world;
    h->xc_rank = rank;
    h->xc_cap = max_rows;
    h->xc_epoch = 0;
    cudaIpcMemHandle_t ipc;
    BB_CUDA(cudaIpcGetMemHandle(&ipc, h->xc_local));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(ipc_handle_out, &ipc, sizeof(ipc));
    return 0;
}

extern "C" int bb_exchange_connect(bb_handle* h, const void* ipc_handles) {
    if (!h || !h->xc_local || !ipc_handles) return bb_fail("bb_exchange_connect: bb_exchange_create first");
    BB_CUDA(cudaSetDevice(h->device));
    for (int r = 0; r < h->xc_world; ++r) {
        if (r == h->xc_rank) {
            h->xc_peer[r] = h->xc_local;
            continue;
        }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, static_cast<const char*>(ipc_handles) + (size_t)r * sizeof(ipc), sizeof(ipc));
        BB_CUDA(cudaIpcOpenMemHandle(&h->xc_peer[r], ipc, cudaIpcMemLazyEnablePeerAccess));
    }
    return 0;
}

extern "C" int bb_exchange_destroy(bb_handle* h) {
    if (!h) return bb_fail("bb_exchange_destroy: null handle");
    cudaSetDevice(h->device);
    bb_exchange_release(h);
    return 0;
}

struct BBPeerFlags {
    unsigned long long* flags[BB_MAX_RANKS];   // flags[r]: the flag array at the head of rank r's buffer
};

// One warp.  Lane r < world releases `epoch` into word `rank` of rank r's flags (the kernel boundary after K1 plus
// the system fence order K1's peer stores before it), then waits until word r of the own flags reaches `epoch`.
__global__ void bb_exchange_signal_kernel(BBPeerFlags pf, int world, int rank, unsigned long long epoch,
                                          long long timeout_cycles, int* error) {
    const int r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    unsigned long long* remote = pf.flags[r] + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(epoch) : "memory");
    const unsigned long long* mine = pf.flags[rank] + r;
    const long long t0 = clock64();
    unsigned long long seen = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
        if (seen >= epoch) break;
        if (clock64() - t0 > timeout_cycles) {       // a peer never arrived: report instead of hanging the GPU
            atomicExch(error, 1 + r);
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

__global__ void bb_exchange_epilogue_kernel(const double* __restrict__ coef, const double* __restrict__ partials,
                                            long slot, int world, long n, int n_det, BBMarg marg,
                                            double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* c = coef + i * BC_NCOEF;
    if (c[BC_STATUS] != 0.0) { out[i] = -DBL_MAX; return; }
    double dre = 0.0, dim = 0.0, hh = 0.0;
    for (int r = 0; r < world; ++r) {
        const double* s = partials + (size_t)r * slot + (size_t)i * n_det * 3;
        for (int d = 0; d < n_det; ++d) {
            dre += __ldcg(s + 3 * d);             // written by the peers: read at L2, never through L1
            dim += __ldcg(s + 3 * d + 1);
            hh += __ldcg(s + 3 * d + 2);
        }
    }
    out[i] = bb_point_lnl(marg, dre, dim, hh, c[BC_DISTANCE]);
}

extern "C" int bb_log_likelihood_ratio_sharded_device(bb_handle* h, const double* params_dev, long n, double* out_dev,
                                                      void* stream) {
    if (!h || !h->have_network) return bb_fail("bb_log_likelihood_ratio_sharded_device: network not set");
    if (!h->xc_local || !h->xc_peer[(h->xc_rank + 1) % h->xc_world])
        return bb_fail("bb_log_likelihood_ratio_sharded_device: bb_exchange_create / bb_exchange_connect first");
    if (n <= 0) return 0;
    if (n > h->xc_cap) return bb_fail("bb_log_likelihood_ratio_sharded_device: batch larger than the exchange buffers");
    if (h->kind != 0 || h->cm_n_curves > 0 || (h->marg.flags & BB_MARG_TIME))
        return bb_fail("frequency sharding applies to the full-grid likelihood without time / calibration marginalisation");
    BB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (bb_ensure_scratch(h, (size_t)n)) return 1;
    if (bb_launch_prologue(h, params_dev, n, st)) return 1;
    h->xc_epoch++;
    h->xc_push_next = true;
    const int rc = bb_launch_inner(h, n, h->d_snr, st);
    h->xc_push_next = false;
    if (rc) return 1;
    BBPeerFlags pf;
    for (int r = 0; r < BB_MAX_RANKS; ++r)
        pf.flags[r] = r < h->xc_world ? static_cast<unsigned long long*>(h->xc_peer[r]) : nullptr;
    bb_exchange_signal_kernel<<<1, 32, 0, st>>>(pf, h->xc_world, h->xc_rank, h->xc_epoch, 20000000000LL, h->xc_error);
    h->launches++;
    BB_CUDA(cudaGetLastError());
    const size_t slot = (size_t)h->xc_cap * h->net.n_det * 3;
    const double* partials = reinterpret_cast<const double*>(static_cast<const char*>(h->xc_local) + BB_XC_HEADER)
                             + (size_t)(h->xc_epoch & 1ull) * h->xc_world * slot;
    const int threads = 128;
    bb_exchange_epilogue_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(
        h->d_coef, partials, (long)slot, h->xc_world, n, h->net.n_det, h->marg, out_dev);
    h->launches++;
    BB_CUDA(cudaGetLastError());
    return 0;
}

// 0 when every wait so far has completed; 1 + r when the flag of rank r never arrived (checked by the host wrapper)
extern "C" int bb_exchange_status(bb_handle* h, int* status_out) {
    if (!h || !h->xc_error || !status_out) return bb_fail("bb_exchange_status: no exchange");
    BB_CUDA(cudaSetDevice(h->device));
    BB_CUDA(cudaMemcpy(status_out, h->xc_error, sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}
