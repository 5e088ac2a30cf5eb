// Device builder for the linear ROQ weights (included by bb_kernels.cu).
//
// Replaces the loop of ROQGravitationalWaveTransient._set_weights_linear (bilby/gw/likelihood/roq.py:849-918): for every
// basis element b and detector the reference zero-pads data * conj(basis_b) / PSD to n_time points, takes one inverse FFT
// and keeps the time samples [lo, hi] around the coalescence-time prior.  Only those n_win << n_time samples are wanted,
// so the transform is evaluated directly as a dense contraction over the basis frequencies,
//
//   lw[t, b] = (4 / T) sum_j E[t, j] G[j, b],   E[t, j] = exp(+2 pi i k_j (lo + t) / n_time),  G[j, b] = (d/S)_j conj(B[b, j]),
//
// in slabs of frequencies: the phase matrix of a slab is generated on the device (phases reduced modulo n_time in integer
// arithmetic, so they are exact) and multiplied with the DMMA GEMM of bb_gemm.cuh (FP64 tensor path), accumulating over
// the slabs.
#pragma once

#define BB_RW_SLAB 8192

// Gt[b][j] = (d/S)_j conj(B[b][j]) in the GEMM's packed layout (bb_pk, B operand: tiles of 128 basis elements x slabs
// of 16 frequencies; the buffer is zero filled by the caller)
__global__ void bb_rw_g_kernel(const double2* __restrict__ d_over_s, const double2* __restrict__ basis /* [nb][n] */,
                               int n, int nb, long S, double2* __restrict__ Gt) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)n * nb) return;
    const int b = (int)(i / n), j = (int)(i - (long)b * n);
    const double2 x = d_over_s[j], y = basis[i];
    Gt[bb_pk(b, j, BB_GEMM_TR_B, S)] = make_double2(x.x * y.x + x.y * y.y, x.y * y.x - x.x * y.y);        // x conj(y)
}

// phase matrix of one slab of frequencies, packed as the GEMM's A operand (tiles of 64 times); ld = nj rounded up to 16
__global__ void bb_rw_phase_kernel(const int* __restrict__ kj, int j0, int nj, int ld, long lo, int n_win, long n_time,
                                   int pow2, double2* __restrict__ E /* zero for nj <= j < ld */) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)n_win * ld) return;
    const int t = (int)(i / ld), jj = (int)(i - (long)t * ld);
    const size_t o = bb_pk(t, jj, BB_GEMM_TR_A(true), ld >> 4);
    if (jj >= nj) { E[o] = make_double2(0.0, 0.0); return; }
    const long full = (long)kj[j0 + jj] * (lo + t);
    const long prod = pow2 ? (full & (n_time - 1)) : (full % n_time);          // exact phase index
    double sn, cs;
    sincospi(2.0 * (double)prod / (double)n_time, &sn, &cs);
    E[o] = make_double2(cs, sn);
}

extern "C" int bb_build_roq_linear_weights(int device, int n_det, int n_freq_sel, const double* d_over_s, int n_basis,
                                           const double* basis, const int* bin_index, long n_time, long lo, int n_win,
                                           double duration, double* out) {
    if (n_det < 1 || n_freq_sel < 1 || n_basis < 1 || n_win < 1 || n_time < 2 || !d_over_s || !basis || !bin_index || !out)
        return bb_fail("bb_build_roq_linear_weights: bad arguments");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        return bb_fail("bb_build_roq_linear_weights: no such CUDA device (bilby_b200 has no CPU path)");
    BB_CUDA(cudaSetDevice(device));
    double2 *g_dos = nullptr, *g_basis = nullptr, *g_G = nullptr, *g_E = nullptr, *g_out = nullptr;
    int* g_k = nullptr;
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(g_dos); cudaFree(g_basis); cudaFree(g_G); cudaFree(g_E); cudaFree(g_out); cudaFree(g_k);
    };
#define BB_RW_TRY(call)                                                                                      \
    do {                                                                                                     \
        if ((call) != cudaSuccess) { rc = bb_fail(std::string(#call) + ": " + cudaGetErrorString(cudaGetLastError())); cleanup(); return rc; } \
    } while (0)
    const size_t n = (size_t)n_freq_sel;
    BB_RW_TRY(cudaMalloc(&g_dos, n * n_det * sizeof(double2)));
    BB_RW_TRY(cudaMalloc(&g_basis, n * n_basis * sizeof(double2)));
    const long S = (long)((n + 15) / 16);
    const size_t g_stride = bb_pk_elems(n_basis, (long)n, BB_GEMM_TR_B);
    BB_RW_TRY(cudaMalloc(&g_G, g_stride * n_det * sizeof(double2)));
    BB_RW_TRY(cudaMemset(g_G, 0, g_stride * n_det * sizeof(double2)));
    BB_RW_TRY(cudaMalloc(&g_k, n * sizeof(int)));
    const int slab = (int)(S * 16) < BB_RW_SLAB ? (int)(S * 16) : BB_RW_SLAB;      // multiple of 16
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device);
    BB_RW_TRY(cudaMalloc(&g_E, bb_pk_elems(n_win, slab, BB_GEMM_TR_A(true)) * sizeof(double2)));
    BB_RW_TRY(cudaMemset(g_E, 0, bb_pk_elems(n_win, slab, BB_GEMM_TR_A(true)) * sizeof(double2)));
    BB_RW_TRY(cudaMalloc(&g_out, (size_t)n_det * n_win * n_basis * sizeof(double2)));
    BB_RW_TRY(cudaMemcpy(g_dos, d_over_s, n * n_det * sizeof(double2), cudaMemcpyHostToDevice));
    BB_RW_TRY(cudaMemcpy(g_basis, basis, n * n_basis * sizeof(double2), cudaMemcpyHostToDevice));
    BB_RW_TRY(cudaMemcpy(g_k, bin_index, n * sizeof(int), cudaMemcpyHostToDevice));
    const bool pow2 = (n_time & (n_time - 1)) == 0;
    for (int det = 0; det < n_det; ++det) {
        const long totg = (long)n * n_basis;
        bb_rw_g_kernel<<<(unsigned)((totg + 255) / 256), 256>>>(g_dos + (size_t)det * n, g_basis, n_freq_sel, n_basis,
                                                                 S, g_G + (size_t)det * g_stride);
    }
    // the phase matrix of a slab is the same for every detector: one batched GEMM per slab
    for (int j0 = 0; j0 < n_freq_sel; j0 += slab) {
        const int nj = (n_freq_sel - j0) < slab ? (n_freq_sel - j0) : slab;
        const int njp = (nj + 15) & ~15;
        const long tot = (long)n_win * njp;
        bb_rw_phase_kernel<<<(unsigned)((tot + 255) / 256), 256>>>(g_k, j0, nj, njp, lo, n_win, n_time, pow2 ? 1 : 0, g_E);
        // out_det[t][b] (+)= (4 / T) sum_j E[t][j] Gt_det[b][j0 + j]
        BBGemmArgs ga{};
        ga.A[0] = g_E;
        ga.B[0] = g_G + (size_t)(j0 >> 4) * BB_GEMM_TR_B * BB_PK;      // slab j0 / 16 of every row tile
        ga.C = reinterpret_cast<double*>(g_out);
        ga.slabs_a = njp >> 4; ga.slabs_b = S;
        ga.slab0 = 0; ga.n_slabs = njp >> 4;
        ga.ldc = n_basis;
        ga.batch_a = 0; ga.batch_b = (long)g_stride; ga.batch_c = (long)n_win * n_basis;
        ga.M = n_win; ga.N = n_basis; ga.n_seg = 1; ga.n_batch = n_det;
        ga.accumulate = j0 ? 1 : 0;
        ga.alpha = 4.0 / duration;
        if (bb_gemm_nt(true, ga, sm_count, nullptr)) { cleanup(); return 1; }
    }
    BB_RW_TRY(cudaMemcpy(out, g_out, (size_t)n_det * n_win * n_basis * sizeof(double2), cudaMemcpyDeviceToHost));
    BB_RW_TRY(cudaDeviceSynchronize());
#undef BB_RW_TRY
    cleanup();
    return 0;
}
