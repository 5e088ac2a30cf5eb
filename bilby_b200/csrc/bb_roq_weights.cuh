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
// arithmetic, so they are exact) and multiplied with cuBLAS ZGEMM (FP64 tensor path), accumulating over the slabs.
#pragma once

#define BB_RW_SLAB 8192

__global__ void bb_rw_g_kernel(const double2* __restrict__ d_over_s, const double2* __restrict__ basis /* [nb][n] */,
                               int n, int nb, double2* __restrict__ G /* [n][nb] */) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)n * nb) return;
    const int j = (int)(i / nb), b = (int)(i - (long)j * nb);
    const double2 x = d_over_s[j], y = basis[(size_t)b * n + j];
    G[i] = make_double2(x.x * y.x + x.y * y.y, x.y * y.x - x.x * y.y);        // x conj(y)
}

__global__ void bb_rw_phase_kernel(const int* __restrict__ kj, int j0, int nj, long lo, int n_win, long n_time, int pow2,
                                   double2* __restrict__ E /* [n_win][nj] */) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)n_win * nj) return;
    const int t = (int)(i / nj), jj = (int)(i - (long)t * nj);
    const long full = (long)kj[j0 + jj] * (lo + t);
    const long prod = pow2 ? (full & (n_time - 1)) : (full % n_time);          // exact phase index
    double sn, cs;
    sincospi(2.0 * (double)prod / (double)n_time, &sn, &cs);
    E[i] = make_double2(cs, sn);
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
    cublasHandle_t cb = nullptr;
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(g_dos); cudaFree(g_basis); cudaFree(g_G); cudaFree(g_E); cudaFree(g_out); cudaFree(g_k);
        if (cb) cublasDestroy(cb);
    };
#define BB_RW_TRY(call)                                                                                      \
    do {                                                                                                     \
        if ((call) != cudaSuccess) { rc = bb_fail(std::string(#call) + ": " + cudaGetErrorString(cudaGetLastError())); cleanup(); return rc; } \
    } while (0)
    const size_t n = (size_t)n_freq_sel;
    BB_RW_TRY(cudaMalloc(&g_dos, n * n_det * sizeof(double2)));
    BB_RW_TRY(cudaMalloc(&g_basis, n * n_basis * sizeof(double2)));
    BB_RW_TRY(cudaMalloc(&g_G, n * n_basis * sizeof(double2)));
    BB_RW_TRY(cudaMalloc(&g_k, n * sizeof(int)));
    const int slab = n_freq_sel < BB_RW_SLAB ? n_freq_sel : BB_RW_SLAB;
    BB_RW_TRY(cudaMalloc(&g_E, (size_t)n_win * slab * sizeof(double2)));
    BB_RW_TRY(cudaMalloc(&g_out, (size_t)n_win * n_basis * sizeof(double2)));
    BB_RW_TRY(cudaMemcpy(g_dos, d_over_s, n * n_det * sizeof(double2), cudaMemcpyHostToDevice));
    BB_RW_TRY(cudaMemcpy(g_basis, basis, n * n_basis * sizeof(double2), cudaMemcpyHostToDevice));
    BB_RW_TRY(cudaMemcpy(g_k, bin_index, n * sizeof(int), cudaMemcpyHostToDevice));
    if (cublasCreate(&cb) != CUBLAS_STATUS_SUCCESS) { cleanup(); return bb_fail("cublasCreate failed"); }
    const cuDoubleComplex alpha = make_cuDoubleComplex(4.0 / duration, 0.0);
    const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0), zero = make_cuDoubleComplex(0.0, 0.0);
    const bool pow2 = (n_time & (n_time - 1)) == 0;
    for (int det = 0; det < n_det; ++det) {
        const long totg = (long)n * n_basis;
        bb_rw_g_kernel<<<(unsigned)((totg + 255) / 256), 256>>>(g_dos + (size_t)det * n, g_basis, n_freq_sel, n_basis, g_G);
        for (int j0 = 0; j0 < n_freq_sel; j0 += slab) {
            const int nj = (n_freq_sel - j0) < slab ? (n_freq_sel - j0) : slab;
            const long tot = (long)n_win * nj;
            bb_rw_phase_kernel<<<(unsigned)((tot + 255) / 256), 256>>>(g_k, j0, nj, lo, n_win, n_time, pow2 ? 1 : 0, g_E);
            // out[t][b] (row-major) = column-major [n_basis x n_win] = G^T [n_basis x nj] * E [nj x n_win]
            if (cublasZgemm(cb, CUBLAS_OP_N, CUBLAS_OP_N, n_basis, n_win, nj, &alpha,
                            reinterpret_cast<const cuDoubleComplex*>(g_G + (size_t)j0 * n_basis), n_basis,
                            reinterpret_cast<const cuDoubleComplex*>(g_E), nj, j0 ? &one : &zero,
                            reinterpret_cast<cuDoubleComplex*>(g_out), n_basis) != CUBLAS_STATUS_SUCCESS) {
                cleanup();
                return bb_fail("bb_build_roq_linear_weights: cublasZgemm failed");
            }
        }
        BB_RW_TRY(cudaMemcpy(out + (size_t)det * n_win * n_basis * 2, g_out, (size_t)n_win * n_basis * sizeof(double2),
                             cudaMemcpyDeviceToHost));
    }
    BB_RW_TRY(cudaDeviceSynchronize());
#undef BB_RW_TRY
    cleanup();
    return 0;
}
