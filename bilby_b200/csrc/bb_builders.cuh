// Device builders for two set-up artefacts (included by bb_kernels.cu):
//   * the summary data of relative binning     bilby/gw/likelihood/relative.py:319-363
//   * the quadratic ROQ weights                bilby/gw/likelihood/roq.py:976-1004
// (the distance lookup table and the linear ROQ weights have their own builders: bb_build_distance_table,
// bb_build_roq_linear_weights).  Both are segmented / strided reductions over the frequency axis: one CTA per output
// element, lanes stride over the bins, a block reduction at the end.  HBM-bound, run once per data set.
#pragma once

#define BB_BLD_THREADS 256

__device__ __forceinline__ double bb_block_sum(double v, double* scratch /* [BB_BLD_THREADS / 32] */) {
    v = bb_warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < BB_BLD_THREADS / 32; ++i) t += scratch[i];
    return t;
}

// a0 = <h0|d>, a1 = <h0|d (f - fc)>, b0 = <h0|h0>, b1 = <h0|h0 (f - fc)> over the grid bins [bin_start[b], bin_start[b+1])
// with the data tiles already holding (4/T) d/S and (4/T)/S (zero outside the mask)
__global__ void __launch_bounds__(BB_BLD_THREADS)
bb_relbin_summary_kernel(const double2* __restrict__ h0 /* [n_det][n_freq] */, BBTiles tiles, int n_freq, double df,
                         const int* __restrict__ bin_start, const double* __restrict__ centre, int n_bins,
                         double* __restrict__ out /* [n_det][4][n_bins][2] */) {
    __shared__ double scratch[BB_BLD_THREADS / 32];
    const int b = blockIdx.x, d = blockIdx.y;
    const double fc = centre[b];
    double s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int k = bin_start[b] + threadIdx.x; k < bin_start[b + 1]; k += BB_BLD_THREADS) {
        const double2 h = h0[(size_t)d * n_freq + k];
        const double2 ds = tiles.ds[(size_t)d * tiles.n_pad + k];
        const double hdr = h.x * ds.x + h.y * ds.y, hdi = h.x * ds.y - h.y * ds.x;     // conj(h0) d / S
        const double hh = (h.x * h.x + h.y * h.y) * tiles.is[(size_t)d * tiles.n_pad + k];
        const double w = (double)k * df - fc;
        s[0] += hdr; s[1] += hdi; s[2] = fma(hdr, w, s[2]); s[3] = fma(hdi, w, s[3]); s[4] += hh; s[5] = fma(hh, w, s[5]);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) s[i] = bb_block_sum(s[i], scratch);
    if (threadIdx.x == 0) {
        double* o = out + (size_t)d * 4 * n_bins * 2;
        o[(0 * n_bins + b) * 2] = s[0]; o[(0 * n_bins + b) * 2 + 1] = s[1];
        o[(1 * n_bins + b) * 2] = s[2]; o[(1 * n_bins + b) * 2 + 1] = s[3];
        o[(2 * n_bins + b) * 2] = s[4]; o[(2 * n_bins + b) * 2 + 1] = 0.0;
        o[(3 * n_bins + b) * 2] = s[5]; o[(3 * n_bins + b) * 2 + 1] = 0.0;
    }
}

extern "C" int bb_build_relbin_summary_data(bb_handle* h, int n_bins, const int* bin_start, const double* centre,
                                            const double* fiducial, double* out) {
    if (!h || !h->have_network) return bb_fail("bb_build_relbin_summary_data: network not set");
    if (n_bins < 1 || !bin_start || !centre || !fiducial || !out) return bb_fail("bb_build_relbin_summary_data: bad arguments");
    BB_CUDA(cudaSetDevice(h->device));
    const int nd = h->net.n_det, nf = h->net.n_freq;
    for (int b = 0; b < n_bins; ++b)
        if (bin_start[b] < 0 || bin_start[b + 1] < bin_start[b] || bin_start[b + 1] > nf)
            return bb_fail("bb_build_relbin_summary_data: bin edges outside the frequency grid");
    double2* d_h0 = nullptr;
    int* d_bs = nullptr;
    double *d_c = nullptr, *d_out = nullptr;
    auto cleanup = [&]() { cudaFree(d_h0); cudaFree(d_bs); cudaFree(d_c); cudaFree(d_out); };
    const size_t out_n = (size_t)nd * 4 * n_bins * 2;
    if (cudaMalloc(&d_h0, (size_t)nd * nf * sizeof(double2)) != cudaSuccess || cudaMalloc(&d_bs, (n_bins + 1) * sizeof(int)) != cudaSuccess
        || cudaMalloc(&d_c, n_bins * sizeof(double)) != cudaSuccess || cudaMalloc(&d_out, out_n * sizeof(double)) != cudaSuccess) {
        cleanup();
        return bb_fail("bb_build_relbin_summary_data: out of device memory");
    }
    cudaMemcpy(d_h0, fiducial, (size_t)nd * nf * sizeof(double2), cudaMemcpyHostToDevice);
    cudaMemcpy(d_bs, bin_start, (n_bins + 1) * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(d_c, centre, n_bins * sizeof(double), cudaMemcpyHostToDevice);
    bb_relbin_summary_kernel<<<dim3(n_bins, nd), BB_BLD_THREADS>>>(d_h0, bb_tiles(h), nf, h->net.df, d_bs, d_c, n_bins, d_out);
    h->launches++;
    const cudaError_t e = cudaMemcpy(out, d_out, out_n * sizeof(double), cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) return bb_fail(std::string("bb_build_relbin_summary_data: ") + cudaGetErrorString(e));
    return 0;
}

// w[d][b] = (4 / T) sum_j Re(B[b][j]) / S_d[j]
__global__ void __launch_bounds__(BB_BLD_THREADS)
bb_roq_quadratic_weights_kernel(const double* __restrict__ inv_psd /* [n_det][n] */, const double* __restrict__ basis_re
                                /* [n_basis][n] */, int n, double norm, int n_basis, double* __restrict__ out) {
    __shared__ double scratch[BB_BLD_THREADS / 32];
    const int b = blockIdx.x, d = blockIdx.y;
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += BB_BLD_THREADS) s = fma(basis_re[(size_t)b * n + j], inv_psd[(size_t)d * n + j], s);
    s = bb_block_sum(s, scratch);
    if (threadIdx.x == 0) out[(size_t)d * n_basis + b] = norm * s;
}

extern "C" int bb_build_roq_quadratic_weights(int device, int n_det, int n_freq_sel, const double* inv_psd, int n_basis,
                                              const double* basis_real, double duration, double* out) {
    if (n_det < 1 || n_freq_sel < 1 || n_basis < 1 || !inv_psd || !basis_real || !out)
        return bb_fail("bb_build_roq_quadratic_weights: bad arguments");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        return bb_fail("bb_build_roq_quadratic_weights: no such CUDA device (bilby_b200 has no CPU path)");
    BB_CUDA(cudaSetDevice(device));
    double *d_p = nullptr, *d_b = nullptr, *d_o = nullptr;
    auto cleanup = [&]() { cudaFree(d_p); cudaFree(d_b); cudaFree(d_o); };
    const size_t n = (size_t)n_freq_sel;
    if (cudaMalloc(&d_p, n * n_det * sizeof(double)) != cudaSuccess || cudaMalloc(&d_b, n * n_basis * sizeof(double)) != cudaSuccess
        || cudaMalloc(&d_o, (size_t)n_det * n_basis * sizeof(double)) != cudaSuccess) {
        cleanup();
        return bb_fail("bb_build_roq_quadratic_weights: out of device memory");
    }
    cudaMemcpy(d_p, inv_psd, n * n_det * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, basis_real, n * n_basis * sizeof(double), cudaMemcpyHostToDevice);
    bb_roq_quadratic_weights_kernel<<<dim3(n_basis, n_det), BB_BLD_THREADS>>>(d_p, d_b, n_freq_sel, 4.0 / duration, n_basis, d_o);
    const cudaError_t e = cudaMemcpy(out, d_o, (size_t)n_det * n_basis * sizeof(double), cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) return bb_fail(std::string("bb_build_roq_quadratic_weights: ") + cudaGetErrorString(e));
    return 0;
}
