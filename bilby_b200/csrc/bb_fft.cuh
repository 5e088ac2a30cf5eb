// Batched complex FP64 FFT for transforms that do not fit one CTA's shared memory (included by bb_kernels.cu).
//
// Used by the IFFT-FFT form of (h, h) of the multi-banded likelihood (bilby/gw/likelihood/multiband.py:613-646,
// 766-787: per sample, detector and band an inverse and a forward transform of M^(b) = 2^9 .. 2^15 points).  K4's
// transform (bb_timemarg.cuh) lives in one CTA's shared memory and stops at 8192 points; this one is the four-step
// algorithm over global memory (the working set of a batch stays in L2):
//   N = N1 N2, x[N2 n1 + n2], X[k1 + N1 k2]:
//   pass 1 (bb_fft_cols_kernel)  for every column n2: FFT_N1 over n1, times W_N^(n2 k1)   -> A[k1][n2]
//   pass 2 (bb_fft_rows_kernel)  for every row k1:    FFT_N2 over n2                      -> X[k1 + N1 k2]
// Each CTA takes a tile of 16 (8 for 512-point sub-transforms) adjacent columns / rows so that every global access is a
// run of >= 128 contiguous bytes, transposes it into shared memory and runs the sub-transforms there with an
// autosorting radix-2 Stockham scheme (two buffers, twiddles from an exact table: sincospi of integer fractions).
// Forward sign (e^{-2 pi i}); the inverse is conj(FFT(conj x)) at the call sites.  HBM/L2-bound, not FP64-bound.
#pragma once

#define BB_FFT_THREADS 256
#define BB_FFT_MAX_SUB 512

// `count` independent transforms of length n (power of two >= 2), transform c at a[c * ld .. c * ld + n); result returned
// in the buffer the function returns (a or b).  tw[m] = exp(-2 pi i m / n), m < n / 2.  All threads of the CTA call it.
__device__ __forceinline__ double2* bb_fft_smem(double2* a, double2* b, int n, int log2n, int count, int ld,
                                                const double2* tw) {
    const int half = n >> 1;
    for (int s = 0; s < log2n; ++s) {
        const int ns = 1 << s, tstride = half >> s;          // twiddle(k, 2 ns) = tw[k * n / (2 ns)]
        for (int i = threadIdx.x; i < count * half; i += blockDim.x) {
            const int c = i / half, j = i - c * half;
            const int k = j & (ns - 1);
            const double2 u0 = a[c * ld + j], x1 = a[c * ld + j + half], w = tw[k * tstride];
            const double2 u1 = make_double2(fma(w.x, x1.x, -w.y * x1.y), fma(w.x, x1.y, w.y * x1.x));
            const int j0 = ((j >> s) << (s + 1)) + k;
            b[c * ld + j0] = make_double2(u0.x + u1.x, u0.y + u1.y);
            b[c * ld + j0 + ns] = make_double2(u0.x - u1.x, u0.y - u1.y);
        }
        __syncthreads();
        double2* t = a; a = b; b = t;
    }
    return a;
}

__device__ __forceinline__ void bb_fft_twiddle_table(double2* tw, int n) {
    for (int m = threadIdx.x; m < (n >> 1); m += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * (double)m / (double)n, &sn, &cs);
        tw[m] = make_double2(cs, sn);
    }
}

// pass 1: tile = CW adjacent columns of one transform
__global__ void __launch_bounds__(BB_FFT_THREADS)
bb_fft_cols_kernel(const double2* __restrict__ in, double2* __restrict__ out, long batch, int N1, int log2n1, int N2,
                   int CW) {
    extern __shared__ __align__(16) double2 fft_smem[];
    const int ld = N1 + 1;
    double2* a = fft_smem;
    double2* b = a + CW * ld;
    double2* tw = b + CW * ld;
    const long N = (long)N1 * N2;
    const int tiles = N2 / CW;
    bb_fft_twiddle_table(tw, N1);
    for (long t = blockIdx.x; t < batch * tiles; t += gridDim.x) {
        const long bi = t / tiles;
        const int n2_0 = (int)(t - bi * tiles) * CW;
        const double2* src = in + bi * N;
        __syncthreads();
        for (int i = threadIdx.x; i < N1 * CW; i += blockDim.x) {
            const int n1 = i / CW, c = i - n1 * CW;
            a[c * ld + n1] = src[(long)n1 * N2 + n2_0 + c];
        }
        __syncthreads();
        double2* r = bb_fft_smem(a, b, N1, log2n1, CW, ld, tw);
        double2* dst = out + bi * N;
        for (int i = threadIdx.x; i < N1 * CW; i += blockDim.x) {
            const int k1 = i / CW, c = i - k1 * CW;
            const long m = ((long)k1 * (n2_0 + c)) % N;                 // exact phase index of W_N^(n2 k1)
            double sn, cs;
            sincospi(-2.0 * (double)m / (double)N, &sn, &cs);
            const double2 v = r[c * ld + k1];
            dst[(long)k1 * N2 + n2_0 + c] = make_double2(fma(v.x, cs, -v.y * sn), fma(v.x, sn, v.y * cs));
        }
    }
}

// pass 2: tile = CW adjacent rows k1 of one transform; output X[k1 + N1 k2]
__global__ void __launch_bounds__(BB_FFT_THREADS)
bb_fft_rows_kernel(const double2* __restrict__ in, double2* __restrict__ out, long batch, int N1, int N2, int log2n2,
                   int CW) {
    extern __shared__ __align__(16) double2 fft_smem[];
    const int ld = N2 + 1;
    double2* a = fft_smem;
    double2* b = a + CW * ld;
    double2* tw = b + CW * ld;
    const long N = (long)N1 * N2;
    const int tiles = N1 / CW;
    bb_fft_twiddle_table(tw, N2);
    for (long t = blockIdx.x; t < batch * tiles; t += gridDim.x) {
        const long bi = t / tiles;
        const int k1_0 = (int)(t - bi * tiles) * CW;
        const double2* src = in + bi * N;
        __syncthreads();
        for (int i = threadIdx.x; i < N2 * CW; i += blockDim.x) {
            const int c = i / N2, n2 = i - c * N2;
            a[c * ld + n2] = src[(long)(k1_0 + c) * N2 + n2];
        }
        __syncthreads();
        double2* r = bb_fft_smem(a, b, N2, log2n2, CW, ld, tw);
        double2* dst = out + bi * N;
        for (int i = threadIdx.x; i < N2 * CW; i += blockDim.x) {
            const int k2 = i / CW, c = i - k2 * CW;
            dst[(long)k2 * N1 + k1_0 + c] = r[c * ld + k2];
        }
    }
}

// out-of-place forward FFT of `batch` contiguous transforms of N = 2^log2n points, 2^8 <= N <= 2^18; `in` is overwritten
// (it holds the intermediate of pass 1 ... no: scratch), `out` receives the natural-order spectrum.
static int bb_fft_forward(const double2* in, double2* scratch, double2* out, long batch, int log2n, int sm_count,
                          cudaStream_t st) {
    if (log2n < 8 || log2n > 18) return bb_fail("bb_fft_forward: transform length must be 2^8 .. 2^18");
    const int l1 = (log2n + 1) / 2, l2 = log2n - l1;
    const int N1 = 1 << l1, N2 = 1 << l2;
    const int cw1 = N1 > 256 ? 8 : 16, cw2 = N2 > 256 ? 8 : 16;
    const size_t sm1 = ((size_t)2 * cw1 * (N1 + 1) + N1 / 2) * sizeof(double2);
    const size_t sm2 = ((size_t)2 * cw2 * (N2 + 1) + N2 / 2) * sizeof(double2);
    BB_CUDA(cudaFuncSetAttribute(bb_fft_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
    BB_CUDA(cudaFuncSetAttribute(bb_fft_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
    long g1 = batch * (N2 / cw1), g2 = batch * (N1 / cw2);
    const long cap = (long)sm_count * 4;
    bb_fft_cols_kernel<<<(unsigned)(g1 < cap ? g1 : cap), BB_FFT_THREADS, sm1, st>>>(in, scratch, batch, N1, l1, N2, cw1);
    BB_CUDA(cudaGetLastError());
    bb_fft_rows_kernel<<<(unsigned)(g2 < cap ? g2 : cap), BB_FFT_THREADS, sm2, st>>>(scratch, out, batch, N1, N2, l2, cw2);
    BB_CUDA(cudaGetLastError());
    return 0;
}
