// Calibration marginalisation (included by bb_kernels.cu after the launch helpers).
//
// Replaces, per sample, GravitationalWaveTransient.calculate_snrs' calibration arrays (bilby/gw/likelihood/
// base.py:333-346) and calibration_marginalized_likelihood (:860-877):
//
//   D[c] = sum_det sum_k (4/T) conj(d_k) h_det,k / S_k * C_det,c(f_k)        (complex; note conj(d) h, base.py:334-339)
//   H[c] = sum_det sum_k (4/T) |h_det,k|^2 / S_k * |C_det,c(f_k)|^2          (real; base.py:341-346)
//   lnL  = logsumexp_c( point likelihood(D[c], H[c]) ) - ln n_curves          (plain | phase | distance(+phase))
//
// over the n_curves response curves drawn once at set-up (base.py:1037-1051, calibration.py:503-591).  This is the one
// place on the path where the arithmetic is a dense contraction over the frequency axis: per chunk of samples
//   X_det [chunk x ldk] (complex) , Y_det [chunk x ldk] (real)             <- bb_calmarg_series_kernel
//   D = sum_det X_det C_det^T , H = sum_det Y_det A_det^T  with C_det, A_det = [n_curves x ldk]
//                                                 <- bb_gemm.cuh (FP64 tensor path, mma.sync.m8n8k4.f64 = DMMA)
//   lnL                                                                           <- bb_calmarg_epilogue_kernel
// The detectors are K-segments of one GEMM, so the sum over detectors comes out of the contraction.  All four operands
// live in the GEMM's packed layout (bb_pk: row tiles x slabs of 16 bins), written that way by the series kernel / the
// upload.
#pragma once

// samples per chunk: 148 x 64, so that the complex GEMM (64-row tiles, 16 column tiles for 1000 curves) and the real one
// (128-row tiles) both split into whole waves of 2 x 148 CTAs (2048 lost 13 % to a half-empty last wave)
#define BB_CM_CHUNK 9472

template <int NDET, int APPROX>
__global__ void bb_calmarg_series_kernel(const double* __restrict__ coef, const unsigned* __restrict__ perm, long s0,
                                         int m, BBTiles tiles, int n_freq, double df, int ldk, int k_lo, int k_hi,
                                         double2* __restrict__ X, size_t x_stride, double* __restrict__ Y, size_t y_stride) {
    const int s = blockIdx.y;
    const int k = k_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m || k >= k_hi) return;
    const long sample = perm ? (long)perm[s0 + s] : s0 + s;
    const double* c = coef + sample * BC_NCOEF;
    const bool active = k < n_freq && c[BC_STATUS] == 0.0 && k >= (int)c[BC_KMIN] && k < (int)c[BC_KMAX];
    double A = 0.0, ph = 0.0;
    const double f = (double)k * df;
    if (active) bb_wave<APPROX>(c, f, tiles.u[k], tiles.lf[k], tiles.q34[k], &A, &ph);
#pragma unroll
    for (int d = 0; d < NDET; ++d) {
        double2 x = make_double2(0.0, 0.0);
        double y = 0.0;
        if (active) {
            const double* cd = c + BC_DET + BC_DSTRIDE * d;
            double sn, cs;
            bb_sincospi(ph + cd[2] * f, &sn, &cs);       // h22 e^{-2 pi i f (dt0 + delay)} = A (cs - i sn)
            const double hr = A * (cd[0] * cs + cd[1] * sn), hi = A * (cd[1] * cs - cd[0] * sn);      // K h
            const double2 dd = tiles.ds[(size_t)d * tiles.n_pad + k];                                 // (4/T) d / S
            x = make_double2(hr * dd.x + hi * dd.y, hi * dd.x - hr * dd.y);                           // h conj(d) / S
            y = (hr * hr + hi * hi) * tiles.is[(size_t)d * tiles.n_pad + k];
        }
        X[d * x_stride + bb_pk(s, k, BB_GEMM_TR_A(true), ldk >> 4)] = x;
        Y[d * y_stride + bb_pk(s, k, BB_GEMM_TR_A(false), ldk >> 4)] = y;
    }
}

// one warp per sample: logsumexp over the curves
__global__ void bb_calmarg_epilogue_kernel(const double* __restrict__ coef, const unsigned* __restrict__ perm, long s0,
                                           int m, const double2* __restrict__ D, const double* __restrict__ H,
                                           int n_curves, BBMarg marg, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= m) return;
    const long sample = perm ? (long)perm[s0 + s] : s0 + s;
    const double* c = coef + sample * BC_NCOEF;
    if (c[BC_STATUS] != 0.0) {
        if (lane == 0) out[sample] = -DBL_MAX;
        return;
    }
    const double dist = c[BC_DISTANCE];
    double mx = -INFINITY, sum = 0.0;
    for (int i = lane; i < n_curves; i += 32) {
        const double2 d = D[(size_t)s * n_curves + i];
        const double l = bb_point_lnl(marg, d.x, d.y, H[(size_t)s * n_curves + i], dist);
        if (l == -INFINITY) continue;
        if (l > mx) { sum = sum * exp(mx - l) + 1.0; mx = l; }
        else sum += exp(l - mx);
    }
    double gmx = mx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gmx = fmax(gmx, __shfl_xor_sync(0xffffffffu, gmx, o));
    double part = (mx == -INFINITY) ? 0.0 : sum * exp(mx - gmx);
    part = bb_warp_sum(part);
    if (lane == 0) out[sample] = (gmx == -INFINITY) ? -INFINITY : (log(part) + gmx) - log((double)n_curves);
}

// Reconstruction (base.py:544-578): one warp per sample draws the index of a response curve from the curves'
// posterior, p_i ~ exp(lnL_i - max), by inverse CDF with the caller's unit-interval draw u (numpy's
// Generator.choice(n, p): cdf = cumsum(p) / cumsum(p)[-1], index = #{cdf <= u}), and hands the inner products of THAT
// curve to the distance / phase reconstruction (base.py:289-290): <h|d> = conj(D[index]), <h|h> = H[index].
struct BBCalSelect {
    const double* uniforms;   // [n][3], column 0 = the calibration draw; nullptr: marginalise (logsumexp) instead
    double* snr;              // [n][n_det][3]: the selected curve's totals in detector 0, zeros elsewhere
    double* out;              // [n][3], column 0 = recalib_index
};

__global__ void bb_calmarg_select_kernel(const double* __restrict__ coef, const unsigned* __restrict__ perm, long s0,
                                         int m, const double2* __restrict__ D, const double* __restrict__ H,
                                         int n_curves, int n_det, BBMarg marg, BBCalSelect sel) {
    const int lane = threadIdx.x & 31;
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= m) return;
    const long sample = perm ? (long)perm[s0 + s] : s0 + s;
    const double* c = coef + sample * BC_NCOEF;
    double* o = sel.snr + sample * n_det * 3;
    if (lane < n_det * 3 && lane >= 3) o[lane] = 0.0;
    if (c[BC_STATUS] != 0.0) {
        if (lane == 0) {
            sel.out[sample * 3] = nan("");
            o[0] = 0.0; o[1] = 0.0; o[2] = nan("");
        }
        return;
    }
    const double dist = c[BC_DISTANCE];
    double mx = -INFINITY;
    for (int i = lane; i < n_curves; i += 32) {
        const double2 d = D[(size_t)s * n_curves + i];
        mx = fmax(mx, bb_point_lnl(marg, d.x, d.y, H[(size_t)s * n_curves + i], dist));
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, k));
    double total = 0.0;
    for (int i = lane; i < n_curves; i += 32) {
        const double2 d = D[(size_t)s * n_curves + i];
        total += exp(bb_point_lnl(marg, d.x, d.y, H[(size_t)s * n_curves + i], dist) - mx);
    }
    total = bb_warp_sum(total);
    const double u = sel.uniforms[sample * 3];
    double run = 0.0;          // cumulative sum before this group of 32 curves
    int count = 0;             // curves whose cdf value is <= u
    for (int base = 0; base < n_curves; base += 32) {
        const int i = base + lane;
        double p = 0.0;
        if (i < n_curves) {
            const double2 d = D[(size_t)s * n_curves + i];
            p = exp(bb_point_lnl(marg, d.x, d.y, H[(size_t)s * n_curves + i], dist) - mx);
        }
        double scan = p;       // inclusive scan over the lanes
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, scan, k);
            if (lane >= k) scan += up;
        }
        const bool below = (i < n_curves) && ((run + scan) / total <= u);
        count += __popc(__ballot_sync(0xffffffffu, below));
        run += __shfl_sync(0xffffffffu, scan, 31);
    }
    if (count > n_curves - 1) count = n_curves - 1;
    if (lane == 0) {
        const double2 d = D[(size_t)s * n_curves + count];
        sel.out[sample * 3] = (double)count;
        o[0] = d.x;
        o[1] = -d.y;
        o[2] = H[(size_t)s * n_curves + count];
    }
}

// Time + calibration marginalisation (base.py:305-323, 860-866, 794-820): one CTA per (sample, response curve).
//   series_c[k] = sum_det X_det[k] C_det,c[k]  (X = h conj(d)/S with 4/T, from bb_calmarg_series_kernel; zero outside
//   the chunk's active window), in-shared-memory FFT (bb_tm_fft_dif), weighted logsumexp over the times inside the
//   geocent_time prior with <h|h>_c from the real contraction (bb_tm_finish) -> L[s][c]; bb_calmarg_lse_kernel then takes
//   logsumexp_c L - log(n_curves).
#define BB_CMT_THREADS 256      // two CTAs per SM: the fill of one pair overlaps the transform of the other
template <int NDET>
__global__ void __launch_bounds__(BB_CMT_THREADS, 2)
bb_calmarg_time_kernel(const double* __restrict__ coef, const unsigned* __restrict__ perm, long s0, int m,
                       const double2* __restrict__ Xs, size_t x_stride, const double2* __restrict__ C, size_t c_stride,
                       const double* __restrict__ H, int n_curves, int ldk, int k_lo, int k_hi, int nfft, int log2n,
                       const double2* __restrict__ twiddle, BBMarg marg, double start_time, double duration,
                       double* __restrict__ L) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* X = reinterpret_cast<double2*>(smem_raw);
    int plan_a, plan_b, ps;
    bb_tm_plan(log2n, &plan_a, &plan_b, &ps);
    double* red = reinterpret_cast<double*>(X + bb_tm_series_elems(nfft, ps));      // [32]
    double2* wl = reinterpret_cast<double2*>(red + 32);                              // [nfft / 256] (pruned finish)
    const int tid = threadIdx.x;
    const long S = ldk >> 4;
    const int Lw = nfft >> 8;
    if (log2n >= 9) {
        for (int e = tid; e < Lw; e += BB_CMT_THREADS) {
            const double2 w = twiddle[(e & (Lw / 2 - 1)) << 8];        // exp(-2 pi i e / L) from the nfft-point table
            wl[e] = (e < Lw / 2) ? w : make_double2(-w.x, -w.y);
        }
    }
    const long pairs = (long)m * n_curves;
    for (long pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
        const int s = (int)(pair / n_curves), c = (int)(pair - (long)s * n_curves);
        const long sample = perm ? (long)perm[s0 + s] : s0 + s;
        const double* rec = coef + sample * BC_NCOEF;
        __syncthreads();
        if (rec[BC_STATUS] != 0.0) {
            if (tid == 0) L[pair] = -DBL_MAX;
            continue;
        }
        for (int k = tid; k < nfft; k += BB_CMT_THREADS) {
            double vr = 0.0, vi = 0.0;
            if (k >= k_lo && k < k_hi) {
#pragma unroll
                for (int d = 0; d < NDET; ++d) {
                    const double2 x = Xs[d * x_stride + bb_pk(s, k, BB_GEMM_TR_A(true), S)];
                    const double2 q = C[d * c_stride + bb_pk(c, k, BB_GEMM_TR_B, S)];
                    vr = fma(x.x, q.x, fma(-x.y, q.y, vr));
                    vi = fma(x.x, q.y, fma(x.y, q.x, vi));
                }
            }
            X[bb_tm_pos(k, ps)] = make_double2(vr, vi);
        }
        __syncthreads();
        // K4b's pruned transform (bb_timemarg_split.cuh): eight stages, then only the outputs inside the time prior are
        // formed by a direct nfft/256-term sum - per (sample, curve) pair 8 instead of log2n stages
        int j_lo = 0, j_hi = nfft;
        if (log2n >= 9) bb_tm_window(marg, rec[BC_JITTER], start_time, duration, nfft, &j_lo, &j_hi);
        if (log2n >= 9 && j_hi - j_lo <= BB_SFT_PRUNE_MAX) {
            bb_tm_pass<4, BB_CMT_THREADS>(X, nfft, 0, ps, twiddle);
            bb_tm_pass<4, BB_CMT_THREADS>(X, nfft, 4, ps, twiddle);
            bb_tm_finish<BB_CMT_THREADS, true>(X, nfft, log2n, marg, H[pair], rec[BC_DISTANCE], rec[BC_JITTER], start_time,
                                               duration, red, L + pair, wl, ps);
        } else {
            bb_tm_fft_dif<BB_CMT_THREADS>(X, nfft, log2n, twiddle);
            bb_tm_finish<BB_CMT_THREADS>(X, nfft, log2n, marg, H[pair], rec[BC_DISTANCE], rec[BC_JITTER], start_time, duration,
                                         red, L + pair);
        }
    }
}

// one warp per sample: logsumexp over the curves of the per-curve (time-marginalised) likelihoods
__global__ void bb_calmarg_lse_kernel(const double* __restrict__ coef, const unsigned* __restrict__ perm, long s0,
                                      int m, const double* __restrict__ L, int n_curves, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= m) return;
    const long sample = perm ? (long)perm[s0 + s] : s0 + s;
    if (coef[sample * BC_NCOEF + BC_STATUS] != 0.0) {
        if (lane == 0) out[sample] = -DBL_MAX;
        return;
    }
    double mx = -INFINITY;
    for (int i = lane; i < n_curves; i += 32) mx = fmax(mx, L[(size_t)s * n_curves + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double sum = 0.0;
    for (int i = lane; i < n_curves; i += 32) {
        const double l = L[(size_t)s * n_curves + i];
        if (l != -INFINITY) sum += exp(l - mx);
    }
    sum = bb_warp_sum(sum);
    if (lane == 0) out[sample] = (mx == -INFINITY) ? -INFINITY : (log(sum) + mx) - log((double)n_curves);
}

// active bin range of every chunk: the samples arrive sorted by active-bin count (longest first), so the first sample
// of a chunk bounds the others from above; kmin is the same for all
__global__ void bb_calmarg_window_kernel(const double* __restrict__ coef, const unsigned* __restrict__ perm, long n,
                                         int n_chunks, int* __restrict__ win /* [n_chunks][2] */) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const long s = perm[(long)c * BB_CM_CHUNK];
    win[2 * c] = (int)coef[s * BC_NCOEF + BC_KMIN];
    win[2 * c + 1] = (int)coef[s * BC_NCOEF + BC_KMAX];
}

template <int NDET, int APPROX>
static int bb_launch_calmarg_t(bb_handle* h, long n, double* out, cudaStream_t st, BBCalSelect sel) {
    const int nc = h->cm_n_curves, ldk = h->cm_ldk;
    long cap = n < BB_CM_CHUNK ? (n + 127) / 128 * 128 : BB_CM_CHUNK;        // rows the scratch buffers hold
    if (cap < h->cm_cap) cap = h->cm_cap;
    const size_t x_stride = bb_pk_elems(cap, ldk, BB_GEMM_TR_A(true)), y_stride = bb_pk_elems(cap, ldk, BB_GEMM_TR_A(false));
    const size_t c_stride = bb_pk_elems(nc, ldk, BB_GEMM_TR_B), a_stride = bb_pk_elems(nc, ldk, BB_GEMM_TR_B_OF(false));
    if (!h->d_cm_X || cap > h->cm_cap) {
        cudaFree(h->d_cm_X); cudaFree(h->d_cm_Y); cudaFree(h->d_cm_D); cudaFree(h->d_cm_H);
        h->d_cm_X = h->d_cm_D = nullptr;
        h->d_cm_Y = h->d_cm_H = nullptr;
        h->cm_cap = 0;
        BB_CUDA(cudaMalloc(&h->d_cm_X, NDET * x_stride * sizeof(double2)));
        BB_CUDA(cudaMalloc(&h->d_cm_Y, NDET * y_stride * sizeof(double)));
        BB_CUDA(cudaMemsetAsync(h->d_cm_X, 0, NDET * x_stride * sizeof(double2), st));
        BB_CUDA(cudaMemsetAsync(h->d_cm_Y, 0, NDET * y_stride * sizeof(double), st));
        BB_CUDA(cudaMalloc(&h->d_cm_D, (size_t)cap * nc * sizeof(double2)));
        BB_CUDA(cudaMalloc(&h->d_cm_H, (size_t)cap * nc * sizeof(double)));
        h->cm_cap = cap;
    }
    BBMarg point = h->marg;
    point.flags &= ~BB_MARG_TIME;
    const bool time_marg = (h->marg.flags & BB_MARG_TIME) != 0;
    int tm_log2n = 0, tm_per_sm = 1;
    size_t tm_smem = 0;
    if (time_marg) {
        if (sel.uniforms) return bb_fail("reconstruction with time + calibration marginalisation is not supported");
        while ((1 << tm_log2n) < h->nfft) ++tm_log2n;
        int pa, pb, ps;
        bb_tm_plan(tm_log2n, &pa, &pb, &ps);
        tm_smem = bb_tm_series_elems(h->nfft, ps) * sizeof(double2) + 32 * sizeof(double) + (size_t)(h->nfft >> 8) * sizeof(double2);
        if (tm_smem > 227 * 1024) return bb_fail("time + calibration marginalisation: series does not fit shared memory");
        BB_CUDA(cudaFuncSetAttribute(bb_calmarg_time_kernel<NDET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tm_smem));
        tm_per_sm = (int)((227 * 1024) / (tm_smem + 1024));
        if (tm_per_sm < 1) tm_per_sm = 1;
        if (tm_per_sm > 2) tm_per_sm = 2;
    }
    // contraction window per chunk (needs the sorted order; one small read-back per call)
    const int n_chunks = (int)((n + BB_CM_CHUNK - 1) / BB_CM_CHUNK);
    const unsigned* perm = (h->perm_valid && !getenv("BB_CM_NOTRIM")) ? h->d_perm : nullptr;
    std::vector<int> win(2 * (size_t)n_chunks);
    if (perm) {
        int* d_win = nullptr;
        BB_CUDA(cudaMalloc(&d_win, win.size() * sizeof(int)));
        bb_calmarg_window_kernel<<<(n_chunks + 127) / 128, 128, 0, st>>>(h->d_coef, perm, n, n_chunks, d_win);
        BB_CUDA(cudaMemcpyAsync(win.data(), d_win, win.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        BB_CUDA(cudaFree(d_win));
        h->launches++;
    }
    BBProfScope prof(h, st);
    for (int c = 0; c < n_chunks; ++c) {
        const long s0 = (long)c * BB_CM_CHUNK;
        const int m = (int)((n - s0) < BB_CM_CHUNK ? (n - s0) : BB_CM_CHUNK);
        int k_lo = 0, k_hi = ldk;
        if (perm) {
            k_lo = win[2 * c] & ~15;                 // whole slabs of 16 bins
            k_hi = (win[2 * c + 1] + 15) & ~15;
            if (k_lo < 0) k_lo = 0;
            if (k_hi > ldk) k_hi = ldk;
            if (k_hi <= k_lo) k_hi = k_lo + 16 <= ldk ? k_lo + 16 : ldk;
        }
        const int kw = k_hi - k_lo;
        dim3 grid((kw + 127) / 128, (unsigned)m);
        bb_calmarg_series_kernel<NDET, APPROX><<<grid, 128, 0, st>>>(h->d_coef, perm, s0, m, bb_tiles(h), h->net.n_freq,
                                                                   h->net.df, ldk, k_lo, k_hi, h->d_cm_X, x_stride, h->d_cm_Y, y_stride);
        BB_CUDA(cudaGetLastError());
        // D[s][c] = sum_det sum_k X_det[s][k] C_det[c][k] (complex) and H[s][c] = sum_det sum_k Y_det[s][k] A_det[c][k]
        // (real): the detectors are K-segments of ONE DMMA GEMM each (bb_gemm.cuh), trimmed to the chunk's active bins
        {
            BBGemmArgs ga{};
            ga.slabs_a = ga.slabs_b = ldk >> 4;
            ga.slab0 = k_lo >> 4; ga.n_slabs = kw >> 4;
            ga.ldc = nc;
            ga.M = m; ga.N = nc; ga.n_seg = NDET; ga.n_batch = 1; ga.accumulate = 0; ga.alpha = 1.0;
            if (!time_marg) {
                for (int d = 0; d < NDET; ++d) {
                    ga.A[d] = h->d_cm_X + d * x_stride;
                    ga.B[d] = h->d_cm_C + d * c_stride;
                }
                ga.C = reinterpret_cast<double*>(h->d_cm_D);
                if (bb_gemm_nt(true, ga, h->sm_count, st)) return 1;
            }
            for (int d = 0; d < NDET; ++d) {
                ga.A[d] = h->d_cm_Y + d * y_stride;
                ga.B[d] = h->d_cm_A + d * a_stride;
            }
            ga.C = h->d_cm_H;
            if (bb_gemm_nt(false, ga, h->sm_count, st)) return 1;
        }
        if (time_marg) {
            // the per-curve likelihoods go to the (unused) <d|h> buffer
            double* L = reinterpret_cast<double*>(h->d_cm_D);
            long grid_t = (long)m * nc;
            if (grid_t > (long)h->sm_count * tm_per_sm) grid_t = (long)h->sm_count * tm_per_sm;
            bb_calmarg_time_kernel<NDET><<<(unsigned)grid_t, BB_CMT_THREADS, tm_smem, st>>>(
                h->d_coef, perm, s0, m, h->d_cm_X, x_stride, h->d_cm_C, c_stride, h->d_cm_H, nc, ldk, k_lo, k_hi, h->nfft, tm_log2n,
                h->d_twiddle, h->marg, h->net.start_time, h->net.duration, L);
            BB_CUDA(cudaGetLastError());
            bb_calmarg_lse_kernel<<<(unsigned)((m * 32L + 127) / 128), 128, 0, st>>>(h->d_coef, perm, s0, m, L, nc, out);
        } else if (sel.uniforms)
            bb_calmarg_select_kernel<<<(unsigned)((m * 32L + 127) / 128), 128, 0, st>>>(h->d_coef, perm, s0, m, h->d_cm_D,
                                                                                       h->d_cm_H, nc, NDET, point, sel);
        else
            bb_calmarg_epilogue_kernel<<<(unsigned)((m * 32L + 127) / 128), 128, 0, st>>>(h->d_coef, perm, s0, m, h->d_cm_D,
                                                                                         h->d_cm_H, nc, point, out);
        BB_CUDA(cudaGetLastError());
        h->launches += 2 + (time_marg ? 2 : 2);
    }
    return 0;
}

static int bb_launch_calmarg(bb_handle* h, long n, double* out, cudaStream_t st, BBCalSelect sel = BBCalSelect{nullptr, nullptr, nullptr}) {
    if (h->marg.flags & BB_MARG_TIME) {
        // base.py:305-323; with distance marginalisation the reference itself fails (shape mismatch in base.py:775-784)
        if (h->marg.flags & BB_MARG_DISTANCE)
            return bb_fail("time + calibration + distance marginalisation is not defined by the reference (shape mismatch)");
        if (h->nfft == 0) return bb_fail("time marginalisation needs n_freq - 1 to be a power of two");
    }
    if (h->kind != 0) return bb_fail("calibration marginalisation: full-grid likelihood only");
    if (h->cal_params) return bb_fail("calibration marginalisation excludes per-sample calibration parameters");
    if (h->shard_lo != 0 || h->shard_hi != h->net.n_freq) return bb_fail("calibration marginalisation cannot be frequency-sharded");
    const bool pd = h->wf.approximant == BB_IMRPHENOMD;
    switch (h->net.n_det) {
        case 1: return pd ? bb_launch_calmarg_t<1, BB_IMRPHENOMD>(h, n, out, st, sel) : bb_launch_calmarg_t<1, BB_TAYLORF2>(h, n, out, st, sel);
        case 2: return pd ? bb_launch_calmarg_t<2, BB_IMRPHENOMD>(h, n, out, st, sel) : bb_launch_calmarg_t<2, BB_TAYLORF2>(h, n, out, st, sel);
        case 3: return pd ? bb_launch_calmarg_t<3, BB_IMRPHENOMD>(h, n, out, st, sel) : bb_launch_calmarg_t<3, BB_TAYLORF2>(h, n, out, st, sel);
        case 4: return pd ? bb_launch_calmarg_t<4, BB_IMRPHENOMD>(h, n, out, st, sel) : bb_launch_calmarg_t<4, BB_TAYLORF2>(h, n, out, st, sel);
    }
    return bb_fail("bad n_det");
}

// curves: [n_det][n_curves][n_freq] complex (re, im) on the full frequency grid (anything outside the mask is
// multiplied by a zero of the data tiles); n_curves = 0 switches the marginalisation off
static int bb_calmarg_upload(bb_handle* h, int n_curves, const double* curves) {
    cudaFree(h->d_cm_C); cudaFree(h->d_cm_A); cudaFree(h->d_cm_X); cudaFree(h->d_cm_Y); cudaFree(h->d_cm_D); cudaFree(h->d_cm_H);
    h->d_cm_C = nullptr; h->d_cm_A = nullptr; h->d_cm_X = nullptr; h->d_cm_Y = nullptr; h->d_cm_D = nullptr; h->d_cm_H = nullptr;
    h->cm_n_curves = 0;
    h->cm_cap = 0;
    if (n_curves <= 0) return 0;
    if (!curves) return bb_fail("bb_set_calibration_marginalization: null curves");
    const int n_det = h->net.n_det, nf = h->net.n_freq;
    const int ldk = (nf + 15) & ~15;                       // whole slabs of 16 bins
    const long S = ldk >> 4;
    const size_t c_stride = bb_pk_elems(n_curves, ldk, BB_GEMM_TR_B), a_stride = bb_pk_elems(n_curves, ldk, BB_GEMM_TR_B_OF(false));
    std::vector<double2> C((size_t)n_det * c_stride, make_double2(0.0, 0.0));
    std::vector<double> A((size_t)n_det * a_stride, 0.0);
    for (int d = 0; d < n_det; ++d)
        for (int c = 0; c < n_curves; ++c) {
            const double* src = curves + ((size_t)d * n_curves + c) * nf * 2;
            for (int k = 0; k < nf; ++k) {
                C[(size_t)d * c_stride + bb_pk(c, k, BB_GEMM_TR_B, S)] = make_double2(src[2 * k], src[2 * k + 1]);
                A[(size_t)d * a_stride + bb_pk(c, k, BB_GEMM_TR_B_OF(false), S)] = src[2 * k] * src[2 * k] + src[2 * k + 1] * src[2 * k + 1];
            }
        }
    BB_CUDA(cudaMalloc(&h->d_cm_C, C.size() * sizeof(double2)));
    BB_CUDA(cudaMalloc(&h->d_cm_A, A.size() * sizeof(double)));
    BB_CUDA(cudaMemcpy(h->d_cm_C, C.data(), C.size() * sizeof(double2), cudaMemcpyHostToDevice));
    BB_CUDA(cudaMemcpy(h->d_cm_A, A.data(), A.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->cm_n_curves = n_curves;
    h->cm_ldk = ldk;
    return 0;
}
