// K4s: time-marginalised likelihood as two co-resident kernels (included by bb_kernels.cu after bb_timemarg.cuh).
//
//   K4a  bb_series_fill_kernel   waveform -> projection -> series X[k] = sum_det h_det conj(d)/S and <h|h>
//                                (base.py:325-330); K1's mapping (warp = sample, lane = bin of a row, tiles by
//                                TMA bulk copies shared by the block's samples); FP64-pipe bound.  The series goes to
//                                a scratch buffer in global memory (active bins only).
//   K4b  bb_series_fft_kernel    one CTA per sample: first radix-16 pass straight from the scratch buffer into
//                                shared memory, remaining passes in shared memory, logsumexp over the prior window
//                                (base.py:794-820); shared-memory bandwidth bound.
//
// The batch is cut into chunks; K4a of chunk c+1 (12 warps, 68 KB shared memory) and K4b of chunk c (4 warps,
// 148 KB) run on two streams and fit one SM together (registers 48K + 16K, shared memory 216 KB), so the FFT of one
// chunk hides behind the FP64 work of the next.  In the fused kernel (bb_timemarg.cuh, kept for small batches) the
// two phases alternate and neither pipe is busy more than half of the time.
#pragma once

#ifndef BB_SF_WARPS
#define BB_SF_WARPS 12
#endif
#define BB_SF_THREADS (BB_SF_WARPS * 32)
#define BB_SF_SB BB_SF_WARPS               // samples per block (one per warp)
#define BB_SF_CHUNK 256                    // bins per tile
#ifndef BB_SFT_THREADS
#define BB_SFT_THREADS 128                 // K4b
#endif
#ifndef BB_SFT_REGS
#define BB_SFT_REGS 128
#endif
#ifndef BB_SFT_PLAN
#define BB_SFT_PLAN 0                      // first eight stages: 0 = two radix-16 passes, 1 = radix-8, radix-8, radix-4
#endif
#define BB_SF_BLOCKS_PER_CHUNK 296         // sample blocks per pipeline chunk (2 per SM)
#define BB_SF_SLOTREC 8                    // doubles per slot handed from K4a to K4b
#define BB_SFT_PRUNE_MAX 640               // widest prior window (in samples) summed directly after two radix-16 passes

template <int NDET>
struct SFTile {
    double ff[BB_SF_CHUNK];       // f, f^(-1/3), f^(1/3), f^(-7/6): K1's columns (bb_k1.cuh); 1/f = t^3 is formed per bin
    double t3[BB_SF_CHUNK];       // where a region needs it, so that K4a (79 KB) still fits one SM together with K4b
    double x3[BB_SF_CHUNK];
    double u7[BB_SF_CHUNK];
    double lf[BB_SF_CHUNK];
    double q34[BB_SF_CHUNK];
    double2 ds[NDET][BB_SF_CHUNK];
    double is[NDET][BB_SF_CHUNK];
};

template <int NDET>
struct SFSmem {
    SFTile<NDET> tile[2];
    double coef[BB_SF_SB][BC_NCOEF];
    unsigned long long bar[2];
    int krange[2];
};

// position in the sorted order of slot i of chunk c: the chunks interleave the sorted sample blocks so that every
// chunk holds the same mix of short and long signals
__host__ __device__ __forceinline__ long bb_sf_pos(int slot, int c, int n_chunks) {
    return ((long)(slot / BB_SF_SB) * n_chunks + c) * BB_SF_SB + slot % BB_SF_SB;
}

template <int NDET>
__device__ __forceinline__ void bb_sf_issue_tile(SFTile<NDET>& t, unsigned long long* bar, const BBTiles& g, int c0) {
    bb_mbar_expect_tx(bar, (unsigned)sizeof(SFTile<NDET>));
    bb_bulk_g2s(t.ff, g.ff + c0, BB_SF_CHUNK * 8, bar);
    bb_bulk_g2s(t.t3, g.t3 + c0, BB_SF_CHUNK * 8, bar);
    bb_bulk_g2s(t.x3, g.x3 + c0, BB_SF_CHUNK * 8, bar);
    bb_bulk_g2s(t.u7, g.u7 + c0, BB_SF_CHUNK * 8, bar);
    bb_bulk_g2s(t.lf, g.lf + c0, BB_SF_CHUNK * 8, bar);
    bb_bulk_g2s(t.q34, g.q34 + c0, BB_SF_CHUNK * 8, bar);
#pragma unroll
    for (int d = 0; d < NDET; ++d) {
        bb_bulk_g2s(t.ds[d], g.ds + (size_t)d * g.n_pad + c0, BB_SF_CHUNK * 16, bar);
        bb_bulk_g2s(t.is[d], g.is + (size_t)d * g.n_pad + c0, BB_SF_CHUNK * 8, bar);
    }
}

template <int NDET>
struct SFState {
    double ramp[NDET][2];    // conj(K_d) exp(+2 pi i f dt_d) at this lane's bin of the current row
    double step[NDET][2];    // advance over one row
    double k2[NDET];         // |K_d|^2
    double hh;
    const double* cal;
    BBCalGrid grid;
    double2* series;         // this sample's row of the scratch buffer
    int nstore;              // bins k < nstore go to the series (nfft; nfft + 1 when the Nyquist bin is wanted)
};

template <int NDET, bool CAL>
__device__ __forceinline__ void bb_sf_bin(SFState<NDET>& st, const SFTile<NDET>& tile, int i, int k, bool act,
                                          double A, double ph) {
    double sn, cs;
    bb_sincospi(act ? ph : 0.0, &sn, &cs);
    A = act ? A : 0.0;
    const double zr = A * cs, zi = A * sn;      // conj(h22 incl. geocentric shift)
    double sr = 0.0, si = 0.0, hs = 0.0;
    BBCalW cw;
    if (CAL) cw = bb_cal_weights(st.grid.n_points, st.grid.l0[0], st.grid.inv_delta[0], tile.lf[i]);
#pragma unroll
    for (int d = 0; d < NDET; ++d) {
        const double rc = st.ramp[d][0], rs = st.ramp[d][1];
        const double2 dd = tile.ds[d][i];
        const double gr = rc * dd.x - rs * dd.y, gi = rc * dd.y + rs * dd.x;    // conj(K) ramp d/S
        if (CAL) {
            double amp1, cr, ci;
            if (d > 0 && !st.grid.shared)
                cw = bb_cal_weights(st.grid.n_points, st.grid.l0[d], st.grid.inv_delta[d], tile.lf[i]);
            bb_cal_apply_v(st.cal + d * 4 * st.grid.n_points, st.grid.n_points, cw, &amp1, &cr, &ci);
            const double qr = amp1 * cr, qi = amp1 * ci;       // conj(C) = amp1 (cr - i ci)
            sr = fma(gr, qr, fma(gi, qi, sr));
            si = fma(gi, qr, fma(-gr, qi, si));
            hs = fma(st.k2[d] * (amp1 * amp1), tile.is[d][i], hs);
        } else {
            sr += gr;
            si += gi;
            hs = fma(st.k2[d], tile.is[d][i], hs);
        }
        st.ramp[d][0] = rc * st.step[d][0] - rs * st.step[d][1];
        st.ramp[d][1] = rc * st.step[d][1] + rs * st.step[d][0];
    }
    st.hh = fma(A * A, hs, st.hh);
    // h conj(d)/S = conj(conj(h) d/S); the Nyquist bin k = nfft is in <h|h> but not in the series
    if (act && k < st.nstore) st.series[k] = make_double2(zr * sr - zi * si, -(zr * si + zi * sr));
}

template <int NDET, int AR, int PR, bool CAL>
__device__ __forceinline__ void bb_sf_rows_pd(SFState<NDET>& st, const SFTile<NDET>& tile, const double* rec, int r0,
                                              int r1, int c0, int lane, int kmin, int kmax, double df) {
    K1Amp<AR> amp;
    K1Ph<PR> phs;
    amp.load(rec);
    phs.load(rec);
    amp.begin((double)(r0 * BB_ROW + lane) * df, (double)BB_ROW * df);
    int k = r0 * BB_ROW + lane, i = k - c0;
    for (int r = r0; r < r1; ++r) {
        const bool act = (k >= kmin) && (k < kmax);
        const double f = tile.ff[i], t = tile.t3[i], x = tile.x3[i];
        const double rf = (PR != 0) ? t * t * t : 0.0;          // 1 / f (intermediate and merger-ringdown phase only)
        const double A = amp.eval(f, x) * tile.u7[i];           // a0 is folded into the region's coefficients (bb_k1.cuh)
        const double ph = phs.eval(f, t, x, rf, tile.lf[i], tile.q34[i]);
        bb_sf_bin<NDET, CAL>(st, tile, i, k, act, A, ph);
        amp.next();
        k += BB_ROW;
        i += BB_ROW;
    }
}

template <int NDET, int APPROX, bool CAL>
__device__ __forceinline__ void bb_sf_rows_generic(SFState<NDET>& st, const SFTile<NDET>& tile, const double* rec,
                                                   int r0, int r1, int c0, int lane, int kmin, int kmax, double df) {
    for (int r = r0; r < r1; ++r) {
        const int k = r * BB_ROW + lane, i = k - c0;
        const bool act = (k >= kmin) && (k < kmax);
        const double f = tile.ff[i];
        double A, ph;
        if (APPROX == BB_TAYLORF2) {
            A = rec[BC_A0] * tile.u7[i];
            ph = bb_taylorf2_phase(rec, f, tile.t3[i], tile.x3[i], tile.lf[i]);
        } else {
            // rows that straddle a region boundary: u = f^(-1/6) = sqrt(u^2) exactly (the square root of a correctly
            // rounded square returns the operand)
            bb_wave<APPROX>(rec, f, sqrt(tile.t3[i]), tile.lf[i], tile.q34[i], &A, &ph);
        }
        bb_sf_bin<NDET, CAL>(st, tile, i, k, act, A, ph);
    }
}

template <int NDET, int APPROX, bool CAL>
__global__ void __maxnreg__(128)
bb_series_fill_kernel(const double* __restrict__ coef, const unsigned* __restrict__ perm, long n, int chunk,
                      int n_chunks, int n_slots, BBTiles tiles, double df, int nstore, int ld,
                      const double* __restrict__ calrec, BBCalGrid grid, double2* __restrict__ series,
                      double* __restrict__ slotrec) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SFSmem<NDET>& sm = *reinterpret_cast<SFSmem<NDET>*>(smem_raw);
    double* sm_cal = reinterpret_cast<double*>(smem_raw + sizeof(SFSmem<NDET>));   // [SB][NDET*4*n_points] (CAL)
    const int cal_len = CAL ? NDET * 4 * grid.n_points : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_blocks = (n_slots + BB_SF_SB - 1) / BB_SF_SB;

    if (tid == 0) {
        bb_mbar_init(&sm.bar[0], 1);
        bb_mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned phase0 = 0, phase1 = 0;
    int issued = 0;

    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const int slot0 = blk * BB_SF_SB;
        const long p0 = bb_sf_pos(slot0, chunk, n_chunks);
        const int ns = (int)max(0L, min((long)BB_SF_SB, n - p0));
        if (tid == 0) { sm.krange[0] = INT_MAX; sm.krange[1] = 0; }
        for (int i = tid; i < ns * BC_NCOEF; i += BB_SF_THREADS) {
            const int sl = i / BC_NCOEF, j = i - sl * BC_NCOEF;
            const long s = perm ? (long)perm[p0 + sl] : p0 + sl;
            sm.coef[sl][j] = coef[s * BC_NCOEF + j];
        }
        if (CAL) {
            for (int i = tid; i < ns * cal_len; i += BB_SF_THREADS) {
                const int sl = i / cal_len, j = i - sl * cal_len;
                const long s = perm ? (long)perm[p0 + sl] : p0 + sl;
                sm_cal[sl * cal_len + j] = calrec[s * cal_len + j];
            }
        }
        __syncthreads();
        if (tid < ns && sm.coef[tid][BC_STATUS] == 0.0) {
            const int k0 = (int)sm.coef[tid][BC_KMIN], k1 = (int)sm.coef[tid][BC_KMAX];
            if (k1 > k0) {
                atomicMin(&sm.krange[0], k0);
                atomicMax(&sm.krange[1], k1);
            }
        }
        __syncthreads();
        const int kb0 = sm.krange[0], kb1 = sm.krange[1];
        const int cb0 = kb0 / BB_SF_CHUNK, cb1 = (kb1 + BB_SF_CHUNK - 1) / BB_SF_CHUNK;

        const bool have = warp < ns && sm.coef[warp < ns ? warp : 0][BC_STATUS] == 0.0;
        const double* rec = sm.coef[warp < ns ? warp : 0];
        int kmin = 0, kmax = 0;
        if (have) {
            kmin = (int)rec[BC_KMIN];
            kmax = (int)rec[BC_KMAX];
            if (kmax < kmin) kmax = kmin;
        }
        const int row_first = kmin / BB_ROW, row_last = (kmax + BB_ROW - 1) / BB_ROW;
        SFState<NDET> st;
        st.cal = sm_cal + (warp < ns ? warp : 0) * cal_len;
        st.grid = grid;
        st.hh = 0.0;
        st.series = series + (size_t)(slot0 + warp) * ld;
        st.nstore = nstore;
#pragma unroll
        for (int d = 0; d < NDET; ++d) {
            const double f0 = (double)(row_first * BB_ROW + lane) * df;
            double sn, cs;
            sincospi(rec[BC_DET + BC_DSTRIDE * d + 2] * f0, &sn, &cs);
            const double kr = rec[BC_DET + BC_DSTRIDE * d], ki = rec[BC_DET + BC_DSTRIDE * d + 1];
            st.ramp[d][0] = kr * cs + ki * sn;        // conj(K) e^{i theta}
            st.ramp[d][1] = kr * sn - ki * cs;
            st.k2[d] = rec[BC_DET + BC_DSTRIDE * d + 3];
            st.step[d][0] = rec[BC_DET + BC_DSTRIDE * d + 4];
            st.step[d][1] = rec[BC_DET + BC_DSTRIDE * d + 5];
        }
        int ka1 = 0, ka2 = 0, kp1 = 0, kp2 = 0;
        if (APPROX == BB_IMRPHENOMD) {
            ka1 = (int)rec[BC_KA1]; ka2 = (int)rec[BC_KA2]; kp1 = (int)rec[BC_KP1]; kp2 = (int)rec[BC_KP2];
        }

        if (cb1 > cb0 && tid == 0)
            bb_sf_issue_tile<NDET>(sm.tile[issued & 1], &sm.bar[issued & 1], tiles, cb0 * BB_SF_CHUNK);
        for (int cb = cb0; cb < cb1; ++cb) {
            const int stage = issued & 1;
            if (cb + 1 < cb1 && tid == 0)
                bb_sf_issue_tile<NDET>(sm.tile[stage ^ 1], &sm.bar[stage ^ 1], tiles, (cb + 1) * BB_SF_CHUNK);
            if (stage == 0) { bb_mbar_wait(&sm.bar[0], phase0); phase0 ^= 1; }
            else { bb_mbar_wait(&sm.bar[1], phase1); phase1 ^= 1; }
            ++issued;
            const SFTile<NDET>& tile = sm.tile[stage];
            const int c0 = cb * BB_SF_CHUNK;
            int r = max(row_first, c0 / BB_ROW);
            const int rend = min(row_last, (c0 + BB_SF_CHUNK) / BB_ROW);
            if (have && r < rend) {
                if (APPROX == BB_IMRPHENOMD) {
                    while (r < rend) {
                        const int kf = r * BB_ROW;
                        const int ar = kf < ka1 ? 0 : (kf < ka2 ? 1 : 2);
                        const int pr = kf < kp1 ? 0 : (kf < kp2 ? 1 : 2);
                        int nb = INT_MAX;
                        if (ka1 > kf) nb = min(nb, ka1);
                        if (ka2 > kf) nb = min(nb, ka2);
                        if (kp1 > kf) nb = min(nb, kp1);
                        if (kp2 > kf) nb = min(nb, kp2);
                        if (nb < kf + BB_ROW) {
                            bb_sf_rows_generic<NDET, BB_IMRPHENOMD, CAL>(st, tile, rec, r, r + 1, c0, lane, kmin, kmax, df);
                            r += 1;
                        } else {
                            const int rstop = (nb == INT_MAX) ? rend : min(rend, nb / BB_ROW);
                            switch (ar * 3 + pr) {
                                case 0: bb_sf_rows_pd<NDET, 0, 0, CAL>(st, tile, rec, r, rstop, c0, lane, kmin, kmax, df); break;
                                case 3: bb_sf_rows_pd<NDET, 1, 0, CAL>(st, tile, rec, r, rstop, c0, lane, kmin, kmax, df); break;
                                case 4: bb_sf_rows_pd<NDET, 1, 1, CAL>(st, tile, rec, r, rstop, c0, lane, kmin, kmax, df); break;
                                case 5: bb_sf_rows_pd<NDET, 1, 2, CAL>(st, tile, rec, r, rstop, c0, lane, kmin, kmax, df); break;
                                case 8: bb_sf_rows_pd<NDET, 2, 2, CAL>(st, tile, rec, r, rstop, c0, lane, kmin, kmax, df); break;
                                default: bb_sf_rows_generic<NDET, BB_IMRPHENOMD, CAL>(st, tile, rec, r, rstop, c0, lane, kmin, kmax, df); break;
                            }
                            r = rstop;
                        }
                    }
                } else {
                    bb_sf_rows_generic<NDET, APPROX, CAL>(st, tile, rec, r, rend, c0, lane, kmin, kmax, df);
                }
            }
            __syncthreads();
        }
        if (warp < ns) {
            // slot record for K4b: status, kmin, kmax, distance, jitter, <h|h>, sample index, t_c - t_start
            const double hh = bb_warp_sum(st.hh);
            const long s = perm ? (long)perm[p0 + warp] : p0 + warp;
            const double v = lane == 0 ? rec[BC_STATUS] : lane == 1 ? rec[BC_KMIN] : lane == 2 ? rec[BC_KMAX]
                           : lane == 3 ? rec[BC_DISTANCE] : lane == 4 ? rec[BC_JITTER] : lane == 5 ? hh : lane == 6 ? (double)s : rec[BC_DT0];
            if (lane < BB_SF_SLOTREC) slotrec[(size_t)(slot0 + warp) * BB_SF_SLOTREC + lane] = v;
        }
        __syncthreads();
    }
}

// first FFT pass (radix 2^R, stages 0 .. R-1) with the inputs taken from the scratch buffer in global memory: bins
// outside [k0, k1) are zero and were never written
template <int R, int NT>
__device__ __forceinline__ void bb_tm_pass_from_global(double2* X, const double2* __restrict__ src, int k0, int k1,
                                                       int nfft, int ps, const double2* __restrict__ twiddle) {
    constexpr int M = 1 << R;
    const int q = nfft >> R;
    for (int t = threadIdx.x; t < q; t += NT) {
        double2 v[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const int k = t + m * q;
            v[m] = (k >= k0 && k < k1) ? src[k] : make_double2(0.0, 0.0);
        }
        double2 wt = twiddle[t];
#pragma unroll
        for (int st = 0; st < R; ++st) {
            const int half = M >> (st + 1);
            double2 wk[M / 2];
#pragma unroll
            for (int m = 0; m < half; ++m) wk[m] = bb_mul_omega16(wt, (m << st) * (16 / M));
#pragma unroll
            for (int g = 0; g < M; g += 2 * half) {
#pragma unroll
                for (int m = 0; m < half; ++m) {
                    const double2 a = v[g + m], b = v[g + m + half];
                    v[g + m] = make_double2(a.x + b.x, a.y + b.y);
                    v[g + m + half] = bb_cmul(make_double2(a.x - b.x, a.y - b.y), wk[m]);
                }
            }
            wt = bb_cmul(wt, wt);
        }
#pragma unroll
        for (int m = 0; m < M; ++m) X[bb_tm_pos(t + m * q, ps)] = v[m];
    }
    __syncthreads();
}

__device__ __forceinline__ void bb_prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Needs log2n >= 9.  Shared memory: series (padding shift 3), red[32], meta[8], wl[L] (L = nfft / 256).
__global__ void __maxnreg__(BB_SFT_REGS)
bb_series_fft_kernel(long n, int chunk, int n_chunks, int n_slots, const double2* __restrict__ series,
                     const double* __restrict__ slotrec, int nfft, int log2n, const double2* __restrict__ twiddle,
                     BBMarg marg, double start_time, double duration, double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* X = reinterpret_cast<double2*>(smem_raw);
    constexpr int ps = 3;
    const int n_series = (int)bb_tm_series_elems(nfft, ps);
    double* red = reinterpret_cast<double*>(X + n_series);     // [32]
    double* meta = red + 32;                                   // [BB_SF_SLOTREC]
    double2* wl = reinterpret_cast<double2*>(meta + BB_SF_SLOTREC);
    const int tid = threadIdx.x;
    const int L = nfft >> 8;
    for (int e = tid; e < L; e += BB_SFT_THREADS) {
        // exp(-2 pi i e / L) from the nfft-point table (entries m < nfft / 2)
        const double2 w = twiddle[(e & (L / 2 - 1)) << 8];
        wl[e] = (e < L / 2) ? w : make_double2(-w.x, -w.y);
    }

    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const long p = bb_sf_pos(slot, chunk, n_chunks);
        if (p >= n) continue;
        __syncthreads();
        if (tid < BB_SF_SLOTREC) meta[tid] = slotrec[(size_t)slot * BB_SF_SLOTREC + tid];
        if (tid == 32) {
            // pull the next sample's series towards L2 while this one is transformed
            const int next = slot + (int)gridDim.x;
            if (next < n_slots && bb_sf_pos(next, chunk, n_chunks) < n) {
                const double* nr = slotrec + (size_t)next * BB_SF_SLOTREC;
                const int a = max(0, min((int)nr[1], nfft)), b = max(a, min((int)nr[2], nfft));
                if (nr[0] == 0.0 && b > a) bb_prefetch_l2(series + (size_t)next * nfft + a, (unsigned)(b - a) * 16u);
            }
        }
        __syncthreads();
        const long s = (long)meta[6];
        if (meta[0] != 0.0) {
            if (tid == 0) out[s] = -DBL_MAX;
            continue;
        }
        const int k0 = (int)meta[1], k1 = min((int)meta[2], nfft);
        const double2* src = series + (size_t)slot * nfft;
#if BB_SFT_PLAN == 0
        bb_tm_pass_from_global<4, BB_SFT_THREADS>(X, src, k0, k1, nfft, ps, twiddle);
        bb_tm_pass<4, BB_SFT_THREADS>(X, nfft, 4, ps, twiddle);
#else
        bb_tm_pass_from_global<3, BB_SFT_THREADS>(X, src, k0, k1, nfft, ps, twiddle);
        bb_tm_pass<3, BB_SFT_THREADS>(X, nfft, 3, ps, twiddle);
        bb_tm_pass<2, BB_SFT_THREADS>(X, nfft, 6, ps, twiddle);
#endif
        int j_lo, j_hi;
        bb_tm_window(marg, meta[4], start_time, duration, nfft, &j_lo, &j_hi);
        if (j_hi - j_lo <= BB_SFT_PRUNE_MAX) {
            bb_tm_finish<BB_SFT_THREADS, true>(X, nfft, log2n, marg, meta[5], meta[3], meta[4], start_time, duration,
                                               red, out + s, wl, ps);
        } else {
            int st = 8;
            while (log2n - st >= 4 && log2n - st != 5) { bb_tm_pass<4, BB_SFT_THREADS>(X, nfft, st, ps, twiddle); st += 4; }
            if (log2n - st >= 3) { bb_tm_pass<3, BB_SFT_THREADS>(X, nfft, st, ps, twiddle); st += 3; }
            if (log2n - st >= 2) { bb_tm_pass<2, BB_SFT_THREADS>(X, nfft, st, ps, twiddle); st += 2; }
            bb_tm_radix2_tail<BB_SFT_THREADS>(X, nfft, log2n, st, ps, twiddle);
            bb_tm_finish<BB_SFT_THREADS, false>(X, nfft, log2n, marg, meta[5], meta[3], meta[4], start_time, duration,
                                                red, out + s, nullptr, ps);
        }
    }
}

template <int NDET, int APPROX, bool CAL>
static int bb_launch_time_marg_split_t(bb_handle* h, long n, double* out, cudaStream_t st) {
    const int nfft = h->nfft;
    int log2n = 0;
    while ((1 << log2n) < nfft) ++log2n;
    if (log2n < 9) return bb_fail("K4 split: nfft < 512");
    const size_t smem_b = bb_tm_series_elems(nfft, 3) * sizeof(double2) + (32 + BB_SF_SLOTREC) * sizeof(double)
                          + (size_t)(nfft >> 8) * sizeof(double2);
    const size_t smem_a = sizeof(SFSmem<NDET>) + (CAL ? (size_t)BB_SF_SB * NDET * 4 * h->cal.n_points * sizeof(double) : 0);
    if (smem_b > 227 * 1024) return bb_fail("time marginalisation: series does not fit shared memory (nfft > 8192)");
    if (smem_a > 227 * 1024) return bb_fail("K4a: shared memory budget exceeded (too many calibration nodes)");
    BB_CUDA(cudaFuncSetAttribute(bb_series_fill_kernel<NDET, APPROX, CAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    BB_CUDA(cudaFuncSetAttribute(bb_series_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    // both kernels ask for the largest shared-memory carve-out: with the driver's per-kernel default the SM would
    // have to drain and reconfigure between them and the two would never be resident together
    BB_CUDA(cudaFuncSetAttribute(bb_series_fill_kernel<NDET, APPROX, CAL>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 (int)cudaSharedmemCarveoutMaxShared));
    BB_CUDA(cudaFuncSetAttribute(bb_series_fft_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 (int)cudaSharedmemCarveoutMaxShared));

    const long total_blocks = (n + BB_SF_SB - 1) / BB_SF_SB;
    // blocks per chunk: whole multiples of the SM count, large enough that every CTA streams several sample blocks per
    // launch (8 per SM: 8.19 M eval/s at 1e6 samples, 2 per SM: 7.44 M) but at least ~20 chunks so that the first fill and
    // the last transform, which run alone, stay a small part of the batch (1e5 samples: 3 per SM, 19 chunks; three scratch
    // buffers, so that a fill never waits for the transform two chunks back: 7.66 -> 7.8 M eval/s, at the edge of the run-to-run noise)
    long bpc = (total_blocks / 20 + h->sm_count - 1) / h->sm_count * h->sm_count;
    if (bpc < 2L * h->sm_count) bpc = 2L * h->sm_count;
    if (bpc > 8L * h->sm_count) bpc = 8L * h->sm_count;
    { static const long force = [] { const char* e = getenv("BB_K4_BPC"); return e ? atol(e) : 0L; }();      // experiments
      if (force > 0) bpc = force * h->sm_count; }
    const int n_chunks = (int)((total_blocks + bpc - 1) / bpc);
    const size_t slots_cap = (size_t)bpc * BB_SF_SB;
    // scratch buffers in flight: K4a of chunk c may start once K4b of chunk c - nbuf has drained its buffer
    static const int nbuf = [] { const char* e = getenv("BB_K4_NBUF"); const int v = e ? atoi(e) : 0; return (v >= 2 && v <= 4) ? v : 3; }();
    const size_t need = (size_t)nbuf * slots_cap * (size_t)nfft;
    if (need > h->series_cap) {
        cudaFree(h->d_series);
        h->d_series = nullptr;
        h->series_cap = 0;
        BB_CUDA(cudaMalloc(&h->d_series, need * sizeof(double2)));
        h->series_cap = need;
    }
    if ((size_t)nbuf * slots_cap * BB_SF_SLOTREC > h->slotrec_cap) {
        cudaFree(h->d_slotrec);
        h->d_slotrec = nullptr;
        h->slotrec_cap = 0;
        BB_CUDA(cudaMalloc(&h->d_slotrec, (size_t)nbuf * slots_cap * BB_SF_SLOTREC * sizeof(double)));
        BB_CUDA(cudaMemset(h->d_slotrec, 0, (size_t)nbuf * slots_cap * BB_SF_SLOTREC * sizeof(double)));
        h->slotrec_cap = (size_t)nbuf * slots_cap * BB_SF_SLOTREC;
    }
    if (!h->aux) BB_CUDA(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
    while ((int)h->tm_events.size() < 2 * n_chunks) {
        cudaEvent_t e;
        BB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->tm_events.push_back(e);
    }
    const unsigned* perm = h->perm_valid ? h->d_perm : nullptr;
    {
        BBProfScope prof(h, st);
        for (int c = 0; c < n_chunks; ++c) {
            // blocks j with j * n_chunks + c < total_blocks
            const long nb = (total_blocks - 1 - c) / n_chunks + 1;
            const int n_slots = (int)(nb * BB_SF_SB);
            double2* buf = h->d_series + (size_t)(c % nbuf) * slots_cap * nfft;
            double* srec = h->d_slotrec + (size_t)(c % nbuf) * slots_cap * BB_SF_SLOTREC;
            if (c >= nbuf) BB_CUDA(cudaStreamWaitEvent(st, h->tm_events[2 * (c - nbuf) + 1], 0));
            // BB_K4_FILL_SMS (experiment): SMs given to K4a; K4b gets the rest (with kernels that fill an SM each, the
            // fill and the transform then run side by side on disjoint SMs instead of sharing every SM)
            static const int fill_sms = [] { const char* e = getenv("BB_K4_FILL_SMS"); return e ? atoi(e) : 0; }();
            const long sms_a = (fill_sms > 0 && fill_sms < h->sm_count) ? fill_sms : h->sm_count;
            const long sms_b = (fill_sms > 0 && fill_sms < h->sm_count) ? h->sm_count - fill_sms : h->sm_count;
            const unsigned grid_a = (unsigned)(nb < sms_a ? nb : sms_a);
            bb_series_fill_kernel<NDET, APPROX, CAL><<<grid_a, BB_SF_THREADS, smem_a, st>>>(
                h->d_coef, perm, n, c, n_chunks, n_slots, bb_tiles(h), h->net.df, nfft, nfft, h->d_calrec, h->cal, buf, srec);
            BB_CUDA(cudaGetLastError());
            BB_CUDA(cudaEventRecord(h->tm_events[2 * c], st));
            BB_CUDA(cudaStreamWaitEvent(h->aux, h->tm_events[2 * c], 0));
            const unsigned grid_b = (unsigned)(n_slots < sms_b ? n_slots : sms_b);
            bb_series_fft_kernel<<<grid_b, BB_SFT_THREADS, smem_b, h->aux>>>(
                n, c, n_chunks, n_slots, buf, srec, nfft, log2n, h->d_twiddle, h->marg,
                h->net.start_time, h->net.duration, out);
            BB_CUDA(cudaGetLastError());
            BB_CUDA(cudaEventRecord(h->tm_events[2 * c + 1], h->aux));
            h->launches += 2;
        }
        // join: the caller's stream continues after the last FFT chunk (the auxiliary stream is in order)
        BB_CUDA(cudaStreamWaitEvent(st, h->tm_events[2 * (n_chunks - 1) + 1], 0));
    }
    return 0;
}

static int bb_launch_time_marg_split(bb_handle* h, long n, double* out, cudaStream_t st) {
    const bool pd = h->wf.approximant == BB_IMRPHENOMD;
    const bool cal = h->cal_params != nullptr;
#define BB_TMS_CASE(N)                                                                                         \
    case N:                                                                                                    \
        if (cal) return pd ? bb_launch_time_marg_split_t<N, BB_IMRPHENOMD, true>(h, n, out, st)                \
                           : bb_launch_time_marg_split_t<N, BB_TAYLORF2, true>(h, n, out, st);                 \
        return pd ? bb_launch_time_marg_split_t<N, BB_IMRPHENOMD, false>(h, n, out, st)                        \
                  : bb_launch_time_marg_split_t<N, BB_TAYLORF2, false>(h, n, out, st);
    switch (h->net.n_det) {
        BB_TMS_CASE(1)
        BB_TMS_CASE(2)
        BB_TMS_CASE(3)
        BB_TMS_CASE(4)
    }
#undef BB_TMS_CASE
    return bb_fail("bad n_det");
}
