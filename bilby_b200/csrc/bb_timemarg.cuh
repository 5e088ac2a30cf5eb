// K4: time-marginalised likelihood, one CTA per sample (included by bb_kernels.cu).
//
//   series   X[k] = sum_det h_det[k] conj(d[k]) / S[k] (4/T folded into the tiles), k < N-1
//            <- GravitationalWaveTransient.calculate_snrs, bilby/gw/likelihood/base.py:325-330
//               (note h * conj(d), not conj(h) * d, and the dropped Nyquist bin)
//   FFT      d_inner_h_tc_array[j] = sum_k X[k] exp(-2 pi i j k / (N-1))   (numpy.fft.fft), summed over
//            detectors BEFORE the transform (linearity; the reference transforms per detector, base.py:439)
//   weights  times[j] = start_time + (j+1) T/(N-1) (+ time_jitter), kept where inside the geocent_time
//            prior; b = prior.prob * delta_tc                         <- base.py:794-806, 1027-1035
//   reduce   logsumexp_j( lnl_j, b )  with lnl_j = plain | phase-marginalised | distance(-phase)-marginalised
//            point likelihood                                          <- base.py:808-820
//
// The series lives in shared memory (18 B * nfft with the padding, 147 KB at 8 s / 2048 Hz); 16 warps fill it
// with the region-specialised evaluators of K1 (rows of 32 bins, phase ramps by recurrence), then a radix-8
// decimation-in-frequency FFT runs in place; the logsumexp reads the bit-reversed outputs inside the prior only.
#pragma once

#define BB_TM_THREADS 512
#define BB_TM_WARPS (BB_TM_THREADS / 32)

__device__ __forceinline__ unsigned bb_bitrev(unsigned v, int bits) { return __brev(v) >> (32 - bits); }

// The series lives in shared memory with one 16-byte pad after every 8 elements: with it every access pattern of
// the radix-8 passes below (element stride q = nfft/8, nfft/64, ... 1) and the natural-order fill is free of bank
// conflicts.
__device__ __forceinline__ int bb_tm_pos(int i) { return i + (i >> 3); }
__host__ __device__ inline size_t bb_tm_series_elems(int nfft) { return (size_t)nfft + ((size_t)nfft >> 3) + 1; }

__device__ __forceinline__ double2 bb_cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place decimation-in-frequency FFT (natural order in, bit-reversed order out: output j is at index
// bitrev(j)).  Three radix-2 stages at a time are fused into one radix-8 pass held in registers (8 elements of
// stride q per thread, one twiddle load per thread and pass: the other six follow from w^2, w^4 and the eighth
// roots of unity); the remaining one or two stages run as plain radix-2 passes.
__device__ __forceinline__ void bb_tm_fft_dif(double2* X, int nfft, int log2n, const double2* __restrict__ twiddle) {
    const int tid = threadIdx.x;
    const double r = 0.70710678118654752440;
    int s = 0;
    for (; s + 3 <= log2n; s += 3) {
        const int q = nfft >> (s + 3);
        for (int t = tid; t < (nfft >> 3); t += BB_TM_THREADS) {
            const int blk = t / q, jp = t - blk * q;
            const int base = blk * 8 * q + jp;
            double2 v[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = X[bb_tm_pos(base + m * q)];
            const double2 w1 = twiddle[jp << s];
            const double2 w2 = bb_cmul(w1, w1), w4 = bb_cmul(w2, w2);
            // stage s: (m, m+4), twiddle w1 * w8^m
            const double2 t1 = make_double2(r * (w1.x + w1.y), r * (w1.y - w1.x));      // w1 * (1 - i)/sqrt2
            const double2 t2 = make_double2(w1.y, -w1.x);                                // w1 * (-i)
            const double2 t3 = make_double2(r * (w1.y - w1.x), -r * (w1.x + w1.y));     // w1 * (-1 - i)/sqrt2
            const double2 ws[4] = {w1, t1, t2, t3};
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const double2 a = v[m], b = v[m + 4];
                v[m] = make_double2(a.x + b.x, a.y + b.y);
                v[m + 4] = bb_cmul(make_double2(a.x - b.x, a.y - b.y), ws[m]);
            }
            // stage s+1: (m, m+2) inside each half, twiddle w2 * (-i)^(m&1)
            const double2 w2i = make_double2(w2.y, -w2.x);
#pragma unroll
            for (int hb = 0; hb < 8; hb += 4) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const double2 a = v[hb + m], b = v[hb + m + 2];
                    v[hb + m] = make_double2(a.x + b.x, a.y + b.y);
                    v[hb + m + 2] = bb_cmul(make_double2(a.x - b.x, a.y - b.y), m ? w2i : w2);
                }
            }
            // stage s+2: (m, m+1), twiddle w4
#pragma unroll
            for (int m = 0; m < 8; m += 2) {
                const double2 a = v[m], b = v[m + 1];
                v[m] = make_double2(a.x + b.x, a.y + b.y);
                v[m + 1] = bb_cmul(make_double2(a.x - b.x, a.y - b.y), w4);
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) X[bb_tm_pos(base + m * q)] = v[m];
        }
        __syncthreads();
    }
    for (; s < log2n; ++s) {
        const int h = nfft >> (s + 1);
        for (int b = tid; b < (nfft >> 1); b += BB_TM_THREADS) {
            const int blk = b / h, j = b - blk * h;
            const int i0 = blk * 2 * h + j;
            const double2 w = twiddle[j << s];
            const double2 a = X[bb_tm_pos(i0)], c = X[bb_tm_pos(i0 + h)];
            X[bb_tm_pos(i0)] = make_double2(a.x + c.x, a.y + c.y);
            X[bb_tm_pos(i0 + h)] = bb_cmul(make_double2(a.x - c.x, a.y - c.y), w);
        }
        __syncthreads();
    }
}

// FFT of the series in shared memory followed by the weighted logsumexp over the times inside the geocent_time
// prior; shared by the full-grid and the relative-binning time-marginalised kernels.  Must be entered with the
// series complete (after a __syncthreads()).
__device__ __forceinline__ void bb_tm_finish(double2* X, int nfft, int log2n, const double2* __restrict__ twiddle,
                                             const BBMarg& marg, double hh, double dist, double jitter,
                                             double start_time, double duration, double* red, double* out_s) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bb_tm_fft_dif(X, nfft, log2n, twiddle);

    // weighted logsumexp over the times inside the prior
    const double dtc = duration / (double)nfft;     // = 2 / sampling_frequency
    const double jit = marg.jitter ? jitter : 0.0;
    const double bw = dtc / (marg.time_max - marg.time_min);
    // times = start_time + linspace(0, T, nfft + 1)[1:] (+ jitter): only j with times inside the prior
    const int j_lo = (int)fmin(fmax(floor((marg.time_min - jit - start_time) / dtc) - 2.0, 0.0), (double)nfft);
    const int j_hi = (int)fmin(fmax(ceil((marg.time_max - jit - start_time) / dtc) + 1.0, 0.0), (double)nfft);
    double mx = -INFINITY, sum = 0.0;
    for (int j = j_lo + tid; j < j_hi; j += BB_TM_THREADS) {
        const double tj = (start_time + (double)(j + 1) * dtc) + jit;
        if (tj < marg.time_min || tj > marg.time_max) continue;
        const double2 v = X[bb_tm_pos((int)bb_bitrev((unsigned)j, log2n))];
        const double l = bb_point_lnl(marg, v.x, v.y, hh, dist);
        if (l == -INFINITY) continue;
        if (l > mx) { sum = sum * exp(mx - l) + bw; mx = l; }
        else sum += bw * exp(l - mx);
    }
    double gmx = mx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gmx = fmax(gmx, __shfl_xor_sync(0xffffffffu, gmx, o));
    if (lane == 0) red[warp] = gmx;
    __syncthreads();
    gmx = red[0];
    for (int w = 1; w < BB_TM_WARPS; ++w) gmx = fmax(gmx, red[w]);
    __syncthreads();
    double part = (mx == -INFINITY) ? 0.0 : sum * exp(mx - gmx);
    part = bb_warp_sum(part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < BB_TM_WARPS; ++w) tot += red[w];
        *out_s = (gmx == -INFINITY) ? -INFINITY : log(tot) + gmx;
    }
}

// per-warp running state of the sample the CTA is working on
template <int NDET>
struct TMState {
    double ramp[NDET][2];    // exp(+2 pi i f dt_d) at this lane's bin of the current row
    double step[NDET][2];    // its advance over BB_TM_WARPS rows
    double hh;
    const double* cal;
    BBCalGrid grid;
};

template <int NDET, bool CAL>
__device__ __forceinline__ void bb_tm_bin(TMState<NDET>& st, const BBTiles& g, const double* rec, double2* X, int k,
                                          bool act, int nfft, double A, double ph, double lfk) {
    double sn, cs;
    sincospi(act ? ph : 0.0, &sn, &cs);
    A = act ? A : 0.0;
    const double zr = A * cs, zi = A * sn;      // conj(h22 incl. geocentric shift)
    const double A2 = A * A;
    double vr = 0.0, vi = 0.0;
#pragma unroll
    for (int d = 0; d < NDET; ++d) {
        const double rc = st.ramp[d][0], rs = st.ramp[d][1];
        double wr = zr * rc - zi * rs, wi = zr * rs + zi * rc;
        double hw = A2;
        if (CAL) {
            double amp1, cr, ci;
            bb_cal_factor(st.cal + d * 4 * st.grid.n_points, st.grid.n_points, st.grid.l0[d], st.grid.inv_delta[d],
                          lfk, &amp1, &cr, &ci);
            const double tr = amp1 * (wr * cr + wi * ci), ti = amp1 * (wi * cr - wr * ci);
            wr = tr;
            wi = ti;
            hw = A2 * amp1 * amp1;
        }
        const double2 dd = g.ds[(size_t)d * g.n_pad + k];
        const double pr = wr * dd.x - wi * dd.y, pi = wr * dd.y + wi * dd.x;     // conj(h/K) d/S
        const double kr = rec[BC_DET + BC_DSTRIDE * d], ki = rec[BC_DET + BC_DSTRIDE * d + 1];
        vr += kr * pr + ki * pi;                                                 // conj(K) p
        vi += kr * pi - ki * pr;
        st.hh += rec[BC_DET + BC_DSTRIDE * d + 3] * hw * g.is[(size_t)d * g.n_pad + k];
        st.ramp[d][0] = rc * st.step[d][0] - rs * st.step[d][1];
        st.ramp[d][1] = rc * st.step[d][1] + rs * st.step[d][0];
    }
    if (act && k < nfft) X[bb_tm_pos(k)] = make_double2(vr, -vi);     // h conj(d)/S = conj(conj(h) d/S)
}

// this warp's rows r, r + BB_TM_WARPS, ... < rstop, all inside amplitude region AR and phase region PR
template <int NDET, int AR, int PR, bool CAL>
__device__ __forceinline__ int bb_tm_rows_pd(TMState<NDET>& st, const BBTiles& g, const double* rec, double2* X, int r,
                                             int rstop, int lane, int kmin, int kmax, int nfft, double df) {
    K1Amp<AR> amp;
    K1Ph<PR> phs;
    amp.load(rec);
    phs.load(rec);
    const double a0 = rec[BC_A0];
    for (; r < rstop; r += BB_TM_WARPS) {
        const int k = r * BB_ROW + lane;
        const bool act = (k >= kmin) && (k < kmax);
        const double f = (double)k * df;
        const double u = g.u[k], t = u * u, x = f * t * t;
        const double lfk = g.lf[k];
        const double A = amp.eval(f, x) * a0 * (u * (t * t * t));
        const double ph = phs.eval(f, t, x, lfk, g.q34[k]);
        bb_tm_bin<NDET, CAL>(st, g, rec, X, k, act, nfft, A, ph, lfk);
    }
    return r;
}

template <int NDET, int APPROX, bool CAL>
__device__ __forceinline__ void bb_tm_row_generic(TMState<NDET>& st, const BBTiles& g, const double* rec, double2* X,
                                                  int r, int lane, int kmin, int kmax, int nfft, double df) {
    const int k = r * BB_ROW + lane;
    const bool act = (k >= kmin) && (k < kmax);
    const double f = (double)k * df;
    double A, ph;
    const double lfk = g.lf[k];
    bb_wave<APPROX>(rec, f, g.u[k], lfk, g.q34[k], &A, &ph);
    bb_tm_bin<NDET, CAL>(st, g, rec, X, k, act, nfft, A, ph, lfk);
}

template <int NDET, int APPROX, bool CAL>
__global__ void __launch_bounds__(BB_TM_THREADS, 1)
bb_time_marg_kernel(const double* __restrict__ coef, long n, BBTiles tiles, int n_freq, double df, int nfft,
                    int log2n, const double2* __restrict__ twiddle, BBMarg marg, double start_time,
                    double duration, const double* __restrict__ calrec, BBCalGrid grid, double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* X = reinterpret_cast<double2*>(smem_raw);
    const int n_series = (int)bb_tm_series_elems(nfft);
    double* c = reinterpret_cast<double*>(X + n_series);
    double* red = c + BC_NCOEF;      // [32]
    double* cal = red + 32;          // CAL: [NDET][4][n_points]
    const int cal_len = CAL ? NDET * 4 * grid.n_points : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (long s = blockIdx.x; s < n; s += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < BC_NCOEF; i += BB_TM_THREADS) c[i] = coef[s * BC_NCOEF + i];
        for (int i = tid; i < n_series; i += BB_TM_THREADS) X[i] = make_double2(0.0, 0.0);
        if (CAL) for (int i = tid; i < cal_len; i += BB_TM_THREADS) cal[i] = calrec[s * cal_len + i];
        __syncthreads();
        if (c[BC_STATUS] != 0.0) {
            if (tid == 0) out[s] = -DBL_MAX;
            continue;
        }
        // rows of 32 bins; warp w owns rows r = w (mod BB_TM_WARPS).  Nyquist bin: in <h|h>, not in the series.
        const int kmin = (int)c[BC_KMIN], kmax = (int)c[BC_KMAX];
        const int row_first = kmin / BB_ROW, row_last = (kmax + BB_ROW - 1) / BB_ROW;
        int r = row_first + ((warp - row_first) % BB_TM_WARPS + BB_TM_WARPS) % BB_TM_WARPS;
        TMState<NDET> st;
        st.hh = 0.0;
        st.cal = cal;
        st.grid = grid;
#pragma unroll
        for (int d = 0; d < NDET; ++d) {
            const double two_dt = c[BC_DET + BC_DSTRIDE * d + 2];
            sincospi(two_dt * ((double)(r * BB_ROW + lane) * df), &st.ramp[d][1], &st.ramp[d][0]);
            sincospi(two_dt * ((double)(BB_TM_WARPS * BB_ROW) * df), &st.step[d][1], &st.step[d][0]);
        }
        if (APPROX == BB_IMRPHENOMD) {
            const int ka1 = (int)c[BC_KA1], ka2 = (int)c[BC_KA2], kp1 = (int)c[BC_KP1], kp2 = (int)c[BC_KP2];
            while (r < row_last) {
                const int kf = r * BB_ROW;
                const int ar = kf < ka1 ? 0 : (kf < ka2 ? 1 : 2);
                const int pr = kf < kp1 ? 0 : (kf < kp2 ? 1 : 2);
                int nb = INT_MAX;
                if (ka1 > kf) nb = min(nb, ka1);
                if (ka2 > kf) nb = min(nb, ka2);
                if (kp1 > kf) nb = min(nb, kp1);
                if (kp2 > kf) nb = min(nb, kp2);
                if (nb < kf + BB_ROW) {
                    bb_tm_row_generic<NDET, BB_IMRPHENOMD, CAL>(st, tiles, c, X, r, lane, kmin, kmax, nfft, df);
                    r += BB_TM_WARPS;
                } else {
                    const int rstop = (nb == INT_MAX) ? row_last : min(row_last, nb / BB_ROW);
                    switch (ar * 3 + pr) {
                        case 0: r = bb_tm_rows_pd<NDET, 0, 0, CAL>(st, tiles, c, X, r, rstop, lane, kmin, kmax, nfft, df); break;
                        case 3: r = bb_tm_rows_pd<NDET, 1, 0, CAL>(st, tiles, c, X, r, rstop, lane, kmin, kmax, nfft, df); break;
                        case 4: r = bb_tm_rows_pd<NDET, 1, 1, CAL>(st, tiles, c, X, r, rstop, lane, kmin, kmax, nfft, df); break;
                        case 5: r = bb_tm_rows_pd<NDET, 1, 2, CAL>(st, tiles, c, X, r, rstop, lane, kmin, kmax, nfft, df); break;
                        case 8: r = bb_tm_rows_pd<NDET, 2, 2, CAL>(st, tiles, c, X, r, rstop, lane, kmin, kmax, nfft, df); break;
                        default:
                            for (; r < rstop; r += BB_TM_WARPS)
                                bb_tm_row_generic<NDET, BB_IMRPHENOMD, CAL>(st, tiles, c, X, r, lane, kmin, kmax, nfft, df);
                            break;
                    }
                }
            }
        } else {
            for (; r < row_last; r += BB_TM_WARPS)
                bb_tm_row_generic<NDET, APPROX, CAL>(st, tiles, c, X, r, lane, kmin, kmax, nfft, df);
        }
        // block-reduce <h|h>
        double hh = bb_warp_sum(st.hh);
        if (lane == 0) red[warp] = hh;
        __syncthreads();
        hh = 0.0;
        for (int w = 0; w < BB_TM_WARPS; ++w) hh += red[w];
        __syncthreads();
        bb_tm_finish(X, nfft, log2n, twiddle, marg, hh, c[BC_DISTANCE], c[BC_JITTER], start_time, duration, red, out + s);
    }
}

template <int NDET, int APPROX, bool CAL>
static int bb_launch_time_marg_t(bb_handle* h, long n, double* out, cudaStream_t st) {
    const int nfft = h->nfft;
    int log2n = 0;
    while ((1 << log2n) < nfft) ++log2n;
    const size_t smem = bb_tm_series_elems(nfft) * sizeof(double2) + (BC_NCOEF + 32 + (CAL ? NDET * 4 * h->cal.n_points : 0)) * sizeof(double);
    if (smem > 227 * 1024) return bb_fail("time marginalisation: series does not fit shared memory (nfft > 8192)");
    BB_CUDA(cudaFuncSetAttribute(bb_time_marg_kernel<NDET, APPROX, CAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    long grid = (long)h->sm_count * per_sm;
    if (grid > n) grid = n;
    {
        BBProfScope prof(h, st);
        bb_time_marg_kernel<NDET, APPROX, CAL><<<(unsigned)grid, BB_TM_THREADS, smem, st>>>(
            h->d_coef, n, bb_tiles(h), h->net.n_freq, h->net.df, nfft, log2n, h->d_twiddle, h->marg,
            h->net.start_time, h->net.duration, h->d_calrec, h->cal, out);
    }
    h->launches++;
    BB_CUDA(cudaGetLastError());
    return 0;
}

static int bb_launch_time_marg(bb_handle* h, long n, double* out, cudaStream_t st) {
    if (h->nfft == 0) return bb_fail("time marginalisation needs n_freq - 1 to be a power of two");
    if (h->shard_lo != 0 || h->shard_hi != h->net.n_freq)
        return bb_fail("time marginalisation cannot be frequency-sharded (SURVEY.md section 8e)");
    const bool pd = h->wf.approximant == BB_IMRPHENOMD;
    const bool cal = h->cal_params != nullptr;
#define BB_TM_CASE(N)                                                                                          \
    case N:                                                                                                    \
        if (cal) return pd ? bb_launch_time_marg_t<N, BB_IMRPHENOMD, true>(h, n, out, st)                      \
                           : bb_launch_time_marg_t<N, BB_TAYLORF2, true>(h, n, out, st);                       \
        return pd ? bb_launch_time_marg_t<N, BB_IMRPHENOMD, false>(h, n, out, st)                              \
                  : bb_launch_time_marg_t<N, BB_TAYLORF2, false>(h, n, out, st);
    switch (h->net.n_det) {
        BB_TM_CASE(1)
        BB_TM_CASE(2)
        BB_TM_CASE(3)
        BB_TM_CASE(4)
    }
#undef BB_TM_CASE
    return bb_fail("bad n_det");
}
