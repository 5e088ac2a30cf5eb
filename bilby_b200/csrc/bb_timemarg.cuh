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

// The series lives in shared memory with one 16-byte pad after every 2^PS elements (PS = 3, or 4 when every pass
// is radix-16): with it every access pattern of the passes below (element stride q = nfft/16, ... 1) and the
// natural-order fill are free of bank conflicts.
__device__ __forceinline__ int bb_tm_pos(int i, int ps) { return i + (i >> ps); }
__host__ __device__ inline size_t bb_tm_series_elems(int nfft, int ps) { return (size_t)nfft + ((size_t)nfft >> ps) + 1; }
// pass plan: a radix-16 passes followed by b radix-8 passes with 4a + 3b = log2n (a as large as possible); any
// remainder runs as radix-2 stages
__host__ __device__ inline void bb_tm_plan(int log2n, int* a, int* b, int* ps) {
    *a = 0;
    *b = 0;
    for (int aa = log2n / 4; aa >= 0; --aa)
        if ((log2n - 4 * aa) % 3 == 0) { *a = aa; *b = (log2n - 4 * aa) / 3; break; }
    *ps = (*b == 0 && *a > 0) ? 4 : 3;
}

__device__ __forceinline__ double2 bb_cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// x * exp(-2 pi i k / 16); k is a compile-time constant after unrolling
__device__ __forceinline__ double2 bb_mul_omega16(double2 x, int k) {
    const double r = 0.70710678118654752440, c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
    switch (k & 15) {
        case 0: return x;
        case 1: return make_double2(x.x * c1 + x.y * s1, x.y * c1 - x.x * s1);
        case 2: return make_double2(r * (x.x + x.y), r * (x.y - x.x));
        case 3: return make_double2(x.x * s1 + x.y * c1, x.y * s1 - x.x * c1);
        case 4: return make_double2(x.y, -x.x);
        case 5: return make_double2(x.y * c1 - x.x * s1, -x.x * c1 - x.y * s1);
        case 6: return make_double2(r * (x.y - x.x), -r * (x.x + x.y));
        case 7: return make_double2(x.y * s1 - x.x * c1, -x.x * s1 - x.y * c1);
        default: return make_double2(-x.x, -x.y);
    }
}

// one radix-2^R decimation-in-frequency pass (R fused radix-2 stages s .. s+R-1) held in registers: 2^R elements of
// stride q per thread, one twiddle load per thread (the others follow by squaring and by the 16th roots of unity)
template <int R, int NT = BB_TM_THREADS>
__device__ __forceinline__ void bb_tm_pass(double2* X, int nfft, int s, int ps, const double2* __restrict__ twiddle) {
    constexpr int M = 1 << R;
    const int q = nfft >> (s + R);
    for (int t = threadIdx.x; t < (nfft >> R); t += NT) {
        const int blk = t / q, jp = t - blk * q;
        const int base = blk * M * q + jp;
        double2 v[M];
#pragma unroll
        for (int m = 0; m < M; ++m) v[m] = X[bb_tm_pos(base + m * q, ps)];
        double2 wt = twiddle[jp << s];
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
        for (int m = 0; m < M; ++m) X[bb_tm_pos(base + m * q, ps)] = v[m];
    }
    __syncthreads();
}

// radix-2 stages s .. log2n-1 (whatever the radix-16 / radix-8 plan leaves over)
template <int NT = BB_TM_THREADS>
__device__ __forceinline__ void bb_tm_radix2_tail(double2* X, int nfft, int log2n, int s, int ps,
                                                  const double2* __restrict__ twiddle) {
    for (; s < log2n; ++s) {
        const int h = nfft >> (s + 1);
        for (int bb = threadIdx.x; bb < (nfft >> 1); bb += NT) {
            const int blk = bb / h, j = bb - blk * h;
            const int i0 = blk * 2 * h + j;
            const double2 w = twiddle[j << s];
            const double2 x0 = X[bb_tm_pos(i0, ps)], x1 = X[bb_tm_pos(i0 + h, ps)];
            X[bb_tm_pos(i0, ps)] = make_double2(x0.x + x1.x, x0.y + x1.y);
            X[bb_tm_pos(i0 + h, ps)] = bb_cmul(make_double2(x0.x - x1.x, x0.y - x1.y), w);
        }
        __syncthreads();
    }
}

// In-place decimation-in-frequency FFT (natural order in, bit-reversed order out: output j is at index bitrev(j)).
template <int NT = BB_TM_THREADS>
__device__ __forceinline__ void bb_tm_fft_dif(double2* X, int nfft, int log2n, const double2* __restrict__ twiddle) {
    int a, b, ps;
    bb_tm_plan(log2n, &a, &b, &ps);
    int s = 0;
    for (int i = 0; i < a; ++i, s += 4) bb_tm_pass<4, NT>(X, nfft, s, ps, twiddle);
    for (int i = 0; i < b; ++i, s += 3) bb_tm_pass<3, NT>(X, nfft, s, ps, twiddle);
    bb_tm_radix2_tail<NT>(X, nfft, log2n, s, ps, twiddle);
}

// After the FFT of the series in shared memory: the weighted logsumexp over the times inside the geocent_time
// prior; shared by the full-grid and the relative-binning time-marginalised kernels.  Must be entered with the
// series complete (after a __syncthreads()).
// Output j of the transform after only the first 8 radix-2 stages: block c = bitrev8(j & 255) of L = nfft / 256
// consecutive elements holds the sequence whose L-point DFT gives the outputs (j & 255) + 256 n2;
// wl[e] = exp(-2 pi i e / L) in shared memory.
__device__ __forceinline__ double2 bb_tm_pruned_value(const double2* X, int j, int log2n, int ps, const double2* wl) {
    const int logl = log2n - 8, L = 1 << logl;
    const int base = (int)(__brev((unsigned)(j & 255)) >> 24) << logl;
    const int n2 = j >> 8;
    double vr = 0.0, vi = 0.0;
    // start at a lane-dependent element so that neighbouring outputs do not hit the same banks
    const int rot = (j * 5) & (L - 1);
    for (int i = 0; i < L; ++i) {
        const int jp = (i + rot) & (L - 1);
        const double2 z = X[bb_tm_pos(base + jp, ps)];
        const double2 w = wl[(jp * n2) & (L - 1)];
        vr += z.x * w.x - z.y * w.y;
        vi += z.x * w.y + z.y * w.x;
    }
    return make_double2(vr, vi);
}

// times = start_time + linspace(0, T, nfft + 1)[1:] (+ jitter): the range [j_lo, j_hi) that can lie inside the prior
__device__ __forceinline__ void bb_tm_window(const BBMarg& marg, double jitter, double start_time, double duration,
                                             int nfft, int* j_lo, int* j_hi) {
    const double dtc = duration / (double)nfft;     // = 2 / sampling_frequency
    const double jit = marg.jitter ? jitter : 0.0;
    *j_lo = (int)fmin(fmax(floor((marg.time_min - jit - start_time) / dtc) - 2.0, 0.0), (double)nfft);
    *j_hi = (int)fmin(fmax(ceil((marg.time_max - jit - start_time) / dtc) + 1.0, 0.0), (double)nfft);
}

// PRUNED = false: the FFT is complete (output j at index bitrev(j)).
// PRUNED = true: only the first 8 radix-2 stages ran; block c = bitrev8(j & 255) of L = nfft / 256 consecutive
// elements holds the sequence whose L-point DFT gives the outputs j = (j & 255) + 256 n2, and the few outputs inside
// the prior are summed directly (wl[e] = exp(-2 pi i e / L) in shared memory).
template <int NT = BB_TM_THREADS, bool PRUNED = false>
__device__ __forceinline__ void bb_tm_finish(double2* X, int nfft, int log2n, const BBMarg& marg, double hh, double dist, double jitter,
                                             double start_time, double duration, double* red, double* out_s,
                                             const double2* wl = nullptr, int ps_in = -1) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int pa, pb, ps;
    bb_tm_plan(log2n, &pa, &pb, &ps);
    if (ps_in >= 0) ps = ps_in;

    // weighted logsumexp over the times inside the prior
    const double dtc = duration / (double)nfft;     // = 2 / sampling_frequency
    const double jit = marg.jitter ? jitter : 0.0;
    const double bw = dtc / (marg.time_max - marg.time_min);
    int j_lo, j_hi;
    bb_tm_window(marg, jitter, start_time, duration, nfft, &j_lo, &j_hi);
    double mx = -INFINITY, sum = 0.0;
    for (int j = j_lo + tid; j < j_hi; j += NT) {
        const double tj = (start_time + (double)(j + 1) * dtc) + jit;
        if (tj < marg.time_min || tj > marg.time_max) continue;
        double2 v;
        if (PRUNED) {
            v = bb_tm_pruned_value(X, j, log2n, ps, wl);
        } else {
            v = X[bb_tm_pos((int)bb_bitrev((unsigned)j, log2n), ps)];
        }
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
    for (int w = 1; w < NT / 32; ++w) gmx = fmax(gmx, red[w]);
    __syncthreads();
    double part = (mx == -INFINITY) ? 0.0 : sum * exp(mx - gmx);
    part = bb_warp_sum(part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < NT / 32; ++w) tot += red[w];
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
    int ps;
};

template <int NDET, bool CAL>
__device__ __forceinline__ void bb_tm_bin(TMState<NDET>& st, const BBTiles& g, const double* rec, double2* X, int k,
                                          bool act, int nfft, double A, double ph, double lfk) {
    double sn, cs;
    bb_sincospi(act ? ph : 0.0, &sn, &cs);
    A = act ? A : 0.0;
    const double zr = A * cs, zi = A * sn;      // conj(h22 incl. geocentric shift)
    const double A2 = A * A;
    double vr = 0.0, vi = 0.0;
    BBCalW cw;
    if (CAL) cw = bb_cal_weights(st.grid.n_points, st.grid.l0[0], st.grid.inv_delta[0], lfk);
#pragma unroll
    for (int d = 0; d < NDET; ++d) {
        const double rc = st.ramp[d][0], rs = st.ramp[d][1];
        double wr = zr * rc - zi * rs, wi = zr * rs + zi * rc;
        double hw = A2;
        if (CAL) {
            double amp1, cr, ci;
            if (d > 0 && !st.grid.shared) cw = bb_cal_weights(st.grid.n_points, st.grid.l0[d], st.grid.inv_delta[d], lfk);
            bb_cal_apply(st.cal + d * 4 * st.grid.n_points, st.grid.n_points, cw, &amp1, &cr, &ci);
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
    if (act && k < nfft) X[bb_tm_pos(k, st.ps)] = make_double2(vr, -vi);     // h conj(d)/S = conj(conj(h) d/S)
}

// this warp's rows r, r + BB_TM_WARPS, ... < rstop, all inside amplitude region AR and phase region PR
template <int NDET, int AR, int PR, bool CAL>
__device__ __forceinline__ int bb_tm_rows_pd(TMState<NDET>& st, const BBTiles& g, const double* rec, double2* X, int r,
                                             int rstop, int lane, int kmin, int kmax, int nfft, double df) {
    K1Amp<AR> amp;
    K1Ph<PR> phs;
    amp.load(rec);
    phs.load(rec);
    amp.begin((double)(r * BB_ROW + lane) * df, (double)(BB_TM_WARPS * BB_ROW) * df);
    for (; r < rstop; r += BB_TM_WARPS) {
        const int k = r * BB_ROW + lane;
        const bool act = (k >= kmin) && (k < kmax);
        const double f = (double)k * df;
        const double u = g.u[k], t = u * u, x = f * t * t;
        const double lfk = g.lf[k];
        const double t3 = t * t * t;                    // 1 / f
        const double A = amp.eval(f, x) * (u * t3);     // a0 is folded into the region's coefficients (bb_k1.cuh)
        const double ph = phs.eval(f, t, x, t3, lfk, g.q34[k]);
        bb_tm_bin<NDET, CAL>(st, g, rec, X, k, act, nfft, A, ph, lfk);
        amp.next();
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
                    double duration, const double* __restrict__ calrec, BBCalGrid grid, double* __restrict__ out,
                    long long* __restrict__ phase_clk) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* X = reinterpret_cast<double2*>(smem_raw);
    int plan_a, plan_b, ps;
    bb_tm_plan(log2n, &plan_a, &plan_b, &ps);
    const int n_series = (int)bb_tm_series_elems(nfft, ps);
    double* c = reinterpret_cast<double*>(X + n_series);
    double* red = c + BC_NCOEF;      // [32]
    double* stepbuf = red + 32;      // [2 * BB_MAX_DET] ramp advance over BB_TM_WARPS rows
    double* cal = stepbuf + 2 * BB_MAX_DET;          // CAL: [NDET][4][n_points]
    const int cal_len = CAL ? NDET * 4 * grid.n_points : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // BB_TM_PHASES=1 (debug): cycles per phase, summed over samples by thread 0 of every CTA
    long long pclk[4] = {0, 0, 0, 0}, pt = 0;
#define BB_TM_MARK(i)                                                                  \
    if (phase_clk != nullptr && tid == 0) {                                            \
        const long long now = clock64();                                               \
        pclk[i] += now - pt;                                                           \
        pt = now;                                                                      \
    }
    for (long s = blockIdx.x; s < n; s += gridDim.x) {
        __syncthreads();
        if (phase_clk != nullptr && tid == 0) pt = clock64();
        for (int i = tid; i < BC_NCOEF; i += BB_TM_THREADS) c[i] = coef[s * BC_NCOEF + i];
        if (CAL) for (int i = tid; i < cal_len; i += BB_TM_THREADS) cal[i] = calrec[s * cal_len + i];
        __syncthreads();
        if (c[BC_STATUS] != 0.0) {
            if (tid == 0) out[s] = -DBL_MAX;
            continue;
        }
        {
            // zero the series outside the active range (the fill below writes every active bin)
            const int z0 = min((int)c[BC_KMIN], nfft), z1 = min((int)c[BC_KMAX], nfft);
            for (int i = tid; i < z0; i += BB_TM_THREADS) X[bb_tm_pos(i, ps)] = make_double2(0.0, 0.0);
            for (int i = z1 + tid; i < nfft; i += BB_TM_THREADS) X[bb_tm_pos(i, ps)] = make_double2(0.0, 0.0);
            if (tid < NDET)
                sincospi(c[BC_DET + BC_DSTRIDE * tid + 2] * ((double)(BB_TM_WARPS * BB_ROW) * df),
                         &stepbuf[2 * tid + 1], &stepbuf[2 * tid]);
            __syncthreads();
        }
        BB_TM_MARK(0)
        // rows of 32 bins; warp w owns rows r = w (mod BB_TM_WARPS).  Nyquist bin: in <h|h>, not in the series.
        const int kmin = (int)c[BC_KMIN], kmax = (int)c[BC_KMAX];
        const int row_first = kmin / BB_ROW, row_last = (kmax + BB_ROW - 1) / BB_ROW;
        int r = row_first + ((warp - row_first) % BB_TM_WARPS + BB_TM_WARPS) % BB_TM_WARPS;
        TMState<NDET> st;
        st.hh = 0.0;
        st.cal = cal;
        st.grid = grid;
        st.ps = ps;
#pragma unroll
        for (int d = 0; d < NDET; ++d) {
            const double two_dt = c[BC_DET + BC_DSTRIDE * d + 2];
            sincospi(two_dt * ((double)(r * BB_ROW + lane) * df), &st.ramp[d][1], &st.ramp[d][0]);
            st.step[d][0] = stepbuf[2 * d];
            st.step[d][1] = stepbuf[2 * d + 1];
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
        BB_TM_MARK(1)
        bb_tm_fft_dif(X, nfft, log2n, twiddle);
        BB_TM_MARK(2)
        bb_tm_finish(X, nfft, log2n, marg, hh, c[BC_DISTANCE], c[BC_JITTER], start_time, duration, red, out + s);
        BB_TM_MARK(3)
    }
    if (phase_clk != nullptr && tid == 0)
        for (int i = 0; i < 4; ++i) atomicAdd((unsigned long long*)phase_clk + i, (unsigned long long)pclk[i]);
#undef BB_TM_MARK
}

static int bb_launch_time_marg_split(bb_handle* h, long n, double* out, cudaStream_t st);

template <int NDET, int APPROX, bool CAL>
static int bb_launch_time_marg_t(bb_handle* h, long n, double* out, cudaStream_t st) {
    const int nfft = h->nfft;
    int log2n = 0;
    while ((1 << log2n) < nfft) ++log2n;
    int plan_a, plan_b, ps;
    bb_tm_plan(log2n, &plan_a, &plan_b, &ps);
    const size_t smem = bb_tm_series_elems(nfft, ps) * sizeof(double2)
                        + (BC_NCOEF + 32 + 2 * BB_MAX_DET + (CAL ? NDET * 4 * h->cal.n_points : 0)) * sizeof(double);
    if (smem > 227 * 1024) return bb_fail("time marginalisation: series does not fit shared memory (nfft > 8192)");
    BB_CUDA(cudaFuncSetAttribute(bb_time_marg_kernel<NDET, APPROX, CAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    long grid = (long)h->sm_count * per_sm;
    if (grid > n) grid = n;
    long long* d_phase = nullptr;
    if (getenv("BB_TM_PHASES")) {
        BB_CUDA(cudaMalloc(&d_phase, 4 * sizeof(long long)));
        BB_CUDA(cudaMemsetAsync(d_phase, 0, 4 * sizeof(long long), st));
    }
    {
        BBProfScope prof(h, st);
        bb_time_marg_kernel<NDET, APPROX, CAL><<<(unsigned)grid, BB_TM_THREADS, smem, st>>>(
            h->d_coef, n, bb_tiles(h), h->net.n_freq, h->net.df, nfft, log2n, h->d_twiddle, h->marg,
            h->net.start_time, h->net.duration, h->d_calrec, h->cal, out, d_phase);
    }
    h->launches++;
    BB_CUDA(cudaGetLastError());
    if (d_phase) {
        long long pc[4];
        BB_CUDA(cudaStreamSynchronize(st));
        BB_CUDA(cudaMemcpy(pc, d_phase, sizeof(pc), cudaMemcpyDeviceToHost));
        BB_CUDA(cudaFree(d_phase));
        fprintf(stderr, "bb_time_marg phases (cycles/sample): load+zero %.0f  fill %.0f  fft %.0f  logsumexp %.0f  (n=%ld)\n",
                (double)pc[0] / n, (double)pc[1] / n, (double)pc[2] / n, (double)pc[3] / n, n);
    }
    return 0;
}

static int bb_launch_time_marg(bb_handle* h, long n, double* out, cudaStream_t st) {
    if (h->nfft == 0) return bb_fail("time marginalisation needs n_freq - 1 to be a power of two");
    if (h->shard_lo != 0 || h->shard_hi != h->net.n_freq)
        return bb_fail("time marginalisation cannot be frequency-sharded (SURVEY.md section 8e)");
    {
        // large batches: the two-kernel pipeline (bb_timemarg_split.cuh); BB_TM_SPLIT=0/1 forces the choice
        const char* force = getenv("BB_TM_SPLIT");
        const bool split = (h->nfft >= 512) && (force ? (force[0] == '1') : (n >= 2048));
        if (split) return bb_launch_time_marg_split(h, n, out, st);
    }
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
