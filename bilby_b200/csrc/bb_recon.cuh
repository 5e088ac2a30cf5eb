// Marginalised-parameter reconstruction, batched (included by bb_kernels.cu after bb_timemarg_split.cuh).
//
// Replaces GravitationalWaveTransient.generate_posterior_sample_from_marginalized_likelihood and the three
// generate_{time,distance,phase}_sample_from_marginalized_likelihood methods (bilby/gw/likelihood/base.py:502-773),
// which the reference runs one posterior row at a time (a pool over rows, bilby/gw/conversion.py:2366-2449).
// The random part of Interped.sample() (core/prior/base.py:143-164) is a unit-interval draw per marginalised
// parameter; the caller supplies those draws, everything after them is deterministic and restated here:
//
//   time      h conj(d)/S on the 16384 Hz grid = U = 32768/fs transforms of the series modulated by
//             exp(-2 pi i k r / (U nfft)), r < U (K4a fills the series once, bb_series_fine_kernel runs the two
//             radix-16 passes + the pruned final DFT per r), point likelihood per time, prior, > max/1000 cut,
//             Interped (base.py:578-658).  NOTE: the reference's FFT here carries no 4/T factor (base.py:626).
//   distance  <d|h>, <h|h> at the new time (K1), posterior over the 10^4-point distance grid, Interped
//             (base.py:660-708); the signal is then rescaled by ref_dist / new_distance (base.py:1061-1063)
//   phase     posterior over linspace(0, 2 pi, 101), Interped (base.py:746-773)
//
// Interped (core/prior/interpolated.py:12-60, 161-176): the abscissae are replaced by a linspace of the same
// length, the density is re-interpolated linearly onto it (scipy interp1d), normalised with the trapezoid rule,
// integrated with the cumulative trapezoid rule (last element forced to one) and inverted by linear interpolation.
#pragma once

#define BB_RC_THREADS 256
#define BB_RC_WMAX 4096          // most fine-grid times inside the geocent_time prior (0.25 s at 16384 Hz)
#define BB_RC_FINE_RATE 16384.0

// ---- block-level helpers (BB_RC_THREADS threads, arrays in shared memory)
__device__ __forceinline__ double bb_rc_block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = bb_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < BB_RC_THREADS / 32; ++w) t += red[w];
    return t;
}
__device__ __forceinline__ double bb_rc_block_max(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = red[0];
    for (int w = 1; w < BB_RC_THREADS / 32; ++w) t = fmax(t, red[w]);
    return t;
}
// in-place inclusive prefix sum of a[0 .. m)
__device__ __forceinline__ void bb_rc_block_scan(double* a, int m, double* part /* [BB_RC_THREADS] */) {
    const int chunk = (m + BB_RC_THREADS - 1) / BB_RC_THREADS;
    const int lo = min(m, (int)threadIdx.x * chunk), hi = min(m, lo + chunk);
    __syncthreads();
    double run = 0.0;
    for (int i = lo; i < hi; ++i) { run += a[i]; a[i] = run; }
    part[threadIdx.x] = run;
    __syncthreads();
    if (threadIdx.x == 0) {
        double acc = 0.0;
        for (int t = 0; t < BB_RC_THREADS; ++t) { const double v = part[t]; part[t] = acc; acc += v; }
    }
    __syncthreads();
    const double off = part[threadIdx.x];
    for (int i = lo; i < hi; ++i) a[i] += off;
    __syncthreads();
}

// numpy.linspace(x0, x1, m)[i]
__device__ __forceinline__ double bb_rc_linspace(double x0, double x1, double step, int i, int m) {
    return i == m - 1 ? x1 : __dadd_rn(__dmul_rn((double)i, step), x0);
}

// scipy.interpolate.interp1d(kind="linear") at xq over sorted abscissae X(i), ordinates y[i], i < m
template <class XF>
__device__ __forceinline__ double bb_rc_interp(XF X, const double* y, int m, double xq) {
    int lo = 0, hi = m;                          // searchsorted(x, xq, side="left")
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (X(mid) < xq) lo = mid + 1; else hi = mid;
    }
    int idx = lo < 1 ? 1 : (lo > m - 1 ? m - 1 : lo);
    const double x_lo = X(idx - 1), x_hi = X(idx), y_lo = y[idx - 1], y_hi = y[idx];
    const double slope = (y_hi - y_lo) / (x_hi - x_lo);
    return __dadd_rn(__dmul_rn(slope, xq - x_lo), y_lo);
}

// Interped(X, y).rescale(u); dens / cdf: [m] scratch; returns the sample (same value in every thread)
template <class XF>
__device__ __forceinline__ double bb_rc_interped(XF X, const double* y, int m, double u, double* dens, double* cdf,
                                                 double* part, double* red) {
    const double x0 = X(0), x1 = X(m - 1);
    const double step = (x1 - x0) / (double)(m - 1);
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += BB_RC_THREADS) dens[i] = bb_rc_interp(X, y, m, bb_rc_linspace(x0, x1, step, i, m));
    __syncthreads();
    // trapezoid(dens, grid)
    double acc = 0.0;
    for (int i = threadIdx.x; i + 1 < m; i += BB_RC_THREADS) {
        const double d = bb_rc_linspace(x0, x1, step, i + 1, m) - bb_rc_linspace(x0, x1, step, i, m);
        acc += __dmul_rn(d, dens[i + 1] + dens[i]) / 2.0;
    }
    const double norm = bb_rc_block_sum(acc, red);
    for (int i = threadIdx.x; i < m; i += BB_RC_THREADS) dens[i] = dens[i] / norm;
    __syncthreads();
    // cumulative_trapezoid(dens, grid, initial=0), last element forced to one
    for (int i = threadIdx.x; i < m; i += BB_RC_THREADS) {
        if (i == 0) { cdf[0] = 0.0; continue; }
        const double d = bb_rc_linspace(x0, x1, step, i, m) - bb_rc_linspace(x0, x1, step, i - 1, m);
        cdf[i] = __dmul_rn(d, dens[i] + dens[i - 1]) / 2.0;
    }
    bb_rc_block_scan(cdf, m, part);
    if (threadIdx.x == 0) cdf[m - 1] = 1.0;
    __syncthreads();
    // inverse_cumulative_distribution = interp1d(x=cdf, y=grid)
    int lo = 0, hi = m;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] < u) lo = mid + 1; else hi = mid;
    }
    const int idx = lo < 1 ? 1 : (lo > m - 1 ? m - 1 : lo);
    const double g_lo = bb_rc_linspace(x0, x1, step, idx - 1, m), g_hi = bb_rc_linspace(x0, x1, step, idx, m);
    const double slope = (g_hi - g_lo) / (cdf[idx] - cdf[idx - 1]);
    return __dadd_rn(__dmul_rn(slope, u - cdf[idx - 1]), g_lo);
}

// ------------------------------------------------------------------------------------------------
// fine-grid (16384 Hz) series of h conj(d)/S around the geocent_time prior, one CTA per sample
// ------------------------------------------------------------------------------------------------
struct BBFineWindow {
    long j_start;     // first candidate index of the fine grid (mod n16)
    int count;        // number of candidates (<= BB_RC_WMAX)
};

// candidates: every fine-grid index whose time can lie inside [time_min, time_max) (two spare on each side)
__device__ __forceinline__ BBFineWindow bb_rc_window(const BBMarg& marg, double start_time, double dt0, long n16) {
    BBFineWindow w;
    const double o_lo = floor((marg.time_min - start_time - dt0) * BB_RC_FINE_RATE) - 2.0;
    const double o_hi = ceil((marg.time_max - start_time - dt0) * BB_RC_FINE_RATE) + 2.0;
    long lo = (long)o_lo % n16;
    if (lo < 0) lo += n16;
    w.j_start = lo;
    const double c = o_hi - o_lo + 1.0;
    w.count = (int)fmin(fmax(c, 0.0), (double)BB_RC_WMAX);
    return w;
}

template <int NT>
__global__ void __maxnreg__(128)
bb_series_fine_kernel(long n, int chunk, int n_chunks, int n_slots, const double2* __restrict__ series, int ld,
                      const double* __restrict__ slotrec, int nfft, int log2n, int up, const double2* __restrict__ twiddle,
                      BBMarg marg, double start_time, double duration, double2* __restrict__ fine /* [n_slots][WMAX] */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* X = reinterpret_cast<double2*>(smem_raw);
    constexpr int ps = 3;
    const int n_series = (int)bb_tm_series_elems(nfft, ps);
    double* meta = reinterpret_cast<double*>(X + n_series);          // [BB_SF_SLOTREC]
    double2* wl = reinterpret_cast<double2*>(meta + BB_SF_SLOTREC);
    const int tid = threadIdx.x;
    const int L = nfft >> 8;
    const long n16 = (long)up * nfft;
    const int q16 = nfft >> 4;
    for (int e = tid; e < L; e += NT) {
        const double2 w = twiddle[(e & (L / 2 - 1)) << 8];
        wl[e] = (e < L / 2) ? w : make_double2(-w.x, -w.y);
    }
    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const long p = bb_sf_pos(slot, chunk, n_chunks);
        if (p >= n) continue;
        __syncthreads();
        if (tid < BB_SF_SLOTREC) meta[tid] = slotrec[(size_t)slot * BB_SF_SLOTREC + tid];
        __syncthreads();
        if (meta[0] != 0.0) continue;
        const int k0 = (int)meta[1], k1 = min((int)meta[2], nfft);
        const double2* src = series + (size_t)slot * ld;
        const BBFineWindow win = bb_rc_window(marg, start_time, meta[7], n16);
        const double2 nyq = ((int)meta[2] > nfft) ? src[nfft] : make_double2(0.0, 0.0);     // bin k = nfft
        const double scale = duration / 4.0;        // the tiles carry 4/T; the reference's transform here does not
        // coarse outputs needed: jc = (j_start + q) / up, q < count
        const int jc0 = (int)(win.j_start / up);
        const int nc = (int)((win.j_start % up + win.count + up - 1) / up);
        for (int r = 0; r < up; ++r) {
            // first pass from global memory with the modulation exp(-2 pi i k r / n16), k = t + m q16
            double sn, cs;
            bb_sincospi(-2.0 * (double)r / (double)(16 * up), &sn, &cs);
            const double2 stepw = make_double2(cs, sn);
            for (int t = tid; t < q16; t += NT) {
                double2 v[16];
                bb_sincospi(-2.0 * (double)((long)t * r) / (double)n16, &sn, &cs);
                double2 w = make_double2(cs, sn);
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int k = t + m * q16;
                    const double2 x = (k >= k0 && k < k1) ? src[k] : make_double2(0.0, 0.0);
                    v[m] = bb_cmul(x, w);
                    w = bb_cmul(w, stepw);
                }
                double2 wt = twiddle[t];
#pragma unroll
                for (int st = 0; st < 4; ++st) {
                    const int half = 16 >> (st + 1);
                    double2 wk[8];
#pragma unroll
                    for (int m = 0; m < half; ++m) wk[m] = bb_mul_omega16(wt, m << st);
#pragma unroll
                    for (int g = 0; g < 16; g += 2 * half) {
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
                for (int m = 0; m < 16; ++m) X[bb_tm_pos(t + m * q16, ps)] = v[m];
            }
            __syncthreads();
            bb_tm_pass<4, NT>(X, nfft, 4, ps, twiddle);
            for (int i = tid; i < nc; i += NT) {
                const int jc = (jc0 + i) & (nfft - 1);
                const long j = (long)jc * up + r;
                long q = j - win.j_start;
                if (q < 0) q += n16;
                if (q >= win.count) continue;
                double2 v = bb_tm_pruned_value(X, jc, log2n, ps, wl);
                // Nyquist bin: x[nfft] exp(-2 pi i nfft j / n16) = x[nfft] exp(-i pi j / up)
                bb_sincospi(-(double)(j % (2 * up)) / (double)up, &sn, &cs);
                v.x += nyq.x * cs - nyq.y * sn;
                v.y += nyq.x * sn + nyq.y * cs;
                fine[(size_t)slot * BB_RC_WMAX + q] = make_double2(v.x * scale, v.y * scale);
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// time posterior on the fine grid -> new geocent_time   (base.py:596-658)
// ------------------------------------------------------------------------------------------------
struct BBRcTimeSmem {
    double t[BB_RC_WMAX], y[BB_RC_WMAX], xs[BB_RC_WMAX], ys[BB_RC_WMAX], dens[BB_RC_WMAX], cdf[BB_RC_WMAX];
    double part[BB_RC_THREADS];
    double red[32];
    double meta[BB_SF_SLOTREC];
    int count;
};

struct BBSmemX {
    const double* x;
    __device__ __forceinline__ double operator()(int i) const { return x[i]; }
};

__global__ void __launch_bounds__(BB_RC_THREADS)
bb_recon_time_kernel(long n, int chunk, int n_chunks, int n_slots, const double2* __restrict__ fine,
                     const double* __restrict__ slotrec, int nfft, int up, BBMarg marg, double start_time,
                     double duration, const double* __restrict__ uniforms, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BBRcTimeSmem& sm = *reinterpret_cast<BBRcTimeSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const long n16 = (long)up * nfft;
    BBMarg point = marg;
    point.flags &= ~BB_MARG_TIME;
    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const long p = bb_sf_pos(slot, chunk, n_chunks);
        if (p >= n) continue;
        __syncthreads();
        if (tid < BB_SF_SLOTREC) sm.meta[tid] = slotrec[(size_t)slot * BB_SF_SLOTREC + tid];
        __syncthreads();
        const long s = (long)sm.meta[6];
        if (sm.meta[0] != 0.0) {
            if (tid == 0) out[s * 3] = nan("");
            continue;
        }
        const double dt0 = sm.meta[7], hh = sm.meta[5], dist = sm.meta[3];
        const BBFineWindow win = bb_rc_window(marg, start_time, dt0, n16);
        // times = create_time_series(16384, T, starting_time = t_c - t_start) % T + t_start (series.py:91-112)
        const double stop = (duration + dt0) - 1.0 / BB_RC_FINE_RATE;
        const double step = (stop - dt0) / (double)(n16 - 1);
        double mx = -INFINITY;
        for (int q = tid; q < win.count; q += BB_RC_THREADS) {
            long j = win.j_start + q;
            if (j >= n16) j -= n16;
            double tj = j == n16 - 1 ? stop : __dadd_rn(__dmul_rn((double)j, step), dt0);
            tj = fmod(tj, duration);
            tj = tj + start_time;
            const bool inside = (tj >= marg.time_min) && (tj < marg.time_max);
            const double2 v = fine[(size_t)slot * BB_RC_WMAX + q];
            double l = inside ? bb_point_lnl(point, v.x, v.y, hh, dist) : -INFINITY;
            if (isnan(l)) l = -INFINITY;
            sm.t[q] = tj;
            sm.y[q] = inside ? l : nan("");      // nan marks "outside the prior" (dropped before any use)
            mx = fmax(mx, l);
        }
        mx = bb_rc_block_max(mx, sm.red);
        // time_post = exp(time_log_like - max) * prior.prob(times); Uniform prior (the one K4 supports)
        const double pr = 1.0 / (marg.time_max - marg.time_min);
        double pmax = 0.0;
        for (int q = tid; q < win.count; q += BB_RC_THREADS) {
            const double l = sm.y[q];
            const double post = isnan(l) ? -1.0 : exp(l - mx) * pr;
            sm.y[q] = post;
            pmax = fmax(pmax, post);
        }
        pmax = bb_rc_block_max(pmax, sm.red);
        // keep = time_post > max / 1000 (outside-prior points carry -1 and are never kept)
        const double thr = pmax / 1000.0;
        double cnt = 0.0;
        for (int q = tid; q < win.count; q += BB_RC_THREADS) cnt += (sm.y[q] > thr) ? 1.0 : 0.0;
        cnt = bb_rc_block_sum(cnt, sm.red);
        const bool dilate = cnt < 3.0;
        // (fewer than three kept: keep[1:-1] |= keep[2:] | keep[:-2], base.py:652-653, applied in time order)
        for (int q = tid; q < win.count; q += BB_RC_THREADS) {
            bool k = sm.y[q] > thr;
            if (dilate && !k && sm.y[q] >= 0.0) {
                const bool left = q > 0 && sm.y[q - 1] > thr, right = q + 1 < win.count && sm.y[q + 1] > thr;
                const bool interior = q > 0 && q + 1 < win.count && sm.y[q - 1] >= 0.0 && sm.y[q + 1] >= 0.0;
                k = interior && (left || right);
            }
            sm.cdf[q] = k ? 1.0 : 0.0;
        }
        bb_rc_block_scan(sm.cdf, win.count, sm.part);
        const int m = win.count > 0 ? (int)sm.cdf[win.count - 1] : 0;
        for (int q = tid; q < win.count; q += BB_RC_THREADS) {
            const double here = sm.cdf[q], before = q > 0 ? sm.cdf[q - 1] : 0.0;
            if (here > before) {
                const int pos = (int)here - 1;
                sm.xs[pos] = sm.t[q];
                sm.ys[pos] = sm.y[q];
            }
        }
        __syncthreads();
        double res = nan("");
        if (m >= 2) {
            BBSmemX X{sm.xs};
            res = bb_rc_interped(X, sm.ys, m, uniforms[s * 3], sm.dens, sm.cdf, sm.part, sm.red);
        }
        if (tid == 0) out[s * 3] = res;
    }
}

// ------------------------------------------------------------------------------------------------
// distance and phase posteriors -> new luminosity_distance, new phase   (base.py:660-708, 746-773)
// ------------------------------------------------------------------------------------------------
struct BBGlobalX {
    const double* x;
    __device__ __forceinline__ double operator()(int i) const { return x[i]; }
};
struct BBPhaseX {
    __device__ __forceinline__ double operator()(int i) const {
        const double stop = 6.283185307179586;          // numpy: 2 * np.pi
        return i == 100 ? stop : __dadd_rn(__dmul_rn((double)i, stop / 100.0), 0.0);
    }
};

__global__ void __launch_bounds__(BB_RC_THREADS)
bb_recon_distance_phase_kernel(const double* __restrict__ params, const double* __restrict__ snr, long n, int n_det,
                               BBMarg marg, const double* __restrict__ dist_grid, const double* __restrict__ dist_prior,
                               int nd, int nbuf, const double* __restrict__ uniforms,
                               double* __restrict__ yscratch /* [grid][nbuf] */, double* __restrict__ out) {
    extern __shared__ __align__(16) double rc_smem[];
    double* y = yscratch + (size_t)blockIdx.x * nbuf;    // posterior ordinates (global, L2 resident)
    double* dens = rc_smem;              // [nbuf]
    double* cdf = dens + nbuf;           // [nbuf]
    double* part = cdf + nbuf;           // [BB_RC_THREADS]
    double* red = part + BB_RC_THREADS;  // [32]
    const int tid = threadIdx.x;
    for (long s = blockIdx.x; s < n; s += gridDim.x) {
        double dre = 0.0, dim = 0.0, hh = 0.0;
        for (int d = 0; d < n_det; ++d) {
            const double* v = snr + (s * n_det + d) * 3;
            dre += v[0];
            dim += v[1];
            hh += v[2];
        }
        const double dl = params[s * BB_NPARAM + BB_P_DISTANCE];
        double new_dist = dl, new_phase = params[s * BB_NPARAM + BB_P_PHASE];
        const bool bad = isnan(hh);      // waveform-domain error
        __syncthreads();
        if ((marg.flags & BB_MARG_DISTANCE) && !bad) {
            double mx = -INFINITY;
            for (int i = tid; i < nd; i += BB_RC_THREADS) {
                const double di = dist_grid[i];
                const double xr = dre * dl / di, xi = dim * dl / di;
                const double hd = hh * (dl * dl) / (di * di);
                const double l = (marg.flags & BB_MARG_PHASE) ? bb_ln_i0(hypot(xr, xi), bb_i0e_a, bb_i0e_b) - hd / 2
                                                               : xr - hd / 2;
                y[i] = l;
                mx = fmax(mx, l);
            }
            mx = bb_rc_block_max(mx, red);
            for (int i = tid; i < nd; i += BB_RC_THREADS) y[i] = exp(y[i] - mx) * dist_prior[i];
            __syncthreads();
            BBGlobalX X{dist_grid};
            new_dist = bb_rc_interped(X, y, nd, uniforms[s * 3 + 1], dens, cdf, part, red);
            // _rescale_signal: the polarisations are multiplied by ref_dist / new_distance (base.py:1061-1063)
            const double sc = marg.ref_dist / new_dist;
            dre *= sc;
            dim *= sc;
            hh *= sc * sc;
        }
        __syncthreads();
        if ((marg.flags & BB_MARG_PHASE) && !bad) {
            BBPhaseX P;
            double mx = -INFINITY;
            for (int i = tid; i < 101; i += BB_RC_THREADS) {
                // phasor = exp(-2j phases); Re(d_inner_h * phasor) - hh / 2
                const double ph = P(i);
                const double c2 = cos(-2.0 * ph), s2 = sin(-2.0 * ph);
                const double l = (dre * c2 - dim * s2) - hh / 2;
                y[i] = l;
                mx = fmax(mx, l);
            }
            mx = bb_rc_block_max(mx, red);
            for (int i = tid; i < 101; i += BB_RC_THREADS) y[i] = exp(y[i] - mx);
            __syncthreads();
            new_phase = bb_rc_interped(P, y, 101, uniforms[s * 3 + 2], dens, cdf, part, red);
        }
        if (tid == 0) {
            out[s * 3 + 1] = bad ? nan("") : new_dist;
            out[s * 3 + 2] = bad ? nan("") : new_phase;
        }
    }
}

// rows with geocent_time replaced by the reconstructed time (and the jitter cleared)
__global__ void bb_recon_set_time_kernel(const double* __restrict__ params, const double* __restrict__ out3, long n,
                                         int use_new_time, double* __restrict__ rows) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < BB_NPARAM; ++k) rows[i * BB_NPARAM + k] = params[i * BB_NPARAM + k];
    if (use_new_time) {
        const double t = out3[i * 3];
        if (!isnan(t)) rows[i * BB_NPARAM + BB_P_GEOCENT_TIME] = t;
        rows[i * BB_NPARAM + BB_P_TIME_JITTER] = 0.0;
    }
}

template <int NDET, int APPROX, bool CAL>
static int bb_recon_time_t(bb_handle* h, long n, const double* uniforms, double* out3, cudaStream_t st) {
    const int nfft = h->nfft;
    int log2n = 0;
    while ((1 << log2n) < nfft) ++log2n;
    if (log2n < 9) return bb_fail("reconstruction: nfft < 512");
    const double upd = 2.0 * BB_RC_FINE_RATE / h->net.sampling_frequency;
    const int up = (int)upd;
    if (up < 1 || (double)up != upd || (up & (up - 1)))
        return bb_fail("reconstruction: 32768 / sampling_frequency must be a power of two");
    if ((h->marg.time_max - h->marg.time_min) * BB_RC_FINE_RATE + 8 > BB_RC_WMAX)
        return bb_fail("reconstruction: geocent_time prior wider than 0.249 s is not supported on the device");
    constexpr int NT = 256;
    const size_t smem_a = sizeof(SFSmem<NDET>) + (CAL ? (size_t)BB_SF_SB * NDET * 4 * h->cal.n_points * sizeof(double) : 0);
    const size_t smem_f = bb_tm_series_elems(nfft, 3) * sizeof(double2) + BB_SF_SLOTREC * sizeof(double)
                          + (size_t)(nfft >> 8) * sizeof(double2);
    const size_t smem_t = sizeof(BBRcTimeSmem);
    if (smem_f > 227 * 1024) return bb_fail("reconstruction: series does not fit shared memory (nfft > 8192)");
    BB_CUDA(cudaFuncSetAttribute(bb_series_fill_kernel<NDET, APPROX, CAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    BB_CUDA(cudaFuncSetAttribute(bb_series_fine_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
    BB_CUDA(cudaFuncSetAttribute(bb_recon_time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    const long total_blocks = (n + BB_SF_SB - 1) / BB_SF_SB;
    const int n_chunks = (int)((total_blocks + BB_SF_BLOCKS_PER_CHUNK - 1) / BB_SF_BLOCKS_PER_CHUNK);
    const size_t slots_cap = (size_t)BB_SF_BLOCKS_PER_CHUNK * BB_SF_SB;
    const int ld = nfft + 8;
    const size_t need = slots_cap * (size_t)ld;
    if (need > h->series_cap) {
        cudaFree(h->d_series);
        h->d_series = nullptr;
        h->series_cap = 0;
        BB_CUDA(cudaMalloc(&h->d_series, need * sizeof(double2)));
        h->series_cap = need;
    }
    if (!h->d_slotrec_rc) BB_CUDA(cudaMalloc(&h->d_slotrec_rc, slots_cap * BB_SF_SLOTREC * sizeof(double)));
    if (!h->d_fine) BB_CUDA(cudaMalloc(&h->d_fine, slots_cap * BB_RC_WMAX * sizeof(double2)));
    const unsigned* perm = h->perm_valid ? h->d_perm : nullptr;
    for (int c = 0; c < n_chunks; ++c) {
        const long nb = (total_blocks - 1 - c) / n_chunks + 1;
        const int n_slots = (int)(nb * BB_SF_SB);
        const unsigned grid_a = (unsigned)(nb < h->sm_count ? nb : h->sm_count);
        bb_series_fill_kernel<NDET, APPROX, CAL><<<grid_a, BB_SF_THREADS, smem_a, st>>>(
            h->d_coef, perm, n, c, n_chunks, n_slots, bb_tiles(h), h->net.df, nfft + 1, ld, h->d_calrec, h->cal,
            h->d_series, h->d_slotrec_rc);
        BB_CUDA(cudaGetLastError());
        const unsigned grid_b = (unsigned)(n_slots < h->sm_count ? n_slots : h->sm_count);
        bb_series_fine_kernel<NT><<<grid_b, NT, smem_f, st>>>(n, c, n_chunks, n_slots, h->d_series, ld, h->d_slotrec_rc,
                                                             nfft, log2n, up, h->d_twiddle, h->marg, h->net.start_time,
                                                             h->net.duration, h->d_fine);
        BB_CUDA(cudaGetLastError());
        bb_recon_time_kernel<<<grid_b, BB_RC_THREADS, smem_t, st>>>(n, c, n_chunks, n_slots, h->d_fine, h->d_slotrec_rc,
                                                                     nfft, up, h->marg, h->net.start_time,
                                                                     h->net.duration, uniforms, out3);
        BB_CUDA(cudaGetLastError());
        h->launches += 3;
    }
    return 0;
}

static int bb_recon_time(bb_handle* h, long n, const double* uniforms, double* out3, cudaStream_t st) {
    if (h->nfft == 0) return bb_fail("reconstruction: n_freq - 1 must be a power of two");
    const bool pd = h->wf.approximant == BB_IMRPHENOMD;
    const bool cal = h->cal_params != nullptr;
#define BB_RC_CASE(N)                                                                                          \
    case N:                                                                                                    \
        if (cal) return pd ? bb_recon_time_t<N, BB_IMRPHENOMD, true>(h, n, uniforms, out3, st)                 \
                           : bb_recon_time_t<N, BB_TAYLORF2, true>(h, n, uniforms, out3, st);                  \
        return pd ? bb_recon_time_t<N, BB_IMRPHENOMD, false>(h, n, uniforms, out3, st)                         \
                  : bb_recon_time_t<N, BB_TAYLORF2, false>(h, n, uniforms, out3, st);
    switch (h->net.n_det) {
        BB_RC_CASE(1)
        BB_RC_CASE(2)
        BB_RC_CASE(3)
        BB_RC_CASE(4)
    }
#undef BB_RC_CASE
    return bb_fail("bad n_det");
}

static int bb_reconstruct(bb_handle* h, const double* params_dev, const double* cal_params_dev, long n,
                          const double* uniforms_dev, double* out_dev, cudaStream_t st) {
    const int flags = h->marg.flags;
    if ((flags & BB_MARG_DISTANCE) && !h->d_rc_dist) return bb_fail("reconstruction: bb_set_reconstruction_grid was not called");
    if (h->kind != 0) return bb_fail("reconstruction: full-grid likelihood only");
    if (h->cm_n_curves > 0 && ((flags & BB_MARG_TIME) || cal_params_dev))
        return bb_fail("reconstruction with calibration marginalisation: no time marginalisation, no per-sample calibration parameters");
    if (h->shard_lo != 0 || h->shard_hi != h->net.n_freq) return bb_fail("reconstruction cannot be frequency-sharded");
    if (bb_ensure_scratch(h, (size_t)n)) return 1;
    if ((size_t)n > h->rc_rows_cap) {
        cudaFree(h->d_rc_rows);
        h->d_rc_rows = nullptr;
        const size_t cap = n < 4096 ? 4096 : (size_t)n;
        BB_CUDA(cudaMalloc(&h->d_rc_rows, cap * BB_NPARAM * sizeof(double)));
        h->rc_rows_cap = cap;
    }
    // detector-based frames: convert once, then work in (ra, dec, geocent_time)
    const BBFrame saved = h->frame;
    if (saved.sky_frame || saved.detector_time) {
        bb_sky_frame_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(params_dev, n, saved, h->d_rc_rows, nullptr);
        h->launches++;
        BB_CUDA(cudaGetLastError());
        // the converted rows are overwritten below by bb_recon_set_time_kernel (same values, new time), so keep a copy
        if ((size_t)n > h->rc_rows2_cap) {
            cudaFree(h->d_rc_rows2);
            h->d_rc_rows2 = nullptr;
            const size_t cap = n < 4096 ? 4096 : (size_t)n;
            BB_CUDA(cudaMalloc(&h->d_rc_rows2, cap * BB_NPARAM * sizeof(double)));
            h->rc_rows2_cap = cap;
        }
        BB_CUDA(cudaMemcpyAsync(h->d_rc_rows2, h->d_rc_rows, (size_t)n * BB_NPARAM * sizeof(double), cudaMemcpyDeviceToDevice, st));
        params_dev = h->d_rc_rows2;
        h->frame.sky_frame = 0;
        h->frame.detector_time = 0;
    }
    h->cal_params = cal_params_dev;
    int rc = 0;
    if (flags & BB_MARG_TIME) {
        rc = bb_launch_prologue(h, params_dev, n, st);
        if (!rc) rc = bb_recon_time(h, n, uniforms_dev, out_dev, st);
    }
    if (!rc) {
        bb_recon_set_time_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(params_dev, out_dev, n,
                                                                             (flags & BB_MARG_TIME) ? 1 : 0, h->d_rc_rows);
        h->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = bb_fail("reconstruction: launch failed");
    }
    if (!rc) rc = bb_launch_prologue(h, h->d_rc_rows, n, st);
    if (!rc) {
        if (h->cm_n_curves > 0) {
            // column 0 of uniforms / out: the response-curve draw / recalib_index (base.py:526-529, 544-578); the
            // distance and phase steps below see the chosen curve's inner products (base.py:289-290)
            BBCalSelect sel{uniforms_dev, h->d_snr, out_dev};
            rc = bb_launch_calmarg(h, n, nullptr, st, sel);
        } else {
            rc = bb_launch_inner(h, n, h->d_snr, st);
        }
    }
    if (!rc) {
        const int nd = (flags & BB_MARG_DISTANCE) ? h->rc_nd : 101;
        const int nbuf = nd < 101 ? 101 : nd;
        const size_t smem = ((size_t)2 * nbuf + BB_RC_THREADS + 32) * sizeof(double);
        if (smem > 227 * 1024) rc = bb_fail("reconstruction: distance grid too long for shared memory");
        long grid = 2L * h->sm_count;
        if (grid > n) grid = n;
        if (!rc && (size_t)grid * nbuf > h->rc_y_cap) {
            cudaFree(h->d_rc_y);
            h->d_rc_y = nullptr;
            if (cudaMalloc(&h->d_rc_y, (size_t)2 * h->sm_count * nbuf * sizeof(double)) != cudaSuccess)
                rc = bb_fail("reconstruction: out of device memory");
            else h->rc_y_cap = (size_t)2 * h->sm_count * nbuf;
        }
        if (!rc) {
            cudaFuncSetAttribute(bb_recon_distance_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            bb_recon_distance_phase_kernel<<<(unsigned)grid, BB_RC_THREADS, smem, st>>>(
                h->d_rc_rows, h->d_snr, n, h->net.n_det, h->marg, h->d_rc_dist, h->d_rc_prior, h->rc_nd, nbuf,
                uniforms_dev, h->d_rc_y, out_dev);
            h->launches++;
            if (cudaGetLastError() != cudaSuccess) rc = bb_fail("reconstruction: launch failed");
        }
    }
    h->cal_params = nullptr;
    h->frame = saved;
    return rc;
}
