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
// The series lives in shared memory (16 B * nfft, 128 KB at 8 s / 2048 Hz); radix-2 decimation-in-time with
// the bit-reversal folded into the store of X[k]; twiddles from a device table.
#pragma once

#define BB_TM_THREADS 256

__device__ __forceinline__ unsigned bb_bitrev(unsigned v, int bits) { return __brev(v) >> (32 - bits); }

// FFT of the (bit-reversed) series in shared memory followed by the weighted logsumexp over the times inside
// the geocent_time prior; shared by the full-grid and the relative-binning time-marginalised kernels.
__device__ __forceinline__ void bb_tm_finish(double2* X, int nfft, int log2n, const double2* __restrict__ twiddle,
                                             const BBMarg& marg, double hh, double dist, double jitter,
                                             double start_time, double duration, double* red, double* out_s) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // in-place radix-2 DIT butterflies on the bit-reversed series
    for (int stage = 0; stage < log2n; ++stage) {
        const int half = 1 << stage;
        const int tstep = nfft >> (stage + 1);
        for (int b = tid; b < (nfft >> 1); b += BB_TM_THREADS) {
            const int j = b & (half - 1);
            const int i0 = ((b >> stage) << (stage + 1)) + j;
            const int i1 = i0 + half;
            const double2 w = twiddle[j * tstep];
            const double2 a = X[i0], bb = X[i1];
            const double tr = bb.x * w.x - bb.y * w.y, ti = bb.x * w.y + bb.y * w.x;
            X[i0] = make_double2(a.x + tr, a.y + ti);
            X[i1] = make_double2(a.x - tr, a.y - ti);
        }
        __syncthreads();
    }

    // weighted logsumexp over the times inside the prior
    const double dtc = duration / (double)nfft;     // = 2 / sampling_frequency
    const double jit = marg.jitter ? jitter : 0.0;
    const double bw = dtc / (marg.time_max - marg.time_min);
    double mx = -INFINITY, sum = 0.0;
    for (int j = tid; j < nfft; j += BB_TM_THREADS) {
        // times = start_time + linspace(0, T, nfft + 1)[1:]  (+ jitter)
        const double tj = (start_time + (double)(j + 1) * dtc) + jit;
        if (tj < marg.time_min || tj > marg.time_max) continue;
        const double2 v = X[j];
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
    for (int w = 1; w < BB_TM_THREADS / 32; ++w) gmx = fmax(gmx, red[w]);
    __syncthreads();
    double part = (mx == -INFINITY) ? 0.0 : sum * exp(mx - gmx);
    part = bb_warp_sum(part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < BB_TM_THREADS / 32; ++w) tot += red[w];
        *out_s = (gmx == -INFINITY) ? -INFINITY : log(tot) + gmx;
    }
}

template <int NDET, int APPROX, bool CAL>
__global__ void __launch_bounds__(BB_TM_THREADS, 1)
bb_time_marg_kernel(const double* __restrict__ coef, long n, BBTiles tiles, int n_freq, double df, int nfft,
                    int log2n, const double2* __restrict__ twiddle, BBMarg marg, double start_time,
                    double duration, const double* __restrict__ calrec, BBCalGrid grid, double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* X = reinterpret_cast<double2*>(smem_raw);
    double* c = reinterpret_cast<double*>(smem_raw + (size_t)nfft * sizeof(double2));
    double* red = c + BC_NCOEF;      // [32]
    double* cal = red + 32;          // CAL: [NDET][4][n_points]
    const int cal_len = CAL ? NDET * 4 * grid.n_points : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (long s = blockIdx.x; s < n; s += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < BC_NCOEF; i += BB_TM_THREADS) c[i] = coef[s * BC_NCOEF + i];
        for (int i = tid; i < nfft; i += BB_TM_THREADS) X[i] = make_double2(0.0, 0.0);
        if (CAL) for (int i = tid; i < cal_len; i += BB_TM_THREADS) cal[i] = calrec[s * cal_len + i];
        __syncthreads();
        if (c[BC_STATUS] != 0.0) {
            if (tid == 0) out[s] = -DBL_MAX;
            continue;
        }
        const int k0 = (int)c[BC_KMIN];
        const int k1 = min((int)c[BC_KMAX], nfft);    // Nyquist bin dropped from the series, kept in <h|h>
        const int k1h = (int)c[BC_KMAX];
        double hh = 0.0;
        for (int k = k0 + tid; k < k1h; k += BB_TM_THREADS) {
            const double f = (double)k * df;
            double A, ph;
            bb_wave<APPROX>(c, f, tiles.u[k], tiles.lf[k], tiles.q34[k], &A, &ph);
            double sn, cs;
            sincospi(ph, &sn, &cs);
            const double zr = A * cs, zi = A * sn;
            const double A2 = A * A;
            double vr = 0.0, vi = 0.0;
#pragma unroll
            for (int d = 0; d < NDET; ++d) {
                double rs, rc;
                sincospi(c[BC_DET + BC_DSTRIDE * d + 2] * f, &rs, &rc);
                double wr = zr * rc - zi * rs, wi = zr * rs + zi * rc;
                double hw = A2;
                if (CAL) {
                    double amp1, cr, ci;
                    bb_cal_factor(cal + d * 4 * grid.n_points, grid.n_points, grid.l0[d], grid.inv_delta[d],
                                  tiles.lf[k], &amp1, &cr, &ci);
                    const double tr = amp1 * (wr * cr + wi * ci), ti = amp1 * (wi * cr - wr * ci);
                    wr = tr;
                    wi = ti;
                    hw = A2 * amp1 * amp1;
                }
                const double2 dd = tiles.ds[(size_t)d * tiles.n_pad + k];
                const double pr = wr * dd.x - wi * dd.y, pi = wr * dd.y + wi * dd.x;   // conj(h/K) d/S
                const double kr = c[BC_DET + BC_DSTRIDE * d], ki = c[BC_DET + BC_DSTRIDE * d + 1];
                vr += kr * pr + ki * pi;       // conj(K) * p
                vi += kr * pi - ki * pr;
                hh += c[BC_DET + BC_DSTRIDE * d + 3] * hw * tiles.is[(size_t)d * tiles.n_pad + k];
            }
            if (k < k1) X[bb_bitrev((unsigned)k, log2n)] = make_double2(vr, -vi);   // h conj(d)/S = conj(conj(h) d/S)
        }
        // block-reduce <h|h>
        hh = bb_warp_sum(hh);
        if (lane == 0) red[warp] = hh;
        __syncthreads();
        hh = 0.0;
        for (int w = 0; w < BB_TM_THREADS / 32; ++w) hh += red[w];
        __syncthreads();

        bb_tm_finish(X, nfft, log2n, twiddle, marg, hh, c[BC_DISTANCE], c[BC_JITTER], start_time, duration, red, out + s);
    }
}

template <int NDET, int APPROX, bool CAL>
static int bb_launch_time_marg_t(bb_handle* h, long n, double* out, cudaStream_t st) {
    const int nfft = h->nfft;
    int log2n = 0;
    while ((1 << log2n) < nfft) ++log2n;
    const size_t smem = (size_t)nfft * sizeof(double2) + (BC_NCOEF + 32 + (CAL ? NDET * 4 * h->cal.n_points : 0)) * sizeof(double);
    if (smem > 227 * 1024) return bb_fail("time marginalisation: series does not fit shared memory (nfft > 8192)");
    BB_CUDA(cudaFuncSetAttribute(bb_time_marg_kernel<NDET, APPROX, CAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    long grid = (long)h->sm_count * per_sm;
    if (grid > n) grid = n;
    bb_time_marg_kernel<NDET, APPROX, CAL><<<(unsigned)grid, BB_TM_THREADS, smem, st>>>(
        h->d_coef, n, bb_tiles(h), h->net.n_freq, h->net.df, nfft, log2n, h->d_twiddle, h->marg,
        h->net.start_time, h->net.duration, h->d_calrec, h->cal, out);
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
