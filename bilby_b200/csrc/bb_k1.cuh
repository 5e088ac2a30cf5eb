// K1: fused waveform -> detector projection -> <h|d>, <h|h>   (included by bb_kernels.cu)
//
// Replaces, per sample, bilby/gw/source.py:552-690 (waveform on the grid), bilby/gw/detector/
// interferometer.py:303-368 (projection, phase ramp) and :607-640 -> bilby/gw/utils.py:118-138 (inner
// products) without ever materialising h(f).
//
// Mapping: one CTA of 16 warps per SM; each warp owns one sample of the block's 16; lane l owns the bins
// k = 32 r + l of row r.  Data tiles (u = f^-1/6, ln f, f^3/4, d/S, 1/S per detector; 96 B per bin for
// three detectors) are streamed through shared memory in chunks of 512 bins with 1-D TMA bulk copies
// (cp.async.bulk + mbarrier, SASS UBLKCP), double buffered, and reused by the 16 samples of the block.
// Samples arrive sorted by active-bin count, so the 16 warps of a block run in step and whole chunks
// above the block's cut-off frequency are never loaded.  Per sample and detector the phase ramp
// exp(2 pi i f dt_det) is advanced row to row by one complex multiplication (anchored with sincospi once
// per sample); rows that lie inside one (amplitude, phase) region of IMRPhenomD run a loop specialised
// for that region with its coefficients in registers, rows that straddle a region boundary take the
// generic per-lane path.  Partial sums stay in registers; one warp-shuffle reduction per sample.
#pragma once

#ifndef BB_K1_THREADS
#define BB_K1_THREADS 512
#endif
#ifndef BB_K1_UNROLL
#define BB_K1_UNROLL 1
#endif
#define BB_PRAGMA_(x) _Pragma(#x)
#define BB_UNROLL(n) BB_PRAGMA_(unroll n)
#define BB_K1_WARPS (BB_K1_THREADS / 32)
#define BB_K1_SB BB_K1_WARPS               // samples per block (one per warp)
#define BB_K1_CHUNK 512                    // bins per tile: kernels with a calibration record or four detectors
#define BB_K1_CHUNK_LARGE 768              // ... and the others (209 KB of tiles: 51.6 instead of 50.7 M eval/s on configs[1];
                                           // 256-bin tiles: 46.3 M - the per-tile barrier and region set-up are not free)
#define BB_K1_PAD 1536                     // the frequency axis is padded to a common multiple of both
template <int NDET, bool CAL>
struct K1Chunk { static constexpr int value = (!CAL && NDET <= 3) ? BB_K1_CHUNK_LARGE : BB_K1_CHUNK; };

template <int NDET, int CHUNK>
struct K1Tile {
    double u[CHUNK];
    double lf[CHUNK];
    double q34[CHUNK];
    double rf[CHUNK];
    double u7[CHUNK];
    double ff[CHUNK];
    double t3[CHUNK];
    double x3[CHUNK];
    double2 ds[NDET][CHUNK];
    double is[NDET][CHUNK];
};

template <int NDET, int CHUNK>
struct K1Smem {
    K1Tile<NDET, CHUNK> tile[2];
    double coef[BB_K1_SB][BC_NCOEF];
    unsigned long long bar[2];
    int krange[2];
};

// ---- mbarrier / bulk-copy primitives (PTX; sm_90+)
__device__ __forceinline__ unsigned bb_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bb_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bb_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bb_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(bb_smem_u32(dst)), "l"(src), "r"(bytes), "r"(bb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bb_mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "BB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra BB_DONE;\n"
        "bra BB_WAIT;\n"
        "BB_DONE:\n"
        "}" ::"r"(bb_smem_u32(bar)), "r"(parity) : "memory");
}

template <int NDET, int CHUNK>
__device__ __forceinline__ void bb_k1_issue_tile(K1Tile<NDET, CHUNK>& t, unsigned long long* bar, const BBTiles& g, int c0) {
    bb_mbar_expect_tx(bar, (unsigned)sizeof(K1Tile<NDET, CHUNK>));
    bb_bulk_g2s(t.u, g.u + c0, CHUNK * 8, bar);
    bb_bulk_g2s(t.lf, g.lf + c0, CHUNK * 8, bar);
    bb_bulk_g2s(t.q34, g.q34 + c0, CHUNK * 8, bar);
    bb_bulk_g2s(t.rf, g.rf + c0, CHUNK * 8, bar);
    bb_bulk_g2s(t.u7, g.u7 + c0, CHUNK * 8, bar);
    bb_bulk_g2s(t.ff, g.ff + c0, CHUNK * 8, bar);
    bb_bulk_g2s(t.t3, g.t3 + c0, CHUNK * 8, bar);
    bb_bulk_g2s(t.x3, g.x3 + c0, CHUNK * 8, bar);
#pragma unroll
    for (int d = 0; d < NDET; ++d) {
        bb_bulk_g2s(t.ds[d], g.ds + (size_t)d * g.n_pad + c0, CHUNK * 16, bar);
        bb_bulk_g2s(t.is[d], g.is + (size_t)d * g.n_pad + c0, CHUNK * 8, bar);
    }
}

// ---- region-specialised pieces of IMRPhenomD (coefficients already in registers)
struct K1AmpIns { double k[10]; };
struct K1AmpInt { double k[5], f1, invw; };
struct K1AmpMr { double frd, wl2, g, lam; };
struct K1PhIns { double q[13]; };
struct K1PhInt { double q[4]; };
struct K1PhMr { double q[7]; };

// Every eval works from the tile's per-bin columns: f, t = f^(-1/3) and x = f^(1/3) (inspiral only), rf = 1/f, ln f,
// f^(3/4): no conversion, multiplication or power per bin.  The amplitude prefactor a0 is folded into the region's
// coefficients when they are loaded; the common factor f^(-7/6) comes from the tile (K1Tile::u7).
template <int AR>
struct K1Amp;
template <>
struct K1Amp<0> {
    K1AmpIns c;
    __device__ __forceinline__ void load(const double* r) {
        const double a0 = r[BC_A0];
#pragma unroll
        for (int i = 0; i < 10; ++i) c.k[i] = r[BC_AINS + i] * a0;
    }
    __device__ __forceinline__ void begin(double f0, double dfrow) {}
    __device__ __forceinline__ void next() {}
    __device__ __forceinline__ double eval(double f, double x) const {
        double a = c.k[9];
#pragma unroll
        for (int i = 8; i >= 0; --i) a = a * x + c.k[i];
        return a;
    }
};
template <>
struct K1Amp<1> {
    K1AmpInt c;
    __device__ __forceinline__ void load(const double* r) {
        const double a0 = r[BC_A0];
#pragma unroll
        for (int i = 0; i < 5; ++i) c.k[i] = r[BC_AINT + i] * a0;
        c.f1 = r[BC_AINT_F1];
        c.invw = r[BC_AINT_INVW];
    }
    __device__ __forceinline__ void begin(double f0, double dfrow) {}
    __device__ __forceinline__ void next() {}
    __device__ __forceinline__ double eval(double f, double x) const {
        const double xs = (f - c.f1) * c.invw;
        double a = c.k[4];
#pragma unroll
        for (int i = 3; i >= 0; --i) a = a * xs + c.k[i];
        return a;
    }
};
template <>
struct K1Amp<2> {
    K1AmpMr c;
    double a0;
    __device__ __forceinline__ void load(const double* r) {
        c.frd = r[BC_MR_FRD];
        c.wl2 = r[BC_MR_WL2];
        c.g = r[BC_MR_G];
        c.lam = r[BC_MR_LAM];
        a0 = r[BC_A0];
    }
    // a0 g exp(-lam (f - f_RD)) advances from row to row by one multiplication (rows are equally spaced in f)
    double e, ratio;
    __device__ __forceinline__ void begin(double f0, double dfrow) {
        e = a0 * c.g * exp(-c.lam * (f0 - c.frd));
        ratio = exp(-c.lam * dfrow);
    }
    __device__ __forceinline__ void next() { e *= ratio; }
    __device__ __forceinline__ double eval(double f, double x) const {
        const double d = f - c.frd;
        return e * bb_rcp_pos(fma(d, d, c.wl2));
    }
};

// phase in half turns; NEEDS_X: the region's formulas use x = f^(1/3), t = f^(-1/3)
template <int PR>
struct K1Ph;
template <>
struct K1Ph<0> {
    K1PhIns c;
    __device__ __forceinline__ void load(const double* r) {
#pragma unroll
        for (int i = 0; i < 13; ++i) c.q[i] = r[BC_PINS + i];
    }
    __device__ __forceinline__ double eval(double f, double t, double x, double rf, double lf, double q34) const {
        const double* q = c.q;
        double pos = q[6];
        pos = pos * x + q[5]; pos = pos * x + q[4]; pos = pos * x + q[3]; pos = pos * x + q[2]; pos = pos * x + q[1];
        double neg = q[10] * t * t + q[9];
        neg = neg * t + q[8]; neg = neg * t + q[7];
        return q[0] + pos * x + neg * t + lf * (q[11] + q[12] * x);
    }
};
template <>
struct K1Ph<1> {
    K1PhInt c;
    __device__ __forceinline__ void load(const double* r) {
#pragma unroll
        for (int i = 0; i < 4; ++i) c.q[i] = r[BC_PINT + i];
    }
    __device__ __forceinline__ double eval(double f, double t, double x, double rf, double lf, double q34) const {
        return c.q[0] + c.q[1] * f + c.q[2] * (rf * rf * rf) + c.q[3] * lf;
    }
};
template <>
struct K1Ph<2> {
    K1PhMr c;
    __device__ __forceinline__ void load(const double* r) {
#pragma unroll
        for (int i = 0; i < 7; ++i) c.q[i] = r[BC_PMR + i];
    }
    __device__ __forceinline__ double eval(double f, double t, double x, double rf, double lf, double q34) const {
        return c.q[0] + c.q[1] * f + c.q[2] * rf + c.q[3] * q34 + c.q[4] * bb_atan((f - c.q[5]) * c.q[6]);
    }
};

// per-warp running state of one sample.
// The detector's phase ramp exp(+2 pi i f dt_d) at bin (row r, lane l) is ramp_d[r0, l] R_d^(r - r0) with R_d the
// advance over one row, so   sum_r ramp_d[r] t_r = ramp_d[r_last] * sum_r conj(R_d)^(r_last - r) t_r,   t_r = conj(h22) d/S:
// the sum is a Horner scheme in conj(R_d) over the rows,  acc <- acc conj(R_d) + t_r  (8 FP64 operations per detector
// and row instead of 12 for rotate + accumulate + advance, and no ramp registers); the lane's ramp at the LAST row
// multiplies once per sample, where the anchor sincospi used to be.
template <int NDET>
struct K1State {
    double acc[NDET][3];     // Horner accumulators of conj(h/K) d/S (re, im); sum A^2 / S
    double step[NDET][2];    // R_d = exp(+i pi * 2 dt_d * 32 df)
    const double* cal;       // CAL: this sample's calibration record [NDET][n_points][4] (shared memory)
    BBCalGrid grid;
};

template <int NDET, bool CAL, bool MASKED = true, class TILE>
__device__ __forceinline__ void bb_k1_accumulate(K1State<NDET>& st, const TILE& tile, int i, bool act,
                                                 double A, double ph) {
    double sn, cs;
    if (MASKED) {           // rows cut by the band edges: lanes outside [kmin, kmax) contribute exactly zero
        ph = act ? ph : 0.0;
        A = act ? A : 0.0;
    }
    bb_sincospi(ph, &sn, &cs);
    const double zr = A * cs, zi = A * sn;      // A e^{+i Phi} = conj(h22 incl. geocentric shift)
    const double A2 = A * A;
    BBCalW cw;
    if (CAL) cw = bb_cal_weights(st.grid.n_points, st.grid.l0[0], st.grid.inv_delta[0], tile.lf[i]);
#pragma unroll
    for (int d = 0; d < NDET; ++d) {
        double wr = zr, wi = zi;
        double hw = A2;
        if (CAL) {
            // h_det *= C(f)  =>  conj(h) picks up amp1 (cr - i ci), |h|^2 picks up amp1^2
            double amp1, cr, ci;
            if (d > 0 && !st.grid.shared)
                cw = bb_cal_weights(st.grid.n_points, st.grid.l0[d], st.grid.inv_delta[d], tile.lf[i]);
            bb_cal_apply_v(st.cal + d * 4 * st.grid.n_points, st.grid.n_points, cw, &amp1, &cr, &ci);
            const double tr = amp1 * (wr * cr + wi * ci), ti = amp1 * (wi * cr - wr * ci);
            wr = tr;
            wi = ti;
            hw = A2 * amp1 * amp1;
        }
        const double2 dd = tile.ds[d][i];
        // t = w d/S;  acc <- acc conj(R) + t
        const double tr = fma(wr, dd.x, -wi * dd.y), ti = fma(wr, dd.y, wi * dd.x);
        const double ar = st.acc[d][0], ai = st.acc[d][1];
        st.acc[d][0] = fma(ar, st.step[d][0], fma(ai, st.step[d][1], tr));
        st.acc[d][1] = fma(ai, st.step[d][0], fma(-ar, st.step[d][1], ti));
        st.acc[d][2] = fma(hw, tile.is[d][i], st.acc[d][2]);
    }
}

// rows [r0, r1) of one chunk, all inside amplitude region AR and phase region PR.  MASKED: the rows may contain lanes
// outside [kmin, kmax) (only the first and the last row of a sample); interior rows skip the selects.
template <int NDET, int AR, int PR, bool CAL, bool MASKED, class TILE>
__device__ __forceinline__ void bb_k1_rows_pd_m(K1State<NDET>& st, const TILE& tile, K1Amp<AR>& amp,
                                                const K1Ph<PR>& phs, int r0, int r1, int c0, int lane, int kmin,
                                                int kmax, double df) {
    constexpr bool NEEDS_X = (AR == 0) || (PR == 0);
    int k = r0 * BB_ROW + lane;
    int i = k - c0;
    BB_UNROLL(BB_K1_UNROLL)
    for (int r = r0; r < r1; ++r) {
        const bool act = !MASKED || ((k >= kmin) && (k < kmax));
        const double f = tile.ff[i];
        double t = 0.0, x = 0.0;
        if (NEEDS_X) {
            t = tile.t3[i];
            x = tile.x3[i];
        }
        const double A = amp.eval(f, x) * tile.u7[i];
        const double ph = phs.eval(f, t, x, tile.rf[i], tile.lf[i], tile.q34[i]);
        bb_k1_accumulate<NDET, CAL, MASKED>(st, tile, i, act, A, ph);
        amp.next();
        if (MASKED) k += BB_ROW;
        i += BB_ROW;
    }
}

template <int NDET, int AR, int PR, bool CAL, class TILE>
__device__ __forceinline__ void bb_k1_rows_pd(K1State<NDET>& st, const TILE& tile, const double* rec, int r0,
                                              int r1, int c0, int lane, int kmin, int kmax, double df) {
    K1Amp<AR> amp;
    K1Ph<PR> phs;
    amp.load(rec);
    phs.load(rec);
    amp.begin((double)(r0 * BB_ROW + lane) * df, (double)BB_ROW * df);
    // rows fully inside [kmin, kmax): [ceil(kmin / 32), floor(kmax / 32))
    const int ri0 = min(max((kmin + BB_ROW - 1) / BB_ROW, r0), r1), ri1 = max(min(kmax / BB_ROW, r1), ri0);
    if (r0 < ri0) bb_k1_rows_pd_m<NDET, AR, PR, CAL, true>(st, tile, amp, phs, r0, ri0, c0, lane, kmin, kmax, df);
    if (ri0 < ri1) bb_k1_rows_pd_m<NDET, AR, PR, CAL, false>(st, tile, amp, phs, ri0, ri1, c0, lane, kmin, kmax, df);
    if (ri1 < r1) bb_k1_rows_pd_m<NDET, AR, PR, CAL, true>(st, tile, amp, phs, ri1, r1, c0, lane, kmin, kmax, df);
}

// TaylorF2 (+ tides): one region, the 15 phase coefficients and the amplitude prefactor in registers, f / t / x / f^(-7/6)
// from the tile (the generic path re-reads every coefficient from shared memory and forms the powers per bin)
template <int NDET, bool CAL, bool MASKED, class TILE>
__device__ __forceinline__ void bb_k1_rows_tf2_m(K1State<NDET>& st, const TILE& tile, const double (&q)[BT_NP],
                                                 double a0, int r0, int r1, int c0, int lane, int kmin, int kmax) {
    int k = r0 * BB_ROW + lane;
    int i = k - c0;
    for (int r = r0; r < r1; ++r) {
        const bool act = !MASKED || ((k >= kmin) && (k < kmax));
        const double f = tile.ff[i], t = tile.t3[i], x = tile.x3[i], lf = tile.lf[i];
        const double A = a0 * tile.u7[i];
        // the arithmetic of bb_taylorf2_phase, operation for operation
        const double x2 = x * x, x5 = x2 * x2 * x;
        double tid = q[8];
        tid = tid * x + q[7]; tid = tid * x + q[6]; tid = tid * x + q[5]; tid = tid * x2 + q[4];
        double neg = q[12] * t * t + q[11];
        neg = neg * t + q[10]; neg = neg * t + q[9];
        const double ph = q[0] + x * (q[1] + x * q[2]) + q[3] * f + tid * x5 + neg * t + lf * (q[13] + q[14] * x);
        bb_k1_accumulate<NDET, CAL, MASKED>(st, tile, i, act, A, ph);
        if (MASKED) k += BB_ROW;
        i += BB_ROW;
    }
}

template <int NDET, bool CAL, class TILE>
__device__ __forceinline__ void bb_k1_rows_tf2(K1State<NDET>& st, const TILE& tile, const double (&q)[BT_NP],
                                               double a0, int r0, int r1, int c0, int lane, int kmin, int kmax) {
    const int ri0 = min(max((kmin + BB_ROW - 1) / BB_ROW, r0), r1), ri1 = max(min(kmax / BB_ROW, r1), ri0);
    if (r0 < ri0) bb_k1_rows_tf2_m<NDET, CAL, true>(st, tile, q, a0, r0, ri0, c0, lane, kmin, kmax);
    if (ri0 < ri1) bb_k1_rows_tf2_m<NDET, CAL, false>(st, tile, q, a0, ri0, ri1, c0, lane, kmin, kmax);
    if (ri1 < r1) bb_k1_rows_tf2_m<NDET, CAL, true>(st, tile, q, a0, ri1, r1, c0, lane, kmin, kmax);
}

// generic rows: per-lane region selection (rows straddling a region boundary)
template <int NDET, int APPROX, bool CAL, class TILE>
__device__ __forceinline__ void bb_k1_rows_generic(K1State<NDET>& st, const TILE& tile, const double* rec,
                                                   int r0, int r1, int c0, int lane, int kmin, int kmax, double df) {
    for (int r = r0; r < r1; ++r) {
        const int k = r * BB_ROW + lane, i = k - c0;
        const bool act = (k >= kmin) && (k < kmax);
        const double f = (double)k * df;
        double A, ph;
        bb_wave<APPROX>(rec, f, tile.u[i], tile.lf[i], tile.q34[i], &A, &ph);
        bb_k1_accumulate<NDET, CAL>(st, tile, i, act, A, ph);
    }
}

template <int NDET, bool CAL, class TILE>
__device__ __forceinline__ void bb_k1_dispatch_pd(K1State<NDET>& st, const TILE& tile, const double* rec,
                                                  int r0, int r1, int c0, int lane, int kmin, int kmax, double df,
                                                  int ar, int pr) {
    const int combo = ar * 3 + pr;
    switch (combo) {
        case 0: bb_k1_rows_pd<NDET, 0, 0, CAL>(st, tile, rec, r0, r1, c0, lane, kmin, kmax, df); break;
        case 3: bb_k1_rows_pd<NDET, 1, 0, CAL>(st, tile, rec, r0, r1, c0, lane, kmin, kmax, df); break;
        case 4: bb_k1_rows_pd<NDET, 1, 1, CAL>(st, tile, rec, r0, r1, c0, lane, kmin, kmax, df); break;
        case 5: bb_k1_rows_pd<NDET, 1, 2, CAL>(st, tile, rec, r0, r1, c0, lane, kmin, kmax, df); break;
        case 8: bb_k1_rows_pd<NDET, 2, 2, CAL>(st, tile, rec, r0, r1, c0, lane, kmin, kmax, df); break;
        default: bb_k1_rows_generic<NDET, BB_IMRPHENOMD, CAL>(st, tile, rec, r0, r1, c0, lane, kmin, kmax, df); break;
    }
}

// Frequency-sharded runs (bb_exchange.cuh): instead of one local result array the partial inner products are
// stored straight into this rank's slot of EVERY rank's exchange buffer (peer memory over NVLink; dst[r] already
// points at the slot), sample by sample as the warps finish, so the transfer rides under the arithmetic.
#define BB_MAX_RANKS 8
struct BBPush {
    double* dst[BB_MAX_RANKS];
    int n_dst;               // 0: store to `out` only
};

template <int NDET, int APPROX, bool CAL>
__global__ void __launch_bounds__(BB_K1_THREADS, 1)
bb_inner_product_kernel(const double* __restrict__ coef, const unsigned* __restrict__ perm, long n, BBTiles tiles,
                        double df, int shard_lo, int shard_hi, const double* __restrict__ calrec, BBCalGrid grid,
                        double* __restrict__ out, BBPush push) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int CHUNK = K1Chunk<NDET, CAL>::value;
    K1Smem<NDET, CHUNK>& sm = *reinterpret_cast<K1Smem<NDET, CHUNK>*>(smem_raw);
    double* sm_cal = reinterpret_cast<double*>(smem_raw + sizeof(K1Smem<NDET, CHUNK>));   // [SB][NDET*4*n_points] (CAL)
    const int cal_len = CAL ? NDET * 4 * grid.n_points : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long n_blocks = (n + BB_K1_SB - 1) / BB_K1_SB;

    if (tid == 0) {
        bb_mbar_init(&sm.bar[0], 1);
        bb_mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned phase0 = 0, phase1 = 0;      // parity of the next completion of each stage's barrier
    int issued = 0;                       // tiles issued so far by this CTA (stage = issued & 1), uniform

    for (long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const long p0 = blk * BB_K1_SB;
        const int ns = (int)min((long)BB_K1_SB, n - p0);
        if (tid == 0) { sm.krange[0] = INT_MAX; sm.krange[1] = 0; }
        // coefficient records of this block's samples (sorted order -> sample index through perm)
        for (int i = tid; i < ns * BC_NCOEF; i += BB_K1_THREADS) {
            const int sl = i / BC_NCOEF, j = i - sl * BC_NCOEF;
            const long s = perm ? (long)perm[p0 + sl] : p0 + sl;
            sm.coef[sl][j] = coef[s * BC_NCOEF + j];
        }
        if (CAL) {
            for (int i = tid; i < ns * cal_len; i += BB_K1_THREADS) {
                const int sl = i / cal_len, j = i - sl * cal_len;
                const long s = perm ? (long)perm[p0 + sl] : p0 + sl;
                sm_cal[sl * cal_len + j] = calrec[s * cal_len + j];
            }
        }
        __syncthreads();
        if (tid < ns) {
            const int k0 = max((int)sm.coef[tid][BC_KMIN], shard_lo), k1 = min((int)sm.coef[tid][BC_KMAX], shard_hi);
            if (k1 > k0) {
                atomicMin(&sm.krange[0], k0);
                atomicMax(&sm.krange[1], k1);
            }
        }
        __syncthreads();
        const int kb0 = sm.krange[0], kb1 = sm.krange[1];
        const int cb0 = kb0 / CHUNK, cb1 = (kb1 + CHUNK - 1) / CHUNK;   // chunk index range

        // ---- per-warp sample set-up
        const bool have = warp < ns;
        const double* rec = sm.coef[have ? warp : 0];
        int kmin = 0, kmax = 0;
        if (have) {
            kmin = max((int)rec[BC_KMIN], shard_lo);
            kmax = min((int)rec[BC_KMAX], shard_hi);
            if (kmax < kmin) kmax = kmin;
        }
        const int row_first = kmin / BB_ROW, row_last = (kmax + BB_ROW - 1) / BB_ROW;   // [row_first, row_last)
        K1State<NDET> st;
        st.cal = sm_cal + (have ? warp : 0) * cal_len;
        st.grid = grid;
#pragma unroll
        for (int d = 0; d < NDET; ++d) {
            st.acc[d][0] = st.acc[d][1] = st.acc[d][2] = 0.0;
            st.step[d][0] = rec[BC_DET + BC_DSTRIDE * d + 4];
            st.step[d][1] = rec[BC_DET + BC_DSTRIDE * d + 5];
        }
        int ka1 = 0, ka2 = 0, kp1 = 0, kp2 = 0;
        if (APPROX == BB_IMRPHENOMD) {
            ka1 = (int)rec[BC_KA1]; ka2 = (int)rec[BC_KA2]; kp1 = (int)rec[BC_KP1]; kp2 = (int)rec[BC_KP2];
        }
        double tq[BT_NP], ta0 = 0.0;
        if (APPROX == BB_TAYLORF2) {
#pragma unroll
            for (int j = 0; j < BT_NP; ++j) tq[j] = rec[BT_P + j];
            ta0 = rec[BC_A0];
        }

        // ---- stream the tiles
        if (cb1 > cb0 && tid == 0) {
            bb_k1_issue_tile(sm.tile[issued & 1], &sm.bar[issued & 1], tiles, cb0 * CHUNK);
        }
        for (int cb = cb0; cb < cb1; ++cb) {
            const int stage = issued & 1;
            if (cb + 1 < cb1 && tid == 0) {
                // the other stage was released by the __syncthreads that ended the previous iteration
                bb_k1_issue_tile(sm.tile[stage ^ 1], &sm.bar[stage ^ 1], tiles, (cb + 1) * CHUNK);
            }
            if (stage == 0) { bb_mbar_wait(&sm.bar[0], phase0); phase0 ^= 1; }
            else { bb_mbar_wait(&sm.bar[1], phase1); phase1 ^= 1; }
            ++issued;
            const K1Tile<NDET, CHUNK>& tile = sm.tile[stage];
            const int c0 = cb * CHUNK;
            int r = max(row_first, c0 / BB_ROW);
            const int rend = min(row_last, (c0 + CHUNK) / BB_ROW);
            if (have && r < rend) {
                if (APPROX == BB_IMRPHENOMD) {
                    // split [r, rend) at the rows that contain a region boundary
                    while (r < rend) {
                        const int kf = r * BB_ROW;                 // first bin of row r
                        const int ar = kf < ka1 ? 0 : (kf < ka2 ? 1 : 2);
                        const int pr = kf < kp1 ? 0 : (kf < kp2 ? 1 : 2);
                        // next boundary strictly above kf
                        int nb = INT_MAX;
                        if (ka1 > kf) nb = min(nb, ka1);
                        if (ka2 > kf) nb = min(nb, ka2);
                        if (kp1 > kf) nb = min(nb, kp1);
                        if (kp2 > kf) nb = min(nb, kp2);
                        if (nb < kf + BB_ROW) {
                            // a boundary falls inside this row: generic per-lane path for one row
                            bb_k1_rows_generic<NDET, BB_IMRPHENOMD, CAL>(st, tile, rec, r, r + 1, c0, lane, kmin, kmax, df);
                            r += 1;
                        } else {
                            const int rstop = (nb == INT_MAX) ? rend : min(rend, nb / BB_ROW);
                            bb_k1_dispatch_pd<NDET, CAL>(st, tile, rec, r, rstop, c0, lane, kmin, kmax, df, ar, pr);
                            r = rstop;
                        }
                    }
                } else {
                    bb_k1_rows_tf2<NDET, CAL>(st, tile, tq, ta0, r, rend, c0, lane, kmin, kmax);
                }
            }
            __syncthreads();     // everyone is done with this stage before it is refilled
        }

        // ---- reduce over lanes and write (Re<h|d>, Im<h|d>, <h|h>) per detector
        if (have) {
            const long s = perm ? (long)perm[p0 + warp] : p0 + warp;
#pragma unroll
            for (int d = 0; d < NDET; ++d) {
                // the lane's ramp at the last row this sample visited closes the Horner sum
                double rs, rc;
                bb_sincospi(rec[BC_DET + BC_DSTRIDE * d + 2] * ((double)((row_last - 1) * BB_ROW + lane) * df), &rs, &rc);
                const double ar = st.acc[d][0], ai = st.acc[d][1];
                const double sr = bb_warp_sum(ar * rc - ai * rs);
                const double si = bb_warp_sum(ar * rs + ai * rc);
                const double sh = bb_warp_sum(st.acc[d][2]);
                if (lane == 0) {
                    const double kr = rec[BC_DET + BC_DSTRIDE * d], ki = rec[BC_DET + BC_DSTRIDE * d + 1];
                    const double v0 = kr * sr + ki * si;        // <h|d> = conj(K) * sum
                    const double v1 = kr * si - ki * sr;
                    const double v2 = (rec[BC_STATUS] != 0.0) ? nan("") : rec[BC_DET + BC_DSTRIDE * d + 3] * sh;
                    if (push.n_dst == 0) {
                        double* o = out + (s * NDET + d) * 3;
                        o[0] = v0;
                        o[1] = v1;
                        o[2] = v2;
                    } else {
                        for (int r = 0; r < push.n_dst; ++r) {
                            double* o = push.dst[r] + (s * NDET + d) * 3;
                            o[0] = v0;
                            o[1] = v1;
                            o[2] = v2;
                        }
                    }
                }
            }
        }
        __syncthreads();         // records are overwritten by the next block iteration
    }
}
