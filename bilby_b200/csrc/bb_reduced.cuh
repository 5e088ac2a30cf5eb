// Reduced-order likelihood kernels (included by bb_kernels.cu).
//
//   K5 bb_relbin_kernel        relative binning  <- bilby/gw/likelihood/relative.py:365-430
//   K6 bb_roq_kernel           ROQ, one coalescence time per sample  <- bilby/gw/likelihood/roq.py:467-602
//   K7 bb_roq_hlinear_kernel + bb_gemm_nt_kernel + bb_roq_time_marg_kernel   ROQ time marginalisation  <- roq.py:535-651
//
//   K5 (no neighbour term)    multi-banding  <- bilby/gw/likelihood/multiband.py:728-765 (bb_set_multiband)
//
// All evaluate the source model on a frequency SEQUENCE (bin edges / ROQ nodes / banded points), the device form of
// bilby/gw/source.py:1068-1140: every node is evaluated with the per-sample coefficient record written by
// K0 (no f_min / f_max masking).  One warp owns one sample; lanes stride over the nodes; the node tables
// (f, f^-1/6, ln f, f^3/4) and the per-node data are read through L1/L2 (they are a few hundred KB and shared
// by every sample; K6 stages its ROQ weights with cp.async.bulk); partial sums stay in registers; the per-sample
// sums are reduced with a halving butterfly (bb_warp_sum8).
#pragma once

#define BB_RED_THREADS 256
#define BB_RED_WARPS (BB_RED_THREADS / 32)
#define BB_ROQ_THREADS 128     // K6: 4 warps per CTA, BB_ROQ_CTAS CTAs per SM (one shared-memory stage of W per warp)
#ifndef BB_ROQ_CTAS
#define BB_ROQ_CTAS 4
#endif
// K5: relative binning (CROSS) is held at 3 CTAs per SM (80 registers); the multi-banded variant is left to ptxas,
// which takes 164 registers = ONE 8-warp CTA per SM and keeps the next row's tables in flight: 9.3e6 eval/s vs 8.7e6
// under a 2-CTA bound (126 registers) and 5.3e6 under a 3-CTA bound (spills); BB_RELBIN_CTAS forces a bound for experiments
#ifdef BB_RELBIN_CTAS
#define BB_RELBIN_BOUNDS __launch_bounds__(BB_RED_THREADS, BB_RELBIN_CTAS)
#else
#define BB_RELBIN_BOUNDS __launch_bounds__(BB_RED_THREADS, CROSS ? 3 : 1)
#endif
#define BB_ROQ_WARPS (BB_ROQ_THREADS / 32)

struct BBNodes {
    const double* f;
    const double* u;      // f^(-1/6)
    const double* lf;     // ln f
    const double* q34;    // f^(3/4)
    const double* t3;     // f^(-1/3) = u u
    const double* x3;     // f^(1/3) = (f t) t
    const double* u7;     // f^(-7/6) = u (t t t)
    // the six columns the kernels read per node (f, t3, x3, u7, lf, q34) a second time in rows of 32 nodes,
    // blk[row][column][32] (the last node repeated beyond n): one running pointer and immediate offsets per row
    // instead of six parameter loads and 64-bit index multiplies
    const double* blk;
    int n;
};
#define BB_NB_F 0
#define BB_NB_T3 32
#define BB_NB_X3 64
#define BB_NB_U7 96
#define BB_NB_LF 128
#define BB_NB_Q34 160
#define BB_NB_ROW 192

struct BBRelbinDev {
    BBNodes edges;            // bin edges (relative.py:233 frequency_bin_edges)
    const double2* ginv;      // [n_det][n_edges]  1 / per_detector_fiducial_waveform_points
    const double2* a0;        // [n_det][n_bins]   summary data (relative.py:339-361)
    const double2* a1;
    const double* b0;         // [n_det][n_bins]   (real: <h0|h0> pieces)
    const double* b1;
    const double* inv_width;  // [n_bins] 1 / bin_widths
    // time marginalisation (relative.py:380-421): P_d[k] = (4/T) h0_d[k] conj(d_d[k]) / S_d[k], bin of grid bin k
    const double2* pgrid;     // [n_det][n_freq]
    const int* bin_of_k;      // [n_freq], -1 outside [bin_inds[0], bin_inds[-1]]
    const double* centre;     // [n_bins] bin_centers
    // edge form of the two sums (bb_relbin_edge_sample): [n_det][ne_pad], zero beyond the last edge
    const double2* lin_c;     // <d|h> = sum_j lin_c[j] conj(h_j)
    const double* quad_e;     // <h|h> = sum_j quad_e[j] |h_j|^2 + Re(cross_g[j] h_j conj(h_{j-1}))
    const double2* cross_g;
    // the same three tables in rows of 32 edges, etab[row][det]{lin_c[32], cross_g[32] (relative binning only),
    // quad_e[32]}: two running pointers and immediate offsets in K5's loop
    const double* etab;
    int ne_pad;
};

struct BBRoqDev {
    BBNodes lin, quad;
    const double2* W;         // [n_det] packed [n_time x n_lin] (bb_gemm.cuh bb_pk, tiles of 128 times): K7's GEMM operand
    const double2* W2;        // [n_det][ceil(n_lin/32)][n_time][32] node-blocked copy for K6, zero beyond n_lin
    const double* wq;         // [n_det][n_quad]
    int n_time;
    long time_start_index;    // time_samples[i] = (time_start_index + i) * time_step   (roq.py:766)
    double time_step;
    int n_marg;               // time marginalisation grid (roq.py:320-331)
    double marg_start, marg_dtc;
};

// load one sample's coefficient record (and calibration record) into this warp's shared-memory slot
template <bool CAL>
__device__ __forceinline__ void bb_red_load(double* rec, double* cal, const double* coef, const double* calrec,
                                            long s, int cal_len, int lane) {
    __syncwarp();
    for (int i = lane; i < BC_NCOEF; i += 32) rec[i] = coef[s * BC_NCOEF + i];
    if (CAL) for (int i = lane; i < cal_len; i += 32) cal[i] = calrec[s * cal_len + i];
    __syncwarp();
}

// The same, asynchronously (cp.async, 16-byte pieces: a record is 84 doubles, a calibration record a multiple of
// 24): the warp fetches the NEXT sample's records into its second slot while it works on the current one.
template <bool CAL>
__device__ __forceinline__ void bb_red_prefetch(double* rec, double* cal, const double* coef, const double* calrec,
                                                long s, int cal_len, int lane) {
    const char* src = reinterpret_cast<const char*>(coef + s * BC_NCOEF);
    for (int i = lane; i < BC_NCOEF / 2; i += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(bb_smem_u32(rec) + 16u * i), "l"(src + 16 * i) : "memory");
    if (CAL) {
        const char* csrc = reinterpret_cast<const char*>(calrec + s * cal_len);
        for (int i = lane; i < cal_len / 2; i += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(bb_smem_u32(cal) + 16u * i), "l"(csrc + 16 * i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bb_red_wait() {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// K5: relative binning
//   r(f_j) = h_det(f_j) / h0_det(f_j) at the bin edges; r0 = mean, r1 = slope over each bin;
//   <d|h> = sum a0 conj(r0) + a1 conj(r1),  <h|h> = sum b0 |r0|^2 + 2 b1 Re(r0 conj(r1))
// Lane l of row r owns edge j = 32 r + l and the bin whose RIGHT edge it is; the left edge's ratio comes
// from the neighbouring lane (or from lane 31 of the previous row).
// STORE: also keep (r0, r1) of every (bin, detector) in shared memory for the time-marginalised variant.
// ------------------------------------------------------------------------------------------------
template <int NDET, int APPROX, bool CAL, bool STORE>
__device__ __forceinline__ void bb_relbin_sample(const double* rec, const double* cal, const BBCalGrid& grid,
                                                 const BBRelbinDev& rb, int lane, double (*acc)[3],
                                                 double2* r01 /* [n_bins][NDET][2] */) {
    const int ne = rb.edges.n, nb = ne - 1;
    double2 carry[NDET];
#pragma unroll
    for (int d = 0; d < NDET; ++d) carry[d] = make_double2(0.0, 0.0);
    for (int base = 0; base < ne; base += 32) {
        const int j = base + lane;
        const bool act = j < ne;
        const int jj = act ? j : ne - 1;
        const double f = rb.edges.f[jj];
        double A, ph;
        bb_wave_cols<APPROX>(rec, f, rb.edges.t3[jj], rb.edges.x3[jj], rb.edges.u7[jj], rb.edges.lf[jj], rb.edges.q34[jj], &A, &ph);
        const int b = j - 1;                               // bin whose right edge is j
        const bool have_bin = act && b >= 0;
        const double iw = have_bin ? rb.inv_width[b] : 0.0;
#pragma unroll
        for (int d = 0; d < NDET; ++d) {
            const double* cd = rec + BC_DET + BC_DSTRIDE * d;
            double sn, cs;
            bb_sincospi(ph + cd[2] * f, &sn, &cs);            // h22 e^{-2 pi i f (dt0 + delay)} = A (cs - i sn)
            double hr = A * (cd[0] * cs + cd[1] * sn), hi = A * (cd[1] * cs - cd[0] * sn);   // K h
            if (CAL) {
                double amp1, cr, ci;
                bb_cal_factor(cal + d * 4 * grid.n_points, grid.n_points, grid.l0[d], grid.inv_delta[d],
                              rb.edges.lf[jj], &amp1, &cr, &ci);
                const double tr = amp1 * (hr * cr - hi * ci), ti = amp1 * (hr * ci + hi * cr);
                hr = tr;
                hi = ti;
            }
            const double2 g = rb.ginv[(size_t)d * ne + jj];
            const double rr = hr * g.x - hi * g.y, ri = hr * g.y + hi * g.x;        // ratio at edge j
            double lr = __shfl_up_sync(0xffffffffu, rr, 1), li = __shfl_up_sync(0xffffffffu, ri, 1);
            if (lane == 0) { lr = carry[d].x; li = carry[d].y; }
            carry[d].x = __shfl_sync(0xffffffffu, rr, 31);
            carry[d].y = __shfl_sync(0xffffffffu, ri, 31);
            if (have_bin) {
                const double r0r = 0.5 * (rr + lr), r0i = 0.5 * (ri + li);
                const double r1r = (rr - lr) * iw, r1i = (ri - li) * iw;
                const double2 a0 = rb.a0[(size_t)d * nb + b], a1 = rb.a1[(size_t)d * nb + b];
                // a conj(r) = (ax rr + ay ri) + i (ay rr - ax ri)
                acc[d][0] += (a0.x * r0r + a0.y * r0i) + (a1.x * r1r + a1.y * r1i);
                acc[d][1] += (a0.y * r0r - a0.x * r0i) + (a1.y * r1r - a1.x * r1i);
                acc[d][2] += rb.b0[(size_t)d * nb + b] * (r0r * r0r + r0i * r0i)
                             + 2.0 * rb.b1[(size_t)d * nb + b] * (r0r * r1r + r0i * r1i);
                if (STORE) {
                    r01[((size_t)b * NDET + d) * 2] = make_double2(r0r, r0i);
                    r01[((size_t)b * NDET + d) * 2 + 1] = make_double2(r1r, r1i);
                }
            }
        }
    }
}


// K5, edge form (the likelihood-only path).  With r_j = h_j / h0_j at the edges, r0 = (r_{j} + r_{j-1}) / 2 and
// r1 = (r_j - r_{j-1}) / width, both sums of relative.py:423-430 regroup into sums over EDGES:
//   <d|h> = sum_j conj(h_j) C_j,          C_j = conj(1/h0_j) [ (a0/2 + a1/w)_{bin j-1} + (a0/2 - a1/w)_{bin j} ]
//   <h|h> = sum_j |h_j|^2 E_j + Re( G_j h_j conj(h_{j-1}) ),
//           E_j = |1/h0_j|^2 [ (b0/4 + b1/w)_{bin j-1} + (b0/4 - b1/w)_{bin j} ],  G_j = (b0/2)_{bin j-1} (1/h0_j) conj(1/h0_{j-1})
// (2 b1 Re(r0 conj r1) = b1 (|r_j|^2 - |r_{j-1}|^2) / w: the cross terms cancel).  The tables are built once in
// bb_set_relative_binning, zero-padded to a multiple of 32 edges, so the loop has no per-bin branches, no ratio and
// one neighbour exchange.
template <int NDET, int APPROX, bool CAL, bool CROSS>
__device__ __forceinline__ void bb_relbin_edge_sample(const double* rec, const double* cal, const BBCalGrid& grid,
                                                      const BBRelbinDev& rb, int lane, double (*acc)[3]) {
    const int ne = rb.edges.n;
    double2 carry[NDET];
#pragma unroll
    for (int d = 0; d < NDET; ++d) carry[d] = make_double2(0.0, 0.0);
    // node tables of the NEXT row are requested before the current row is evaluated (the multi-banded likelihood
    // streams hundreds of rows per sample from L2)
    // (only there: with the two rows of a relative-binning sample the extra registers cost more than they hide)
    constexpr bool PREFETCH = !CROSS;
    const double* nb = rb.edges.blk + lane;
    constexpr int ET = CROSS ? 160 : 96;                   // doubles per (row, detector) of etab
    const double* et = rb.etab + 2 * lane;                 // lin_c (and cross_g 64 doubles on)
    const double* ee = rb.etab + (CROSS ? 128 : 64) + lane;    // quad_e
    double nf = 0.0, nt = 0.0, nx = 0.0, nu7 = 0.0, nlf = 0.0, nq = 0.0;
    if (PREFETCH) {
        nf = nb[BB_NB_F]; nt = nb[BB_NB_T3]; nx = nb[BB_NB_X3]; nu7 = nb[BB_NB_U7];
        nlf = nb[BB_NB_LF]; nq = nb[BB_NB_Q34];
    }
    for (int base = 0; base < ne; base += 32, nb += BB_NB_ROW, et += NDET * ET, ee += NDET * ET) {
        const bool more = base + 32 < ne;
        double f, t, x, u7, lfj, q34;
        if (PREFETCH) {
            f = nf; t = nt; x = nx; u7 = nu7; lfj = nlf; q34 = nq;
            if (more) {
                const double* nn = nb + BB_NB_ROW;
                nf = nn[BB_NB_F];
                nt = nn[BB_NB_T3];
                nx = nn[BB_NB_X3];
                nu7 = nn[BB_NB_U7];
                nlf = nn[BB_NB_LF];
                if (APPROX == BB_IMRPHENOMD) nq = nn[BB_NB_Q34];
            }
        } else {
            f = nb[BB_NB_F]; t = nb[BB_NB_T3]; x = nb[BB_NB_X3]; u7 = nb[BB_NB_U7]; lfj = nb[BB_NB_LF];
            q34 = (APPROX == BB_IMRPHENOMD) ? nb[BB_NB_Q34] : 0.0;
        }
        double A, ph;
        bb_wave_cols<APPROX>(rec, f, t, x, u7, lfj, q34, &A, &ph);
#pragma unroll
        for (int d = 0; d < NDET; ++d) {
            const double* cd = rec + BC_DET + BC_DSTRIDE * d;
            double sn, cs;
            // relative binning (two rows per sample) keeps the variant with the range branch: the branch holds the
            // register allocation at 80 (three CTAs per SM, 6.6e8 eval/s); branch-free it takes 128 (two CTAs, 6.45e8).
            // The multi-banded likelihood streams hundreds of rows and gains 38 % from the branch-free one.
            if (CROSS) bb_sincospi_branchy(ph + cd[2] * f, &sn, &cs);            // h22 e^{-2 pi i f (dt0 + delay)} = A (cs - i sn)
            else bb_sincospi(ph + cd[2] * f, &sn, &cs);
            double hr = A * (cd[0] * cs + cd[1] * sn), hi = A * (cd[1] * cs - cd[0] * sn);   // K h
            if (CAL) {
                double amp1, cr, ci;
                bb_cal_factor(cal + d * 4 * grid.n_points, grid.n_points, grid.l0[d], grid.inv_delta[d],
                              lfj, &amp1, &cr, &ci);
                const double tr = amp1 * (hr * cr - hi * ci), ti = amp1 * (hr * ci + hi * cr);
                hr = tr;
                hi = ti;
            }
            const double2 c = *reinterpret_cast<const double2*>(et + d * ET);
            const double e = ee[d * ET];
            acc[d][0] = fma(c.x, hr, fma(c.y, hi, acc[d][0]));           // C conj(h)
            acc[d][1] = fma(c.y, hr, fma(-c.x, hi, acc[d][1]));
            if (CROSS) {
                const double2 g = *reinterpret_cast<const double2*>(et + d * ET + 64);
                double lr = __shfl_up_sync(0xffffffffu, hr, 1), li = __shfl_up_sync(0xffffffffu, hi, 1);
                if (lane == 0) { lr = carry[d].x; li = carry[d].y; }
                if (more) {
                    carry[d].x = __shfl_sync(0xffffffffu, hr, 31);
                    carry[d].y = __shfl_sync(0xffffffffu, hi, 31);
                }
                const double pr = fma(hr, lr, hi * li), pi = fma(hi, lr, -hr * li);      // h_j conj(h_{j-1})
                acc[d][2] = fma(e, fma(hr, hr, hi * hi), fma(g.x, pr, fma(-g.y, pi, acc[d][2])));
            } else {
                acc[d][2] = fma(e, fma(hr, hr, hi * hi), acc[d][2]);     // multi-banding: no neighbour term
            }
        }
    }
}

template <int NDET, int APPROX, bool CAL, bool CROSS>
__global__ void BB_RELBIN_BOUNDS
bb_relbin_kernel(const double* __restrict__ coef, long n, BBRelbinDev rb, const double* __restrict__ calrec,
                 BBCalGrid grid, double* __restrict__ out) {
    extern __shared__ __align__(16) double red_smem[];
    const int cal_len = CAL ? NDET * 4 * grid.n_points : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // two record slots per warp: the next sample's records arrive while this one is evaluated
    const int slot_len = BC_NCOEF + cal_len;
    double* slots = red_smem + (size_t)warp * 2 * slot_len;
    const long stride = (long)gridDim.x * BB_RED_WARPS;
    long s = (long)blockIdx.x * BB_RED_WARPS + warp;
    if (s < n) bb_red_prefetch<CAL>(slots, slots + BC_NCOEF, coef, calrec, s, cal_len, lane);
    for (int ping = 0; s < n; s += stride, ping ^= 1) {
        double* rec = slots + ping * slot_len;
        double* cal = rec + BC_NCOEF;
        bb_red_wait();
        if (s + stride < n) {
            double* nxt = slots + (ping ^ 1) * slot_len;
            bb_red_prefetch<CAL>(nxt, nxt + BC_NCOEF, coef, calrec, s + stride, cal_len, lane);
        }
        double acc[NDET][3];
#pragma unroll
        for (int d = 0; d < NDET; ++d) acc[d][0] = acc[d][1] = acc[d][2] = 0.0;
        bb_relbin_edge_sample<NDET, APPROX, CAL, CROSS>(rec, cal, grid, rb, lane, acc);
        if (NDET == 3) {
            // nine sums: eight through the halving butterfly (lane 4 q ends with quantity q = 3 d + c), one plain
            const double* flat = &acc[0][0];
            const double t8 = bb_warp_sum8(flat, lane);
            const double t9 = bb_warp_sum(flat[8]);
            const bool bad = rec[BC_STATUS] != 0.0;
            double* o = out + s * NDET * 3;
            const int q = lane >> 2;
            if ((lane & 3) == 0) o[q] = (bad && (q % 3 == 2)) ? nan("") : t8;
            if (lane == 1) o[8] = bad ? nan("") : t9;
        } else {
#pragma unroll
            for (int d = 0; d < NDET; ++d) {
                const double sr = bb_warp_sum(acc[d][0]), si = bb_warp_sum(acc[d][1]), sh = bb_warp_sum(acc[d][2]);
                if (lane == 0) {
                    double* o = out + (s * NDET + d) * 3;
                    o[0] = sr;
                    o[1] = si;
                    o[2] = (rec[BC_STATUS] != 0.0) ? nan("") : sh;
                }
            }
        }
        __syncwarp();       // every lane is done with this slot before it is refilled two samples later
    }
}

// ------------------------------------------------------------------------------------------------
// K6: ROQ at one coalescence time per sample (roq.py:467-549)
// ------------------------------------------------------------------------------------------------
// <h|h>_d = |K_d|^2 sum_j |h22(f_j)|^2 |C_d(f_j)|^2 w_d[j] over the quadratic nodes (roq.py:504-507)
template <int NDET, int APPROX, bool CAL, bool REDUCE = true>
__device__ __forceinline__ void bb_roq_quadratic(const double* rec, const double* cal, const BBCalGrid& grid,
                                                 const BBRoqDev& rq, int lane, double* hq) {
#pragma unroll
    for (int d = 0; d < NDET; ++d) hq[d] = 0.0;
    const int nq = rq.quad.n;
    for (int j = lane; j < nq; j += 32) {
        const double f = rq.quad.f[j];
        double A, ph;
        bb_wave_cols<APPROX>(rec, f, rq.quad.t3[j], rq.quad.x3[j], rq.quad.u7[j], rq.quad.lf[j], rq.quad.q34[j], &A, &ph);
        const double A2 = A * A;
#pragma unroll
        for (int d = 0; d < NDET; ++d) {
            double w = A2 * rq.wq[(size_t)d * nq + j];
            if (CAL) {
                double amp1, cr, ci;
                bb_cal_factor(cal + d * 4 * grid.n_points, grid.n_points, grid.l0[d], grid.inv_delta[d],
                              rq.quad.lf[j], &amp1, &cr, &ci);
                w *= amp1 * amp1;
            }
            hq[d] += w;
        }
    }
    if (REDUCE) {
#pragma unroll
        for (int d = 0; d < NDET; ++d) hq[d] = bb_warp_sum(hq[d]) * rec[BC_DET + BC_DSTRIDE * d + 3];
    }
}

// the cubic interpolation through five neighbouring ROQ times (roq.py:576-602 / 644-651, LIGO-T2100224)
__device__ __forceinline__ double2 bb_interp5(const double2* v, double a) {
    const double b = 1.0 - a;
    const double c = (a * a * a - a) / 6.0, d = (b * b * b - b) / 6.0;
    double2 o;
    {
        const double r1 = (-v[0].x + 8.0 * v[1].x - 14.0 * v[2].x + 8.0 * v[3].x - v[4].x) / 4.0;
        const double r2 = v[2].x - 2.0 * v[3].x + v[4].x;
        o.x = a * v[2].x + b * v[3].x + c * r1 + d * r2;
    }
    {
        const double r1 = (-v[0].y + 8.0 * v[1].y - 14.0 * v[2].y + 8.0 * v[3].y - v[4].y) / 4.0;
        const double r2 = v[2].y - 2.0 * v[3].y + v[4].y;
        o.y = a * v[2].y + b * v[3].y + c * r1 + d * r2;
    }
    return o;
}

// bb_interp5 as a linear combination: o = sum_k ck[k] v[k]
__device__ __forceinline__ void bb_interp5_coeffs(double a, double* ck) {
    const double b = 1.0 - a;
    const double c = (a * a * a - a) / 6.0, d = (b * b * b - b) / 6.0;
    ck[0] = -0.25 * c;
    ck[1] = 2.0 * c;
    ck[2] = a - 3.5 * c + d;
    ck[3] = b + 2.0 * c - 2.0 * d;
    ck[4] = d - 0.25 * c;
}

// K6 mapping: one warp per sample, lane = linear node 32 p + l of pass p.  W is stored a second time in node blocks,
// W2[det][p][time][32] (bb_set_roq), so the five neighbouring ROQ times of one pass are ONE contiguous 2560-byte
// piece per detector: lane 0 fetches them with cp.async.bulk into the warp's stage (completion on the warp's
// mbarrier) right after the previous pass has been consumed, and the waveform arithmetic of the pass hides the L2
// latency.  No LSU instruction, shared-memory write wavefront or register is spent on the transfer.
template <int NDET, int APPROX, bool CAL>
__global__ void __launch_bounds__(BB_ROQ_THREADS, BB_ROQ_CTAS)
bb_roq_kernel(const double* __restrict__ coef, long n, BBRoqDev rq, const double* __restrict__ calrec,
              BBCalGrid grid, double* __restrict__ out) {
    extern __shared__ __align__(16) double red_smem[];
    const int cal_len = CAL ? NDET * 4 * grid.n_points : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot_len = BC_NCOEF + cal_len;
    double* slots = red_smem + (size_t)warp * 2 * slot_len;
    double2* wst0 = reinterpret_cast<double2*>(red_smem + (size_t)BB_ROQ_WARPS * 2 * slot_len);
    double2* wst = wst0 + (size_t)warp * NDET * 5 * 32;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(wst0 + (size_t)BB_ROQ_WARPS * NDET * 5 * 32) + warp;
    if (lane == 0) {
        bb_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned parity = 0;                                     // parity of the next completion of this warp's barrier
    const int nl = rq.lin.n;
    const int n_pass = (nl + 31) / 32;
    const double* nblk = rq.lin.blk + lane;
    const double ts0 = (double)rq.time_start_index * rq.time_step;
    const double ts1 = (double)(rq.time_start_index + 1) * rq.time_step;
    const double space = ts1 - ts0;                          // samples[1] - samples[0] (roq.py:571)
    const long stride = (long)gridDim.x * BB_ROQ_WARPS;
    long s = (long)blockIdx.x * BB_ROQ_WARPS + warp;
    if (s < n) bb_red_prefetch<CAL>(slots, slots + BC_NCOEF, coef, calrec, s, cal_len, lane);
    for (int ping = 0; s < n; s += stride, ping ^= 1) {
        double* rec = slots + ping * slot_len;
        double* cal = rec + BC_NCOEF;
        bb_red_wait();
        if (s + stride < n) {
            double* nxt = slots + (ping ^ 1) * slot_len;
            bb_red_prefetch<CAL>(nxt, nxt + BC_NCOEF, coef, calrec, s + stride, cal_len, lane);
        }
        // five neighbouring ROQ times per detector (roq.py:509-516, 551-574): first = closest - 2, clipped per index.
        // Lane d works out detector d (window, interpolation coefficients: four FP64 divisions) and the warp picks the
        // results up by shuffle, instead of every lane working out every detector.
        const int dl = lane < NDET ? lane : NDET - 1;
        int first_m;
        bool inb_m;
        double ck_m[5];
        {
            const double ifo_time = rec[BC_DT0] + 0.5 * rec[BC_DET + BC_DSTRIDE * dl + 2];
            const double q = floor((ifo_time - ts0) / space);
            const long closest = (long)fmin(fmax(q, -1.0e9), 1.0e9);
            inb_m = (closest - 2 >= 0) && (closest + 2 < (long)rq.n_time);
            first_m = (int)(closest - 2);
            // a = (time_samples[3] - time) / max(time_samples[1] - time_samples[0], 1e-12) on the CLIPPED indices
            const int i3 = min(max(first_m + 3, 0), rq.n_time - 1), i1 = min(max(first_m + 1, 0), rq.n_time - 1),
                      i0 = min(max(first_m, 0), rq.n_time - 1);
            const double t3 = (double)(rq.time_start_index + i3) * rq.time_step;
            const double t1 = (double)(rq.time_start_index + i1) * rq.time_step;
            const double t0 = (double)(rq.time_start_index + i0) * rq.time_step;
            const double a = (t3 - ifo_time) / fmax(t1 - t0, 1e-12);
            // The five-sample interpolation (bb_interp5) is linear in the five contractions with REAL coefficients that
            // depend only on the sample's time: the five rows of W are combined first (10 DFMA) and contracted once
            // (4 DFMA) instead of five complex multiply-accumulates (20 DFMA) and 30 running sums per lane.
            bb_interp5_coeffs(a, ck_m);
        }
        const unsigned inb_mask = __ballot_sync(0xffffffffu, inb_m);
        const bool all_in = (inb_mask & ((1u << NDET) - 1u)) == ((1u << NDET) - 1u);
        int first[NDET];
        double ck[NDET][5];
#pragma unroll
        for (int d = 0; d < NDET; ++d) {
            first[d] = __shfl_sync(0xffffffffu, first_m, d);
#pragma unroll
            for (int k = 0; k < 5; ++k) ck[d][k] = __shfl_sync(0xffffffffu, ck_m[k], d);
        }
        // Source of pass 0 and the advance per pass, computed once per sample.  The usual case (all five times of every
        // detector inside the weights): lane d issues the one bulk copy of detector d from its own running pointer
        // (the copy's operands travel through uniform registers, one elected lane at a time, so three lanes with one
        // copy each are a third of the instructions of one lane with three)
        const double2* wsrc = rq.W2 + ((size_t)dl * n_pass * rq.n_time + (size_t)max(first_m, 0)) * 32;
        double2* const wdst = wst + dl * 5 * 32;
        const size_t wstride = (size_t)rq.n_time * 32;
        auto stage_fill = [&](int pass) {
            if (all_in) {
                if (lane < NDET) {
                    if (lane == 0) bb_mbar_expect_tx(bar, (unsigned)(NDET * 5 * 32 * sizeof(double2)));
                    bb_bulk_g2s(wdst, wsrc, 5 * 32 * sizeof(double2), bar);
                }
                wsrc += wstride;
            } else if (lane == 0) {
                bb_mbar_expect_tx(bar, (unsigned)(NDET * 5 * 32 * sizeof(double2)));
                for (int d = 0; d < NDET; ++d) {
                    const double2* blk = rq.W2 + ((size_t)d * n_pass + pass) * rq.n_time * 32;
                    for (int k = 0; k < 5; ++k) {
                        const int i = min(max(first[d] + k, 0), rq.n_time - 1);
                        bb_bulk_g2s(wst + (d * 5 + k) * 32, blk + (size_t)i * 32, 32 * sizeof(double2), bar);
                    }
                }
            }
        };
        __syncwarp();
        stage_fill(0);
        double hq[NDET];
        bb_roq_quadratic<NDET, APPROX, CAL, NDET != 3>(rec, cal, grid, rq, lane, hq);      // NDET == 3: lane partials
        double2 acc[NDET];
#pragma unroll
        for (int d = 0; d < NDET; ++d) acc[d] = make_double2(0.0, 0.0);
        for (int pass = 0; pass < n_pass; ++pass) {
            // lanes beyond the last node evaluate the last node again (no branch): their rows of W2 are zero
            double zr0, zi0, lfj;
            {
                const double* nb = nblk + pass * BB_NB_ROW;
                const double f = nb[BB_NB_F];
                double A, ph, sn, cs;
                lfj = nb[BB_NB_LF];
                bb_wave_cols<APPROX>(rec, f, nb[BB_NB_T3], nb[BB_NB_X3], nb[BB_NB_U7], lfj,
                                     (APPROX == BB_IMRPHENOMD) ? nb[BB_NB_Q34] : 0.0, &A, &ph);
                bb_sincospi(ph, &sn, &cs);
                zr0 = A * cs;                                   // conj(h22) = A e^{+i Phi}
                zi0 = A * sn;
            }
            bb_mbar_wait(bar, parity);
            parity ^= 1u;
            const double2* wrow = wst + lane;                   // W2 is zero beyond the last node
#pragma unroll
            for (int d = 0; d < NDET; ++d) {
                double zr = zr0, zi = zi0;
                if (CAL) {
                    double amp1, cr, ci;                    // conj(h C) = conj(h) amp1 (cr - i ci)
                    bb_cal_factor(cal + d * 4 * grid.n_points, grid.n_points, grid.l0[d], grid.inv_delta[d],
                                  lfj, &amp1, &cr, &ci);
                    const double tr = amp1 * (zr * cr + zi * ci), ti = amp1 * (zi * cr - zr * ci);
                    zr = tr;
                    zi = ti;
                }
                double wx = 0.0, wy = 0.0;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const double2 w = wrow[32 * (d * 5 + k)];
                    wx = fma(ck[d][k], w.x, wx);
                    wy = fma(ck[d][k], w.y, wy);
                }
                acc[d].x = fma(zr, wx, fma(-zi, wy, acc[d].x));
                acc[d].y = fma(zr, wy, fma(zi, wx, acc[d].y));
            }
            __syncwarp();                                       // every lane has read the stage
            if (pass + 1 < n_pass) stage_fill(pass + 1);
        }
        if (NDET == 3) {
            // nine sums (3 x Re / Im of the contraction, 3 quadratic) through the halving butterfly: lane 4 q ends with
            // quantity q of {Re0, Im0, Re1, Im1, Re2, Im2, Q0, Q1}; Q2 by a plain reduction
            const double v[8] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y, acc[2].x, acc[2].y, hq[0], hq[1]};
            const double t8 = bb_warp_sum8(v, lane);
            const double q2 = bb_warp_sum(hq[2]);
            const double im = __shfl_down_sync(0xffffffffu, t8, 4);          // lane 8 d: Re_d here, Im_d from lane 8 d + 4
            const bool bad = rec[BC_STATUS] != 0.0;
            double* o = out + s * NDET * 3;
            if ((lane & 7) == 0 && lane < 24) {
                const int d = lane >> 3;
                const double kr = rec[BC_DET + BC_DSTRIDE * d], ki = rec[BC_DET + BC_DSTRIDE * d + 1];
                const bool in = (inb_mask >> d) & 1u;
                // conj(K) * sum; out of the ROQ time window: d_inner_h += log(False) (roq.py:532-533)
                o[3 * d] = in ? kr * t8 + ki * im : -INFINITY;
                o[3 * d + 1] = kr * im - ki * t8;
            }
            if (lane == 24 || lane == 28) {
                const int d = (lane - 24) >> 2;
                o[3 * d + 2] = bad ? nan("") : t8 * rec[BC_DET + BC_DSTRIDE * d + 3];
            }
            if (lane == 1) o[8] = bad ? nan("") : q2 * rec[BC_DET + BC_DSTRIDE * 2 + 3];
        } else {
#pragma unroll
            for (int d = 0; d < NDET; ++d) {
                const double kr = rec[BC_DET + BC_DSTRIDE * d], ki = rec[BC_DET + BC_DSTRIDE * d + 1];
                const double sr = bb_warp_sum(acc[d].x), si = bb_warp_sum(acc[d].y);
                if (lane == 0) {
                    double* o = out + (s * NDET + d) * 3;
                    // conj(K) * sum; out of the ROQ time window: d_inner_h += log(False) (roq.py:532-533)
                    o[0] = ((inb_mask >> d) & 1u) ? kr * sr + ki * si : -INFINITY;
                    o[1] = kr * si - ki * sr;
                    o[2] = (rec[BC_STATUS] != 0.0) ? nan("") : hq[d];
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// K7: ROQ time marginalisation (roq.py:535-549, 604-651; base.py:794-820)
//   (a) bb_roq_hlinear_kernel : V_d[s][i] = conj(h_linear_d[i]) for every sample, <h|h> per sample
//   (b) bb_gemm_nt_kernel (DMMA): Y_d[s][t] = sum_i W_d[t][i] V_d[s][i]      (the dense all-times contraction)
//   (c) bb_roq_time_marg_kernel: five-sample interpolation at the likelihood's time grid, sum over detectors,
//                                point likelihood, logsumexp with the time prior
// ------------------------------------------------------------------------------------------------
template <int NDET, int APPROX, bool CAL>
__global__ void __launch_bounds__(BB_RED_THREADS)
bb_roq_hlinear_kernel(const double* __restrict__ coef, long s_begin, long n, BBRoqDev rq,
                      const double* __restrict__ calrec, BBCalGrid grid, double2* __restrict__ V /* [NDET] packed [n x n_lin] */,
                      double* __restrict__ hh /* [n] */) {
    extern __shared__ __align__(16) double red_smem[];
    const int cal_len = CAL ? NDET * 4 * grid.n_points : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* rec = red_smem + (size_t)warp * (BC_NCOEF + cal_len);
    double* cal = rec + BC_NCOEF;
    const int nl = rq.lin.n;
    const long S = (nl + 15) / 16;                                   // V is written packed: the GEMM's A operand
    const size_t vstride = (size_t)((n + 63) / 64) * S * 64 * BB_PK;
    for (long s = (long)blockIdx.x * BB_RED_WARPS + warp; s < n; s += (long)gridDim.x * BB_RED_WARPS) {
        bb_red_load<CAL>(rec, cal, coef, calrec, s_begin + s, cal_len, lane);
        if (nl + lane < S * 16)                                      // zero K tail of the last slab
#pragma unroll
            for (int d = 0; d < NDET; ++d) V[d * vstride + bb_pk(s, nl + lane, 64, S)] = make_double2(0.0, 0.0);
        double hq[NDET];
        bb_roq_quadratic<NDET, APPROX, CAL>(rec, cal, grid, rq, lane, hq);
        if (lane == 0) {
            double t = 0.0;
#pragma unroll
            for (int d = 0; d < NDET; ++d) t += hq[d];
            hh[s] = (rec[BC_STATUS] != 0.0) ? nan("") : t;
        }
        for (int j = lane; j < nl; j += 32) {
            const double f = rq.lin.f[j];
            double A, ph;
            bb_wave_cols<APPROX>(rec, f, rq.lin.t3[j], rq.lin.x3[j], rq.lin.u7[j], rq.lin.lf[j], rq.lin.q34[j], &A, &ph);
            double sn, cs;
            bb_sincospi(ph, &sn, &cs);
            const double zr0 = A * cs, zi0 = A * sn;
#pragma unroll
            for (int d = 0; d < NDET; ++d) {
                double zr = zr0, zi = zi0;
                if (CAL) {
                    double amp1, cr, ci;
                    bb_cal_factor(cal + d * 4 * grid.n_points, grid.n_points, grid.l0[d], grid.inv_delta[d],
                                  rq.lin.lf[j], &amp1, &cr, &ci);
                    const double tr = amp1 * (zr * cr + zi * ci), ti = amp1 * (zi * cr - zr * ci);
                    zr = tr;
                    zi = ti;
                }
                const double kr = rec[BC_DET + BC_DSTRIDE * d], ki = rec[BC_DET + BC_DSTRIDE * d + 1];
                V[d * vstride + bb_pk(s, j, 64, S)] = make_double2(kr * zr + ki * zi, kr * zi - ki * zr);   // conj(K) conj(h)
            }
        }
    }
}

template <int NDET>
__global__ void __launch_bounds__(BB_RED_THREADS)
bb_roq_time_marg_kernel(const double* __restrict__ coef, long s_begin, long n, BBRoqDev rq,
                        const double2* __restrict__ Y /* [NDET][n][n_row]: ROQ rows row0 .. row0 + n_row - 1 */, int row0,
                        int n_row, const double* __restrict__ hh,
                        BBMarg marg, double start_time, double* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double ts0 = (double)rq.time_start_index * rq.time_step;
    const double ts1 = (double)(rq.time_start_index + 1) * rq.time_step;
    const double space = ts1 - ts0;
    const int wpb = blockDim.x >> 5;          // launched with one warp per CTA so that it fits beside the GEMM's CTAs
    for (long s = (long)blockIdx.x * wpb + warp; s < n; s += (long)gridDim.x * wpb) {
        const double* rec = coef + (s_begin + s) * BC_NCOEF;
        if (rec[BC_STATUS] != 0.0) {
            if (lane == 0) out[s_begin + s] = -DBL_MAX;
            continue;
        }
        const double h2 = hh[s], dist = rec[BC_DISTANCE];
        const double jit = marg.jitter ? rec[BC_JITTER] : 0.0;
        double delay[NDET];
#pragma unroll
        for (int d = 0; d < NDET; ++d) delay[d] = 0.5 * rec[BC_DET + BC_DSTRIDE * d + 2];
        const double bw = marg.roq_dtc / (marg.time_max - marg.time_min);     // prior.prob(t) * delta_tc
        double mx = -INFINITY, sum = 0.0;
        for (int j = lane; j < rq.n_marg; j += 32) {
            const double tj = rq.marg_start + rq.marg_dtc / 2.0 + (double)j * rq.marg_dtc;   // roq.py:330
            const double tt = tj + jit;                                         // base.py:795-797
            if (tt < marg.time_min || tt > marg.time_max) continue;             // base.py:799-806
            double dre = 0.0, dim = 0.0;
#pragma unroll
            for (int d = 0; d < NDET; ++d) {
                double ifo_t = (tj - start_time) + delay[d];                     // roq.py:536-539
                if (marg.jitter) ifo_t += jit;
                const double per = (ifo_t - ts0) / space;
                const double fl = floor(per);
                long c = (long)fl;
                const double2* y = Y + ((size_t)d * n + s) * n_row - row0;
                double2 v[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    long i = c + k - 2;
                    i = i < 0 ? 0 : (i > rq.n_time - 1 ? rq.n_time - 1 : i);
                    v[k] = y[i];
                }
                const double b = per - fl;
                const double2 r = bb_interp5(v, 1.0 - b);
                dre += r.x;
                dim += r.y;
            }
            const double l = bb_point_lnl(marg, dre, dim, h2, dist);
            if (l == -INFINITY) continue;
            if (l > mx) { sum = sum * exp(mx - l) + bw; mx = l; }
            else sum += bw * exp(l - mx);
        }
        double gmx = mx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmx = fmax(gmx, __shfl_xor_sync(0xffffffffu, gmx, o));
        double part = (mx == -INFINITY) ? 0.0 : sum * exp(mx - gmx);
        part = bb_warp_sum(part);
        if (lane == 0) out[s_begin + s] = (gmx == -INFINITY) ? -INFINITY : log(part) + gmx;
    }
}

// ------------------------------------------------------------------------------------------------
// Multi-banded likelihood with time marginalisation (multiband.py:714-726, 789-797).  The reference scatters
// strain * linear_coeffs of every detector into a zero array of Nbs[-1] / 2 points and takes its FFT; only the banded
// points are non-zero (1e4 of 2.6e5 for the 128 s signal) and only the times inside the geocent_time prior are used
// (a few hundred), so the transform is evaluated as a dense contraction instead:
//   (a) bb_mb_series_kernel   : v_s[p] = sum_det h_det(f_p) L_det[p] for every sample (packed, the GEMM's A operand),
//                               <h|h> = sum_det sum_p Q_det[p] |h_det(f_p)|^2
//   (b) bb_gemm_nt_kernel     : D[s][t] = sum_p v_s[p] exp(-2 pi i idx_p (row0 + t) / N)      (FP64 tensor cores)
//   (c) bb_mb_time_marg_kernel: point likelihood per time, logsumexp with the time prior (base.py:794-820)
// ------------------------------------------------------------------------------------------------
template <int NDET, int APPROX, bool CAL>
__global__ void __launch_bounds__(BB_RED_THREADS)
bb_mb_series_kernel(const double* __restrict__ coef, long s_begin, long n, BBRelbinDev rb,
                    const double* __restrict__ calrec, BBCalGrid grid, double2* __restrict__ V /* packed [n x n_points] */,
                    double* __restrict__ hh /* [n] */) {
    extern __shared__ __align__(16) double red_smem[];
    const int cal_len = CAL ? NDET * 4 * grid.n_points : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* rec = red_smem + (size_t)warp * (BC_NCOEF + cal_len);
    double* cal = rec + BC_NCOEF;
    const int ne = rb.edges.n, np = rb.ne_pad;
    const long S = (ne + 15) / 16;
    for (long s = (long)blockIdx.x * BB_RED_WARPS + warp; s < n; s += (long)gridDim.x * BB_RED_WARPS) {
        bb_red_load<CAL>(rec, cal, coef, calrec, s_begin + s, cal_len, lane);
        double h2 = 0.0;
        for (int base = 0; base < S * 16; base += 32) {
            const int j = base + lane;
            if (j >= S * 16) break;
            double vr = 0.0, vi = 0.0;
            if (j < ne) {
                const double f = rb.edges.f[j], lfj = rb.edges.lf[j];
                double A, ph;
                bb_wave_cols<APPROX>(rec, f, rb.edges.t3[j], rb.edges.x3[j], rb.edges.u7[j], lfj, rb.edges.q34[j], &A, &ph);
#pragma unroll
                for (int d = 0; d < NDET; ++d) {
                    const double* cd = rec + BC_DET + BC_DSTRIDE * d;
                    double sn, cs;
                    bb_sincospi(ph + cd[2] * f, &sn, &cs);
                    double hr = A * (cd[0] * cs + cd[1] * sn), hi = A * (cd[1] * cs - cd[0] * sn);   // K h
                    if (CAL) {
                        double amp1, cr, ci;
                        bb_cal_factor(cal + d * 4 * grid.n_points, grid.n_points, grid.l0[d], grid.inv_delta[d], lfj,
                                      &amp1, &cr, &ci);
                        const double tr = amp1 * (hr * cr - hi * ci), ti = amp1 * (hr * ci + hi * cr);
                        hr = tr;
                        hi = ti;
                    }
                    const double2 c = rb.lin_c[(size_t)d * np + j];         // conj(L)
                    vr = fma(hr, c.x, fma(hi, c.y, vr));                      // h L
                    vi = fma(hi, c.x, fma(-hr, c.y, vi));
                    h2 = fma(rb.quad_e[(size_t)d * np + j], fma(hr, hr, hi * hi), h2);
                }
            }
            V[bb_pk(s, j, 64, S)] = make_double2(vr, vi);                    // zero K tail included
        }
        h2 = bb_warp_sum(h2);
        if (lane == 0) hh[s] = (rec[BC_STATUS] != 0.0) ? nan("") : h2;
    }
}

// E[t][p] = exp(-2 pi i idx_p (row0 + t) / N), packed as the GEMM's B operand (phases reduced modulo N in integers: exact)
__global__ void bb_mb_phase_kernel(const int* __restrict__ idx, int n_points, long row0, int n_row, long n_full,
                                   double2* __restrict__ E) {
    const long S = (n_points + 15) / 16;
    const long total = (long)n_row * S * 16;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long t = i / (S * 16), p = i - t * (S * 16);
        double2 v = make_double2(0.0, 0.0);
        if (p < n_points) {
            const long m = ((long)idx[p] * (row0 + t)) % n_full;
            double sn, cs;
            sincospi(-2.0 * (double)m / (double)n_full, &sn, &cs);
            v = make_double2(cs, sn);
        }
        E[bb_pk(t, p, BB_GEMM_TR_B, S)] = v;
    }
}

__global__ void __launch_bounds__(128)
bb_mb_time_marg_kernel(const double* __restrict__ coef, long s_begin, long n, const double2* __restrict__ Y /* [n][n_row] */,
                       long row0, int n_row, long n_full, double dtc, const double* __restrict__ hh, BBMarg marg,
                       double start_time, double* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (long s = (long)blockIdx.x * wpb + warp; s < n; s += (long)gridDim.x * wpb) {
        const double* rec = coef + (s_begin + s) * BC_NCOEF;
        if (rec[BC_STATUS] != 0.0) {
            if (lane == 0) out[s_begin + s] = -DBL_MAX;
            continue;
        }
        const double h2 = hh[s], dist = rec[BC_DISTANCE];
        const double jit = marg.jitter ? rec[BC_JITTER] : 0.0;
        const double bw = dtc / (marg.time_max - marg.time_min);              // prior.prob(t) * delta_tc
        // times start + j dtc (+ jitter) inside the prior (base.py:795-806): j in [j_lo, j_hi]
        long j_lo = (long)ceil((marg.time_min - jit - start_time) / dtc) - 1, j_hi = (long)floor((marg.time_max - jit - start_time) / dtc) + 1;
        if (j_lo < 0) j_lo = 0;
        if (j_hi > n_full - 1) j_hi = n_full - 1;
        bool missing = false;
        double mx = -INFINITY, sum = 0.0;
        for (long j = j_lo + lane; j <= j_hi; j += 32) {
            const double tt = (start_time + (double)j * dtc) + jit;
            if (tt < marg.time_min || tt > marg.time_max) continue;
            if (j < row0 || j >= row0 + n_row) { missing = true; continue; }   // outside the contracted window
            const double2 d = Y[(size_t)s * n_row + (j - row0)];
            const double l = bb_point_lnl(marg, d.x, d.y, h2, dist);
            if (l == -INFINITY) continue;
            if (l > mx) { sum = sum * exp(mx - l) + bw; mx = l; }
            else sum += bw * exp(l - mx);
        }
        double gmx = mx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmx = fmax(gmx, __shfl_xor_sync(0xffffffffu, gmx, o));
        double part = (mx == -INFINITY) ? 0.0 : sum * exp(mx - gmx);
        part = bb_warp_sum(part);
        missing = __any_sync(0xffffffffu, missing);
        if (lane == 0) out[s_begin + s] = missing ? nan("") : ((gmx == -INFINITY) ? -INFINITY : log(part) + gmx);
    }
}

// ------------------------------------------------------------------------------------------------
// Multi-banded likelihood, IFFT-FFT form of (h, h) (multiband.py:613-646, 766-787; linear_interpolation=False).
// Per band b >= 1 the reference takes the inverse real transform (M^(b) points) of sqrt(window) * strain, zero-pads it to
// 2 M^(b) and sums |rfft|^2 against the transform I of the truncated inverse-PSD autocorrelation.  The EVEN bins of that
// spectrum are the band's points themselves, so their terms (like all of band 0) are per-point weights and ride in
// quad_e with the linear-interpolation kernel; only the ODD bins need transforms:
//   Z[m]   = FFT(conj(w))[m],  w = sqrt(window) strain at bins Ks..Ke   -> x[m] = (2 / M) Re Z[m]   (the irfft)
//   Y[l]   = FFT(x[m] e^{-i pi m / M})[l] = X[2 l + 1]
//   <h|h> += (4 / That) sum_l |Y[l]|^2 I[2 l + 1],   l < M / 2
// (bb_fft.cuh for the transforms; these kernels fill, modulate and reduce).
// ------------------------------------------------------------------------------------------------
struct BBMbBand {
    int M, log2M, Ks, Ke, start;      // transform length, band bins, index of the band's first banded point
    double norm;                      // 4 / That^(b)
    const double* i_odd;              // [n_det][M / 2]   I^(b)[2 l + 1]
};

template <int NDET, int APPROX, bool CAL>
__global__ void __launch_bounds__(BB_RED_THREADS)
bb_mb_band_fill_kernel(const double* __restrict__ coef, long s_begin, long n, BBRelbinDev rb, BBMbBand band,
                       const double* __restrict__ sqrt_window, const double* __restrict__ calrec, BBCalGrid grid,
                       double2* __restrict__ Z /* [n][NDET][M], zeroed */) {
    extern __shared__ __align__(16) double red_smem[];
    const int cal_len = CAL ? NDET * 4 * grid.n_points : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* rec = red_smem + (size_t)warp * (BC_NCOEF + cal_len);
    double* cal = rec + BC_NCOEF;
    const int nb = band.Ke - band.Ks + 1;
    for (long s = (long)blockIdx.x * BB_RED_WARPS + warp; s < n; s += (long)gridDim.x * BB_RED_WARPS) {
        bb_red_load<CAL>(rec, cal, coef, calrec, s_begin + s, cal_len, lane);
        for (int q = lane; q < nb; q += 32) {
            const int j = band.start + q;
            const double f = rb.edges.f[j], lfj = rb.edges.lf[j], sw = sqrt_window[j];
            double A, ph;
            bb_wave_cols<APPROX>(rec, f, rb.edges.t3[j], rb.edges.x3[j], rb.edges.u7[j], lfj, rb.edges.q34[j], &A, &ph);
#pragma unroll
            for (int d = 0; d < NDET; ++d) {
                const double* cd = rec + BC_DET + BC_DSTRIDE * d;
                double sn, cs;
                bb_sincospi(ph + cd[2] * f, &sn, &cs);
                double hr = A * (cd[0] * cs + cd[1] * sn), hi = A * (cd[1] * cs - cd[0] * sn);   // K h
                if (CAL) {
                    double amp1, cr, ci;
                    bb_cal_factor(cal + d * 4 * grid.n_points, grid.n_points, grid.l0[d], grid.inv_delta[d], lfj,
                                  &amp1, &cr, &ci);
                    const double tr = amp1 * (hr * cr - hi * ci), ti = amp1 * (hr * ci + hi * cr);
                    hr = tr;
                    hi = ti;
                }
                Z[((size_t)s * NDET + d) * band.M + band.Ks + q] = make_double2(sw * hr, -sw * hi);     // conj(w)
            }
        }
        __syncwarp();
    }
}

// y[m] = (2 / M) Re Z[m] e^{-i pi m / M}, in place
__global__ void bb_mb_band_modulate_kernel(double2* __restrict__ Z, long total, int M) {
    const double scale = 2.0 / (double)M;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int m = (int)(i & (M - 1));
        double sn, cs;
        sincospi(-(double)m / (double)M, &sn, &cs);
        const double x = scale * Z[i].x;
        Z[i] = make_double2(x * cs, x * sn);
    }
}

__global__ void bb_mb_hh_fold_kernel(const double* __restrict__ per_det, long n, int n_det, double* __restrict__ hh) {
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    double t = 0.0;
    for (int d = 0; d < n_det; ++d) t += per_det[s * n_det + d];
    hh[s] += t;
}

// one warp per (sample, detector): target[(s * NDET + d) * stride + offset] += norm sum_l |Y[l]|^2 I_odd[d][l]
__global__ void bb_mb_band_reduce_kernel(const double2* __restrict__ Y, long n, int n_det, BBMbBand band,
                                         double* __restrict__ target, int stride, int offset) {
    const int lane = threadIdx.x & 31;
    const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n * n_det) return;
    const int d = (int)(w % n_det);
    const double2* y = Y + (size_t)w * band.M;
    const double* io = band.i_odd + (size_t)d * (band.M / 2);
    double acc = 0.0;
    for (int l = lane; l < band.M / 2; l += 32) acc = fma(fma(y[l].x, y[l].x, y[l].y * y[l].y), io[l], acc);
    acc = bb_warp_sum(acc);
    if (lane == 0) target[(size_t)w * stride + offset] += band.norm * acc;
}

// ------------------------------------------------------------------------------------------------
// K5t: relative binning with time marginalisation (relative.py:380-421): the full-grid waveform is rebuilt
// as h0_d[k] (r0_b + r1_b (f_k - f_centre,b)), so the series h conj(d)/S is P_d[k] (r0 + r1 (f_k - fc)) with
// P_d = (4/T) h0_d conj(d_d) / S_d precomputed at set-up.  One CTA per sample; warp 0 evaluates the edges.
// ------------------------------------------------------------------------------------------------
template <int NDET, int APPROX, bool CAL>
__global__ void __launch_bounds__(BB_TM_THREADS, 1)
bb_relbin_time_marg_kernel(const double* __restrict__ coef, long n, BBRelbinDev rb, int n_freq, double df, int nfft,
                           int log2n, const double2* __restrict__ twiddle, BBMarg marg, double start_time,
                           double duration, const double* __restrict__ calrec, BBCalGrid grid,
                           double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* X = reinterpret_cast<double2*>(smem_raw);
    const int nb = rb.edges.n - 1;
    int plan_a, plan_b, ps;
    bb_tm_plan(log2n, &plan_a, &plan_b, &ps);
    double2* r01 = X + bb_tm_series_elems(nfft, ps);             // [nb][NDET][2]
    double* c = reinterpret_cast<double*>(r01 + (size_t)nb * NDET * 2);
    double* red = c + BC_NCOEF;      // [33]
    double* cal = red + 33;
    const int cal_len = CAL ? NDET * 4 * grid.n_points : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long s = blockIdx.x; s < n; s += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < BC_NCOEF; i += BB_TM_THREADS) c[i] = coef[s * BC_NCOEF + i];
        if (CAL) for (int i = tid; i < cal_len; i += BB_TM_THREADS) cal[i] = calrec[s * cal_len + i];
        __syncthreads();
        if (c[BC_STATUS] != 0.0) {
            if (tid == 0) out[s] = -DBL_MAX;
            continue;
        }
        if (warp == 0) {
            double acc[NDET][3];
#pragma unroll
            for (int d = 0; d < NDET; ++d) acc[d][0] = acc[d][1] = acc[d][2] = 0.0;
            bb_relbin_sample<NDET, APPROX, CAL, true>(c, cal, grid, rb, lane, acc, r01);
            double hh = 0.0;
#pragma unroll
            for (int d = 0; d < NDET; ++d) hh += bb_warp_sum(acc[d][2]);
            if (lane == 0) red[32] = hh;
        }
        __syncthreads();
        const double hh = red[32];
        for (int k = tid; k < nfft; k += BB_TM_THREADS) {        // Nyquist bin dropped (relative.py:418-420)
            const int b = (k < n_freq) ? rb.bin_of_k[k] : -1;
            double vr = 0.0, vi = 0.0;
            if (b >= 0) {
                const double df_c = (double)k * df - rb.centre[b];
#pragma unroll
                for (int d = 0; d < NDET; ++d) {
                    const double2 r0 = r01[((size_t)b * NDET + d) * 2], r1 = r01[((size_t)b * NDET + d) * 2 + 1];
                    const double qr = r0.x + r1.x * df_c, qi = r0.y + r1.y * df_c;
                    const double2 p = rb.pgrid[(size_t)d * n_freq + k];
                    vr += p.x * qr - p.y * qi;
                    vi += p.x * qi + p.y * qr;
                }
            }
            X[bb_tm_pos(k, ps)] = make_double2(vr, vi);
        }
        __syncthreads();
        bb_tm_fft_dif(X, nfft, log2n, twiddle);
        bb_tm_finish(X, nfft, log2n, marg, hh, c[BC_DISTANCE], c[BC_JITTER], start_time, duration, red, out + s);
    }
}
