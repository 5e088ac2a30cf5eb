// TaylorF2 (+ tidal terms) per-sample prologue and per-bin evaluation.
//
// Replaces the lalsimulation call behind bilby/gw/source.py:351-432 (lal_binary_neutron_star with
// waveform_approximant = "TaylorF2"): 3.5PN aligned-spin phasing (bb_taylorf2_phasing in bb_phenomd.cuh),
// tidal 5/6/6.5/7/7.5PN terms, Newtonian amplitude, exp(-i(Phi - pi/4)), reference-phase subtraction.
// Same record layout idea as IMRPhenomD: coefficients of powers of x = f^(1/3) [Hz^(1/3)], t = 1/x and ln f,
// phases in half turns, geocentric time shift folded into the linear term.
#pragma once
#include "bb_common.cuh"
#include "bb_geometry.cuh"
#include "bb_phenomd.cuh"

// record slots (reuse the phase area of the IMRPhenomD record)
enum {
    BT_P = BC_PINS,   // 15: 1, x, x^2, x^3(=f), x^5, x^7, x^8, x^9, x^10, t, t^2, t^3, t^5, ln f, x ln f
    BT_NP = 15
};

BB_HD double bb_tf2_series(double v, const double* pv, const double* pvl, const double* tid) {
    const double logv = log(v);
    const double v2 = v * v, v3 = v2 * v, v4 = v3 * v, v5 = v4 * v, v6 = v5 * v, v7 = v6 * v;
    double ph = pv[7] * v7 + (pv[6] + pvl[6] * logv) * v6 + (pv[5] + pvl[5] * logv) * v5 + pv[4] * v4
                + pv[3] * v3 + pv[2] * v2 + pv[1] * v + pv[0];
    const double v10 = v5 * v5;
    ph += tid[0] * v10 + tid[1] * v10 * v2 + tid[2] * v10 * v3 + tid[3] * v10 * v4 + tid[4] * v10 * v5;
    return ph / v5;
}

template <bool WHOLE = true>      // false: the waveform part only (see bb_phenomd_prologue)
BB_HD void bb_taylorf2_prologue(const double* p, const BBNetwork& net, const BBWaveformConfig& wf, double* coef) {
    if (WHOLE) for (int i = 0; i < BC_NCOEF; ++i) coef[i] = 0.0;
    const double m1 = p[BB_P_MASS_1], m2 = p[BB_P_MASS_2], chi1 = p[BB_P_CHI_1], chi2 = p[BB_P_CHI_2];
    const double lam1 = p[BB_P_LAMBDA_1], lam2 = p[BB_P_LAMBDA_2];
    const double dist_mpc = p[BB_P_DISTANCE];
    coef[BC_DISTANCE] = dist_mpc;
    coef[BC_JITTER] = p[BB_P_TIME_JITTER];
    const double dt0 = WHOLE ? bb_detector_prologue(p, net, wf, coef) : bb_prologue_dt0(p, net, wf);
    const double M = m1 + m2;
    const double MTSUN = BB_G_SI * BB_MSUN_SI / (BB_C_SI * BB_C_SI * BB_C_SI);
    const double MRSUN = BB_G_SI * BB_MSUN_SI / (BB_C_SI * BB_C_SI);
    const double Ms = M * MTSUN;
    const double pi = BB_PI;
    const double piM = pi * Ms;
    const double f_isco = (1.0 / sqrt(6.0)) * (1.0 / 6.0) / piM;      // vISCO^3 / (pi M)
    const double f_end = (wf.f_max == 0.0) ? f_isco : wf.f_max;
    const bool bad = !(m1 > 0.0) || !(m2 > 0.0) || !(dist_mpc > 0.0) || (!wf.sequence && !(f_end > wf.f_min)) || !isfinite(M)
                     || !isfinite(dist_mpc);
    if (bad) {
        coef[BC_STATUS] = 1.0;
        return;
    }
    const double eta = m1 * m2 / (M * M);
    double pv[8], pvl[8];
    bb_taylorf2_phasing(m1, m2, chi1, chi2, 1.0, 1.0, pv, pvl);
    const double pfaN = 3.0 / (128.0 * eta);
    double tid[5];
    {
        const double xs[2] = {m1 / M, m2 / M};
        const double ls[2] = {lam1, lam2};
        for (int i = 0; i < 5; ++i) tid[i] = 0.0;
        for (int b = 0; b < 2; ++b) {
            const double x = xs[b], x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
            tid[0] += ls[b] * (-288.0 + 264.0 * x) * x4;
            tid[1] += ls[b] * (-15895.0 / 28.0 + 4595.0 / 28.0 * x + 5715.0 / 14.0 * x2 - 325.0 / 7.0 * x3) * x4;
            tid[2] += ls[b] * x4 * 24.0 * (12.0 - 11.0 * x) * pi;
            tid[3] += ls[b] * (-x4 * 5.0 * (193986935.0 / 571536.0 - 14415613.0 / 381024.0 * x - 57859.0 / 378.0 * x2
                                           - 209495.0 / 1512.0 * x3 + 965.0 / 54.0 * x4 - 4.0 * x4 * x));
            tid[4] += ls[b] * x4 * 1.0 / 28.0 * pi * (27719.0 - 22415.0 * x + 7598.0 * x2 - 10520.0 * x3);
        }
        for (int i = 0; i < 5; ++i) tid[i] *= pfaN;
    }
    double ref_phasing = 0.0;
    if (wf.f_ref != 0.0) ref_phasing = bb_tf2_series(cbrt(piM * wf.f_ref), pv, pvl, tid);
    // h = amp e^{-i(Phi - pi/4)}, Phi(f) = series(v)/v^5 - 2 phi_ref - ref_phasing; plus geocentric shift 2 pi f dt0
    const double cst = -2.0 * p[BB_P_PHASE] - ref_phasing - pi / 4.0;
    const double lin = 2.0 * pi * dt0;
    const double ipi = 1.0 / pi;
    const double a = cbrt(piM);            // v = a x
    const double la = log(a);              // logv = la + ln(f)/3
    double ap[16];
    ap[0] = 1.0;
    for (int i = 1; i < 16; ++i) ap[i] = ap[i - 1] * a;
    double* q = coef + BT_P;
    // series(v)/v^5 = sum_k (pv_k + pvl_k logv) v^(k-5) + tidal
    q[0] = (pv[5] + pvl[5] * la + cst) * ipi;                 // k = 5
    q[1] = (pv[6] + pvl[6] * la) * ap[1] * ipi;               // k = 6 -> v
    q[2] = pv[7] * ap[2] * ipi;                               // k = 7 -> v^2
    q[3] = lin * ipi;                                         // f = x^3
    q[4] = tid[0] * ap[5] * ipi;                              // v^10/v^5
    q[5] = tid[1] * ap[7] * ipi;
    q[6] = tid[2] * ap[8] * ipi;
    q[7] = tid[3] * ap[9] * ipi;
    q[8] = tid[4] * ap[10] * ipi;
    q[9] = pv[4] / ap[1] * ipi;                               // v^-1
    q[10] = pv[3] / ap[2] * ipi;                              // v^-2
    q[11] = pv[2] / ap[3] * ipi;                              // v^-3
    q[12] = pv[0] / ap[5] * ipi;                              // v^-5   (pv[1] = 0)
    q[13] = pvl[5] / 3.0 * ipi;
    q[14] = pvl[6] * ap[1] / 3.0 * ipi;
    // amplitude: amp0 sqrt(5/(32 eta)) v^(-7/2) = A0 f^(-7/6)
    const double amp0 = -4.0 * m1 * m2 / (dist_mpc * 1e6 * BB_PARSEC_SI) * MRSUN * MTSUN * sqrt(pi / 12.0);
    coef[BC_A0] = amp0 * sqrt(5.0 / (32.0 * eta)) * pow(piM, -7.0 / 6.0);
    // active bins: i in [ceil(f_min/df), floor(f_end/df)] (upstream n = f_max/df + 1), then the reference's
    // frequency_bounds and the detector masks
    const double df = net.df;
    double k0 = ceil(wf.f_min / df);
    double k1 = floor(f_end / df + 1.0);
    const double kb = floor(wf.f_max / df) + 1.0;
    if (wf.f_max > 0.0 && k1 > kb) k1 = kb;
    if (k0 < (double)net.k_lo) k0 = (double)net.k_lo;
    if (k1 > (double)(net.k_hi + 1)) k1 = (double)(net.k_hi + 1);
    if (k1 > (double)net.n_freq) k1 = (double)net.n_freq;
    if (k1 < k0) k1 = k0;
    coef[BC_KMIN] = k0;
    coef[BC_KMAX] = k1;
}

BB_HD double bb_taylorf2_amp(const double* c, double u, double t) {
    return c[BC_A0] * (u * (t * t * t));
}

BB_HD double bb_taylorf2_phase(const double* c, double f, double t, double x, double lf) {
    const double* q = c + BT_P;
    const double x2 = x * x, x5 = x2 * x2 * x;
    // tidal block x^5 (q4 + x^2 (q5 + x (q6 + x (q7 + x q8))))
    double tid = q[8];
    tid = tid * x + q[7]; tid = tid * x + q[6]; tid = tid * x + q[5]; tid = tid * x2 + q[4];
    double neg = q[12] * t * t + q[11];
    neg = neg * t + q[10]; neg = neg * t + q[9];
    return q[0] + x * (q[1] + x * q[2]) + q[3] * f + tid * x5 + neg * t + lf * (q[13] + q[14] * x);
}
