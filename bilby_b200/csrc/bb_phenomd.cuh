// IMRPhenomD per-sample prologue: physical parameters -> coefficient record (bb_common.cuh).
//
// Replaces, on the device, the arithmetic the reference delegates to lalsimulation through
// bilby/gw/source.py:597-643 / bilby/gw/utils.py:642-684 (SimInspiralChooseFDWaveform, IMRPhenomD):
// final state and QNM frequencies, the 19 phenomenological fits (Khan+ 2016 Table V), TaylorF2
// 3.5PN aligned-spin phasing, the C1 connection coefficients and the time/phase alignment.
// Structure is this library's own: everything is reduced to polynomial coefficients in f [Hz],
// f^(1/3), ln f, f^(3/4) (tabulated once per frequency grid), with phases in half turns.
#pragma once
#include "bb_common.cuh"
#include "bb_geometry.cuh"

#define BB_NFIT 19
// rows: rho1 rho2 rho3 v2 gamma1 gamma2 gamma3 sigma1..4 beta1..3 alpha1..5 ; 11 numbers per row:
// c00 c01 | c10 c11 c12 | c20 c21 c22 | c30 c31 c32   (powers of eta within, of xi = chiPN-1 across)
enum { F_RHO1 = 0, F_RHO2, F_RHO3, F_V2, F_GAMMA1, F_GAMMA2, F_GAMMA3, F_SIGMA1, F_SIGMA2, F_SIGMA3,
       F_SIGMA4, F_BETA1, F_BETA2, F_BETA3, F_ALPHA1, F_ALPHA2, F_ALPHA3, F_ALPHA4, F_ALPHA5 };

BB_HD double bb_fit_row(const double* c, double eta, double xi) {
    const double eta2 = eta * eta;
    return c[0] + c[1] * eta + (c[2] + c[3] * eta + c[4] * eta2) * xi
           + (c[5] + c[6] * eta + c[7] * eta2) * xi * xi
           + (c[8] + c[9] * eta + c[10] * eta2) * xi * xi * xi;
}

BB_HD double bb_final_spin(double eta, double chi1, double chi2) {
    const double seta = sqrt(1.0 - 4.0 * eta);
    const double m1 = 0.5 * (1.0 + seta), m2 = 0.5 * (1.0 - seta);
    const double s = m1 * m1 * chi1 + m2 * m2 * chi2;
    const double eta2 = eta * eta, eta3 = eta2 * eta, eta4 = eta3 * eta;
    const double s2 = s * s, s3 = s2 * s, s4 = s3 * s;
    return 3.4641016151377544 * eta - 4.399247300629289 * eta2 + 9.397292189321194 * eta3
           - 13.180949901606242 * eta4
           + (1 - 0.0850917821418767 * eta - 5.837029316602263 * eta2) * s
           + (0.1014665242971878 * eta - 2.0967746996832157 * eta2) * s2
           + (-1.3546806617824356 * eta + 4.108962025369336 * eta2) * s3
           + (-0.8676969352555539 * eta + 2.064046835273906 * eta2) * s4;
}

BB_HD double bb_e_rad(double eta, double chi1, double chi2) {
    const double seta = sqrt(1.0 - 4.0 * eta);
    const double m1 = 0.5 * (1.0 + seta), m2 = 0.5 * (1.0 - seta);
    const double s = (m1 * m1 * chi1 + m2 * m2 * chi2) / (m1 * m1 + m2 * m2);
    const double eta2 = eta * eta, eta3 = eta2 * eta, eta4 = eta3 * eta;
    return ((0.055974469826360077 * eta + 0.5809510763115132 * eta2 - 0.9606726679372312 * eta3
             + 3.352411249771192 * eta4)
            * (1. + (-0.0030302335878845507 - 2.0066110851351073 * eta + 7.7050567802399215 * eta2) * s))
           / (1. + (-0.6714403054720589 - 1.4756929437702908 * eta + 7.304676214885011 * eta2) * s);
}

// natural cubic spline from nodes, values and second derivatives
BB_HD double bb_spline_eval(const double* x, const double* y, const double* y2, int n, double xq) {
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x[mid] > xq) hi = mid; else lo = mid;
    }
    const double h = x[hi] - x[lo];
    const double a = (x[hi] - xq) / h;
    const double b = (xq - x[lo]) / h;
    return a * y[lo] + b * y[hi] + ((a * a * a - a) * y2[lo] + (b * b * b - b) * y2[hi]) * (h * h) / 6.0;
}

// TaylorF2 3.5PN aligned-spin phasing coefficients, v[k], vlogv[k] (k = 0..7), times 3/(128 eta)
BB_HD void bb_taylorf2_phasing(double m1, double m2, double chi1, double chi2, double qm1, double qm2,
                               double* v, double* vl) {
    const double M = m1 + m2, m1M = m1 / M, m2M = m2 / M;
    const double eta = m1 * m2 / (M * M);
    const double d = (m1 - m2) / M;
    const double pi = BB_PI;
    const double pfaN = 3.0 / (128.0 * eta);
    for (int i = 0; i < 8; ++i) { v[i] = 0.0; vl[i] = 0.0; }
    v[0] = 1.0;
    v[2] = 5.0 * (743.0 / 84.0 + 11.0 * eta) / 9.0;
    v[3] = -16.0 * pi;
    v[4] = 5.0 * (3058.673 / 7.056 + 5429.0 / 7.0 * eta + 617.0 * eta * eta) / 72.0;
    v[5] = 5.0 / 9.0 * (7729.0 / 84.0 - 13.0 * eta) * pi;
    vl[5] = 5.0 / 3.0 * (7729.0 / 84.0 - 13.0 * eta) * pi;
    v[6] = (11583.231236531 / 4.694215680 - 640.0 / 3.0 * pi * pi - 6848.0 / 21.0 * BB_EULER_GAMMA)
           + eta * (-15737.765635 / 3.048192 + 2255. / 12. * pi * pi)
           + eta * eta * 76055.0 / 1728.0 - eta * eta * eta * 127825.0 / 1296.0;
    v[6] += (-6848.0 / 21.0) * 1.3862943611198906188;   // log(4)
    vl[6] = -6848.0 / 21.0;
    v[7] = pi * (77096675. / 254016. + 378515. / 1512. * eta - 74045. / 756. * eta * eta);

    const double chi1sq = chi1 * chi1, chi2sq = chi2 * chi2;
    const double SL = m1M * m1M * chi1 + m2M * m2M * chi2;
    const double dSigmaL = d * (m2M * chi2 - m1M * chi1);
    double pn_sigma = eta * (721. / 48. * chi1 * chi2 - 247. / 48. * chi1 * chi2);
    pn_sigma += (720. * qm1 - 1.) / 96.0 * m1M * m1M * chi1sq;
    pn_sigma += (720. * qm2 - 1.) / 96.0 * m2M * m2M * chi2sq;
    pn_sigma -= (240. * qm1 - 7.) / 96.0 * m1M * m1M * chi1sq;
    pn_sigma -= (240. * qm2 - 7.) / 96.0 * m2M * m2M * chi2sq;
    double pn_ss3 = (326.75 / 1.12 + 557.5 / 1.8 * eta) * eta * chi1 * chi2;
    pn_ss3 += ((4703.5 / 8.4 + 2935. / 6. * m1M - 120. * m1M * m1M) * qm1
               + (-4108.25 / 6.72 - 108.5 / 1.2 * m1M + 125.5 / 3.6 * m1M * m1M)) * m1M * m1M * chi1sq;
    pn_ss3 += ((4703.5 / 8.4 + 2935. / 6. * m2M - 120. * m2M * m2M) * qm2
               + (-4108.25 / 6.72 - 108.5 / 1.2 * m2M + 125.5 / 3.6 * m2M * m2M)) * m2M * m2M * chi2sq;
    const double pn_gamma = (554345. / 1134. + 110. * eta / 9.) * SL + (13915. / 84. - 10. * eta / 3.) * dSigmaL;
    v[7] += (-8980424995. / 762048. + 6586595. * eta / 756. - 305. * eta * eta / 36.) * SL
            - (170978035. / 48384. - 2876425. * eta / 672. - 4735. * eta * eta / 144.) * dSigmaL;
    v[6] += pi * (3760. * SL + 1490. * dSigmaL) / 3. + pn_ss3;
    v[5] += -1. * pn_gamma;
    vl[5] += -3. * pn_gamma;
    v[4] += -10. * pn_sigma;
    v[3] += 188. * SL / 3. + 25. * dSigmaL;
    for (int i = 0; i < 8; ++i) { v[i] *= pfaN; vl[i] *= pfaN; }
}

BB_HD double bb_subtract_3pn_ss(double m1, double m2, double chi1, double chi2) {
    const double M = m1 + m2, m1M = m1 / M, m2M = m2 / M;
    const double eta = m1 * m2 / (M * M);
    double s = (326.75 / 1.12 + 557.5 / 1.8 * eta) * eta * chi1 * chi2;
    s += ((4703.5 / 8.4 + 2935. / 6. * m1M - 120. * m1M * m1M)
          + (-4108.25 / 6.72 - 108.5 / 1.2 * m1M + 125.5 / 3.6 * m1M * m1M)) * m1M * m1M * chi1 * chi1;
    s += ((4703.5 / 8.4 + 2935. / 6. * m2M - 120. * m2M * m2M)
          + (-4108.25 / 6.72 - 108.5 / 1.2 * m2M + 125.5 / 3.6 * m2M * m2M)) * m2M * m2M * chi2 * chi2;
    return s;
}

// ---- helpers used only inside the prologue (functions of Mf, NOT the per-bin hot path)
struct BBPhenomDScratch {
    double eta, etaInv, fRD, fDM;
    double A[10];                       // amplitude inspiral coefficients of Mf^(k/3)
    double gamma1, gamma2, gamma3;
    double P[14];                       // phase inspiral prefactors (see prologue)
    double pv[8], pvl[8];
    double sigma1, sigma2, sigma3, sigma4, beta1, beta2, beta3, alpha1, alpha2, alpha3, alpha4, alpha5;
};

BB_HD double bb_pd_amp_ins(const BBPhenomDScratch& s, double f) {
    const double x = cbrt(f);
    double out = 0.0;
    for (int k = 9; k >= 0; --k) out = out * x + s.A[k];
    return out;
}
BB_HD double bb_pd_damp_ins(const BBPhenomDScratch& s, double f) {
    const double x = cbrt(f);
    double out = 0.0;
    for (int k = 9; k >= 1; --k) out = out * x + s.A[k] * (k / 3.0);
    return out / (x * x);
}
BB_HD double bb_pd_amp_mrd(const BBPhenomDScratch& s, double f) {
    const double w = s.fDM * s.gamma3, d = f - s.fRD;
    return exp(-d * s.gamma2 / w) * (w * s.gamma1) / (d * d + w * w);
}
BB_HD double bb_pd_damp_mrd(const BBPhenomDScratch& s, double f) {
    const double w = s.fDM * s.gamma3, d = f - s.fRD;
    const double den = d * d + w * w;
    return exp(-d * s.gamma2 / w) * w * s.gamma1 * (-s.gamma2 / w / den - 2.0 * d / (den * den));
}
// P: 0 initial_phasing, 1 two_thirds, 2 third, 3 third_with_logv, 4 logv, 5 minus_third,
//    6 minus_two_thirds, 7 minus_one, 8 minus_five_thirds, 9 one, 10 four_thirds, 11 five_thirds, 12 two
BB_HD double bb_pd_phi_ins(const BBPhenomDScratch& s, double f) {
    const double x = cbrt(f);
    const double logv = log(x * cbrt(BB_PI));
    double ph = s.P[0] + s.P[1] * x * x + s.P[2] * x + s.P[3] * logv * x + s.P[4] * logv
                + s.P[5] / x + s.P[6] / (x * x) + s.P[7] / f + s.P[8] / (f * x * x);
    ph += (s.P[9] * f + s.P[10] * f * x + s.P[11] * f * x * x + s.P[12] * f * f) * s.etaInv;
    return ph;
}
BB_HD double bb_pd_dphi_ins(const BBPhenomDScratch& s, double f) {
    const double pi = BB_PI;
    const double v = cbrt(pi * f), logv = log(v);
    const double v2 = v * v, v3 = v2 * v, v4 = v3 * v, v5 = v4 * v, v6 = v5 * v, v7 = v6 * v, v8 = v7 * v;
    double d = 2.0 * s.pv[7] * v7;
    d += (s.pv[6] + s.pvl[6] * (1.0 + logv)) * v6;
    d += s.pvl[5] * v5;
    d += -1.0 * s.pv[4] * v4;
    d += -2.0 * s.pv[3] * v3;
    d += -3.0 * s.pv[2] * v2;
    d += -4.0 * s.pv[1] * v;
    d += -5.0 * s.pv[0];
    d /= v8 * 3.0 / pi;
    const double x = cbrt(f);
    d += (s.sigma1 + s.sigma2 * x + s.sigma3 * x * x + s.sigma4 * f) * s.etaInv;
    return d;
}
BB_HD double bb_pd_phi_int_beta(const BBPhenomDScratch& s, double f) {
    return (s.beta1 * f - s.beta3 / (3.0 * f * f * f) + s.beta2 * log(f)) * s.etaInv;
}
BB_HD double bb_pd_dphi_int(const BBPhenomDScratch& s, double f) {
    return (s.beta1 + s.beta3 / (f * f * f * f) + s.beta2 / f) * s.etaInv;
}
BB_HD double bb_pd_phi_mrd_alpha(const BBPhenomDScratch& s, double f) {
    const double f34 = sqrt(f * sqrt(f));
    return (-(s.alpha2 / f) + (4.0 / 3.0) * (s.alpha3 * f34) + s.alpha1 * f
            + s.alpha4 * atan((f - s.alpha5 * s.fRD) / s.fDM)) * s.etaInv;
}
BB_HD double bb_pd_dphi_mrd(const BBPhenomDScratch& s, double f) {
    const double x = (f - s.alpha5 * s.fRD) / s.fDM;
    return (s.alpha1 + s.alpha2 / (f * f) + s.alpha3 / sqrt(sqrt(f)) + s.alpha4 / (s.fDM * (1.0 + x * x)))
           * s.etaInv;
}

// 5x5 dense solve with partial pivoting (amplitude collocation, well conditioned in the scaled variable)
BB_HD void bb_solve5(double a[5][5], double* b) {
    for (int c = 0; c < 5; ++c) {
        int p = c;
        double best = fabs(a[c][c]);
        for (int r = c + 1; r < 5; ++r) if (fabs(a[r][c]) > best) { best = fabs(a[r][c]); p = r; }
        if (p != c) {
            for (int k = 0; k < 5; ++k) { const double t = a[c][k]; a[c][k] = a[p][k]; a[p][k] = t; }
            const double t = b[c]; b[c] = b[p]; b[p] = t;
        }
        const double inv = 1.0 / a[c][c];
        for (int r = c + 1; r < 5; ++r) {
            const double m = a[r][c] * inv;
            for (int k = c; k < 5; ++k) a[r][k] -= m * a[c][k];
            b[r] -= m * b[c];
        }
    }
    for (int r = 4; r >= 0; --r) {
        double acc = b[r];
        for (int k = r + 1; k < 5; ++k) acc -= a[r][k] * b[k];
        b[r] = acc / a[r][r];
    }
}

// sky / detector part of the record, shared by all approximants.
// interferometer.py:303-368: antenna response at geocent_time, dt = (t_c - t_start) + delay.
// Returns dt0 = t_c - t_start (folded into the waveform phase by the caller).
BB_HD double bb_detector_prologue(const double* p, const BBNetwork& net, const BBWaveformConfig& wf, double* coef) {
    const double tc = wf.add_jitter ? p[BB_P_GEOCENT_TIME] + p[BB_P_TIME_JITTER] : p[BB_P_GEOCENT_TIME];
    const double gmst = bb_wrap_2pi(bb_gmst(wf.fixed_antenna_time ? wf.antenna_time : tc));
    // fixed_antenna_time == 2 (multi-banded time marginalisation, multiband.py:714-726 + interferometer.py:336-355):
    // the antenna response at the reference time, the delay still at geocent_time
    const double gmst_delay = (wf.fixed_antenna_time == 2) ? bb_wrap_2pi(bb_gmst(tc)) : gmst;
    const double cfac = cos(p[BB_P_THETA_JN]);
    const double pfac = 0.5 * (1.0 + cfac * cfac);
    // sines and cosines of the sky position once per sample, not once per detector (same values: same arguments)
    const BBSkyTrig sky = bb_sky_trig(p[BB_P_RA], p[BB_P_DEC], p[BB_P_PSI], gmst);
    double cph_delay = sky.cph, sph_delay = sky.sph;
    if (wf.fixed_antenna_time == 2) {
        const double phi_delay = p[BB_P_RA] - gmst_delay;
        cph_delay = cos(phi_delay);
        sph_delay = sin(phi_delay);
    }
    for (int d = 0; d < BB_MAX_DET; ++d) {
        double* cd = coef + BC_DET + BC_DSTRIDE * d;
        if (d < net.n_det) {
            double fp, fc;
            bb_antenna_trig(net.detector_tensor[d], sky, &fp, &fc);
            const double delay = bb_time_delay_trig(net.vertex[d], sky.sth, cph_delay, sph_delay, sky.cth);
            // h_det = F+ h+ + Fx hx with h+ = pfac h22, hx = -i cfac h22
            cd[0] = fp * pfac;
            cd[1] = -fc * cfac;
            cd[2] = 2.0 * delay;
            cd[3] = cd[0] * cd[0] + cd[1] * cd[1];
            // ramp step over one row: exp(+i pi * cd[2] * 32 df); sincospi is device-only, and pi*x with
            // |x| < 1 keeps sin/cos argument error at the 1e-17 level
            const double a = cd[2] * (double)BB_ROW * net.df;
            const double r = a - 2.0 * floor(0.5 * a + 0.5);     // a mod 2 in [-1, 1)
            cd[4] = cos(BB_PI * r);
            cd[5] = sin(BB_PI * r);
        } else {
            for (int i = 0; i < BC_DSTRIDE; ++i) cd[i] = 0.0;
        }
    }
    coef[BC_DT0] = tc - net.start_time;
    return wf.no_time_shift ? 0.0 : tc - net.start_time;
}

// dt0 alone (the value bb_detector_prologue returns), for the thread that builds the waveform part of a record while its
// partner builds the detector part (K0)
BB_HD double bb_prologue_dt0(const double* p, const BBNetwork& net, const BBWaveformConfig& wf) {
    const double tc = wf.add_jitter ? p[BB_P_GEOCENT_TIME] + p[BB_P_TIME_JITTER] : p[BB_P_GEOCENT_TIME];
    return wf.no_time_shift ? 0.0 : tc - net.start_time;
}

// active bin range shared by the approximants: upstream fills i in [int(f_min/df), int(f_max'/df)),
// the reference then zeroes f < minimum_frequency or f > maximum_frequency (source.py:618-619, 678-679)
BB_HD void bb_bin_range(const BBNetwork& net, const BBWaveformConfig& wf, double f_max_prime, double* coef) {
    const double df = net.df;
    double k0 = floor(wf.f_min / df);
    if (k0 * df < wf.f_min) k0 += 1.0;
    double k1 = floor(f_max_prime / df);                 // exclusive
    const double kb = floor(wf.f_max / df) + 1.0;       // bins with f <= maximum_frequency
    if (k1 > kb) k1 = kb;
    if (k0 < (double)net.k_lo) k0 = (double)net.k_lo;
    if (k1 > (double)(net.k_hi + 1)) k1 = (double)(net.k_hi + 1);
    if (k1 > (double)net.n_freq) k1 = (double)net.n_freq;
    if (k1 < k0) k1 = k0;
    coef[BC_KMIN] = k0;
    coef[BC_KMAX] = k1;
}

// WHOLE = false: the waveform part only, into a record that is already zero and whose detector part
// (bb_detector_prologue) another thread writes
template <bool WHOLE = true>
BB_HD void bb_phenomd_prologue(const double* p, const BBNetwork& net, const BBWaveformConfig& wf,
                               const BBQnmTable& qnm, const double* fit /* [19][11] */, double* coef) {
    if (WHOLE) for (int i = 0; i < BC_NCOEF; ++i) coef[i] = 0.0;
    double m1 = p[BB_P_MASS_1], m2 = p[BB_P_MASS_2], chi1 = p[BB_P_CHI_1], chi2 = p[BB_P_CHI_2];
    if (m2 > m1) { double t = m1; m1 = m2; m2 = t; t = chi1; chi1 = chi2; chi2 = t; }
    const double dist_mpc = p[BB_P_DISTANCE];
    coef[BC_DISTANCE] = dist_mpc;
    coef[BC_JITTER] = p[BB_P_TIME_JITTER];
    const double dt0 = WHOLE ? bb_detector_prologue(p, net, wf, coef) : bb_prologue_dt0(p, net, wf);

    const double M = m1 + m2;
    const double MTSUN = BB_G_SI * BB_MSUN_SI / (BB_C_SI * BB_C_SI * BB_C_SI);
    const double MRSUN = BB_G_SI * BB_MSUN_SI / (BB_C_SI * BB_C_SI);
    const double Ms = M * MTSUN;
    const double f_cut = 0.2 / Ms;
    const double f_ref = (wf.f_ref == 0.0) ? wf.f_min : wf.f_ref;
    const double f_max_prime = (wf.f_max == 0.0 || wf.sequence) ? f_cut : (wf.f_max < f_cut ? wf.f_max : f_cut);
    const bool bad = !(m1 > 0.0) || !(m2 > 0.0) || !(dist_mpc > 0.0) || fabs(chi1) > 1.0 || fabs(chi2) > 1.0
                     || !(f_max_prime > wf.f_min) || !isfinite(M) || !isfinite(dist_mpc);
    if (bad) {
        coef[BC_STATUS] = 1.0;
        coef[BC_KMIN] = 0.0;
        coef[BC_KMAX] = 0.0;
        return;
    }
    BBPhenomDScratch s;
    double eta = m1 * m2 / (M * M);
    if (eta > 0.25) eta = 0.25;
    s.eta = eta;
    s.etaInv = 1.0 / eta;
    const double Seta = sqrt(1.0 - 4.0 * eta);
    const double chi_s = 0.5 * (chi1 + chi2), chi_a = 0.5 * (chi1 - chi2);
    const double chipn = chi_s * (1.0 - eta * 76.0 / 113.0) + Seta * chi_a;
    const double xi = chipn - 1.0;
    double finspin = bb_final_spin(eta, chi1, chi2);
    if (finspin < qnm.x[0]) finspin = qnm.x[0];
    if (finspin > qnm.x[qnm.n - 1]) finspin = qnm.x[qnm.n - 1];
    const double erad = bb_e_rad(eta, chi1, chi2);
    s.fRD = bb_spline_eval(qnm.x, qnm.fring, qnm.fring_d2, qnm.n, finspin) / (1.0 - erad);
    s.fDM = bb_spline_eval(qnm.x, qnm.fdamp, qnm.fdamp_d2, qnm.n, finspin) / (1.0 - erad);

    const double rho1 = bb_fit_row(fit + 11 * F_RHO1, eta, xi);
    const double rho2 = bb_fit_row(fit + 11 * F_RHO2, eta, xi);
    const double rho3 = bb_fit_row(fit + 11 * F_RHO3, eta, xi);
    const double v2 = bb_fit_row(fit + 11 * F_V2, eta, xi);
    s.gamma1 = bb_fit_row(fit + 11 * F_GAMMA1, eta, xi);
    s.gamma2 = bb_fit_row(fit + 11 * F_GAMMA2, eta, xi);
    s.gamma3 = bb_fit_row(fit + 11 * F_GAMMA3, eta, xi);
    s.sigma1 = bb_fit_row(fit + 11 * F_SIGMA1, eta, xi);
    s.sigma2 = bb_fit_row(fit + 11 * F_SIGMA2, eta, xi);
    s.sigma3 = bb_fit_row(fit + 11 * F_SIGMA3, eta, xi);
    s.sigma4 = bb_fit_row(fit + 11 * F_SIGMA4, eta, xi);
    s.beta1 = bb_fit_row(fit + 11 * F_BETA1, eta, xi);
    s.beta2 = bb_fit_row(fit + 11 * F_BETA2, eta, xi);
    s.beta3 = bb_fit_row(fit + 11 * F_BETA3, eta, xi);
    s.alpha1 = bb_fit_row(fit + 11 * F_ALPHA1, eta, xi);
    s.alpha2 = bb_fit_row(fit + 11 * F_ALPHA2, eta, xi);
    s.alpha3 = bb_fit_row(fit + 11 * F_ALPHA3, eta, xi);
    s.alpha4 = bb_fit_row(fit + 11 * F_ALPHA4, eta, xi);
    s.alpha5 = bb_fit_row(fit + 11 * F_ALPHA5, eta, xi);

    // ---- amplitude inspiral series in Mf^(k/3)
    const double pi = BB_PI;
    const double p13 = cbrt(pi), p23 = p13 * p13;
    const double chi12 = chi1 * chi1, chi22 = chi2 * chi2;
    const double eta2 = eta * eta, eta3 = eta2 * eta;
    s.A[0] = 1.0;
    s.A[1] = 0.0;
    s.A[2] = ((-969 + 1804 * eta) * p23) / 672.;
    s.A[3] = ((chi1 * (81 * (1 + Seta) - 44 * eta) + chi2 * (81 - 81 * Seta - 44 * eta)) * pi) / 48.;
    s.A[4] = ((-27312085.0 - 10287648 * chi22 - 10287648 * chi12 * (1 + Seta) + 10287648 * chi22 * Seta
               + 24 * (-1975055 + 857304 * chi12 - 994896 * chi1 * chi2 + 857304 * chi22) * eta
               + 35371056 * eta2) * (p23 * p23)) / 8.128512e6;
    s.A[5] = ((p23 * p23 * p13) * (chi2 * (-285197 * (-1 + Seta) + 4 * (-91902 + 1579 * Seta) * eta - 35632 * eta2)
                                  + chi1 * (285197 * (1 + Seta) - 4 * (91902 + 1579 * Seta) * eta - 35632 * eta2)
                                  + 42840 * (-1.0 + 4 * eta) * pi)) / 32256.;
    s.A[6] = -(pi * pi * (-336 * (-3248849057.0 + 2943675504 * chi12 - 3339284256 * chi1 * chi2
                                  + 2943675504 * chi22) * eta2
                          - 324322727232 * eta3
                          - 7 * (-177520268561 + 107414046432 * chi22 + 107414046432 * chi12 * (1 + Seta)
                                 - 107414046432 * chi22 * Seta
                                 + 11087290368 * (chi1 + chi2 + chi1 * Seta - chi2 * Seta) * pi)
                          + 12 * eta * (-545384828789 - 176491177632 * chi1 * chi2 + 202603761360 * chi22
                                        + 77616 * chi12 * (2610335 + 995766 * Seta)
                                        - 77287373856 * chi22 * Seta
                                        + 5841690624 * (chi1 + chi2) * pi + 21384760320 * pi * pi)))
             / 6.0085960704e10;
    s.A[7] = rho1;
    s.A[8] = rho2;
    s.A[9] = rho3;

    // ---- amplitude peak, intermediate collocation
    const double g2 = s.gamma2, g3 = s.gamma3;
    double fmaxCalc;
    if (!(g2 > 1)) fmaxCalc = fabs(s.fRD + (s.fDM * (-1 + sqrt(1 - g2 * g2)) * g3) / g2);
    else fmaxCalc = fabs(s.fRD + (-s.fDM * g3) / g2);
    const double f1 = 0.014, f3 = fmaxCalc, w = f3 - f1;
    double rhs[5];
    rhs[0] = bb_pd_amp_ins(s, f1);
    rhs[1] = v2;
    rhs[2] = bb_pd_amp_mrd(s, f3);
    rhs[3] = bb_pd_damp_ins(s, f1) * w;
    rhs[4] = bb_pd_damp_mrd(s, f3) * w;
    double mat[5][5] = {{1, 0, 0, 0, 0}, {1, 0.5, 0.25, 0.125, 0.0625}, {1, 1, 1, 1, 1},
                        {0, 1, 0, 0, 0}, {0, 1, 2, 3, 4}};
    bb_solve5(mat, rhs);

    // ---- PN phasing and inspiral prefactors
    bb_taylorf2_phasing(m1, m2, chi1, chi2, 1.0, 1.0, s.pv, s.pvl);
    s.pv[6] -= bb_subtract_3pn_ss(m1, m2, chi1, chi2) * s.pv[0];
    s.P[0] = s.pv[5] - pi / 4.0;
    s.P[1] = s.pv[7] * p23;
    s.P[2] = s.pv[6] * p13;
    s.P[3] = s.pvl[6] * p13;
    s.P[4] = s.pvl[5];
    s.P[5] = s.pv[4] / p13;
    s.P[6] = s.pv[3] / p23;
    s.P[7] = s.pv[2] / pi;
    s.P[8] = s.pv[0] / (p23 * p23 * p13);
    s.P[9] = s.sigma1;
    s.P[10] = s.sigma2 * 0.75;
    s.P[11] = s.sigma3 * 0.6;
    s.P[12] = s.sigma4 * 0.5;

    // ---- C1 connection
    const double fi = 0.018, fm = 0.5 * s.fRD;
    const double C2Int = bb_pd_dphi_ins(s, fi) - bb_pd_dphi_int(s, fi);
    const double C1Int = bb_pd_phi_ins(s, fi) - bb_pd_phi_int_beta(s, fi) - C2Int * fi;
    const double C2MRD = bb_pd_dphi_int(s, fm) + C2Int - bb_pd_dphi_mrd(s, fm);
    const double C1MRD = bb_pd_phi_int_beta(s, fm) + C1Int + C2Int * fm - bb_pd_phi_mrd_alpha(s, fm) - C2MRD * fm;

    // ---- time / phase alignment (upstream IMRPhenomDGenerateFD)
    const double t0 = bb_pd_dphi_mrd(s, fmaxCalc);
    const double MfRef = Ms * f_ref;
    double phifRef;
    if (MfRef < fi) phifRef = bb_pd_phi_ins(s, MfRef);
    else if (MfRef >= fm) phifRef = bb_pd_phi_mrd_alpha(s, MfRef) + C1MRD + C2MRD * MfRef;
    else phifRef = bb_pd_phi_int_beta(s, MfRef) + C1Int + C2Int * MfRef;
    const double phi_precalc = 2.0 * p[BB_P_PHASE] + phifRef;
    // Phi_tot(f) = phi(Ms f) - t0 (Ms f - MfRef) - phi_precalc + 2 pi f dt0
    const double lin = -t0 * Ms + 2.0 * pi * dt0;       // coefficient of f [rad / Hz]
    const double cst = t0 * MfRef - phi_precalc;        // constant [rad]
    const double ipi = 1.0 / pi;

    // ---- scale everything to Hz and store
    const double m3 = cbrt(Ms), m32 = m3 * m3;
    const double amp0 = 2. * sqrt(5. / (64. * pi)) * M * MRSUN * M * MTSUN / (dist_mpc * 1e6 * BB_PARSEC_SI);
    coef[BC_A0] = amp0 * sqrt(2.0 * eta / 3.0) / sqrt(p13) * pow(Ms, -7.0 / 6.0);
    coef[BC_FA1] = f1 / Ms;
    coef[BC_FA2] = fmaxCalc / Ms;
    double mk = 1.0;
    for (int k = 0; k < 10; ++k) { coef[BC_AINS + k] = s.A[k] * mk; mk *= m3; }
    for (int k = 0; k < 5; ++k) coef[BC_AINT + k] = rhs[k];
    coef[BC_AINT_F1] = f1 / Ms;
    coef[BC_AINT_INVW] = Ms / w;
    const double W = s.fDM * g3;
    coef[BC_MR_FRD] = s.fRD / Ms;
    coef[BC_MR_WL2] = (W / Ms) * (W / Ms);
    coef[BC_MR_G] = s.gamma1 * W / (Ms * Ms);
    coef[BC_MR_LAM] = s.gamma2 * Ms / W;
    coef[BC_FP1] = fi / Ms;
    coef[BC_FP2] = fm / Ms;

    const double lv0 = log(p13 * m3);   // logv = lv0 + ln(f)/3
    double* q = coef + BC_PINS;
    q[0] = (s.P[0] + s.P[4] * lv0 + cst) * ipi;
    q[1] = (s.P[2] + s.P[3] * lv0) * m3 * ipi;
    q[2] = s.P[1] * m32 * ipi;
    q[3] = (s.P[9] * s.etaInv * Ms + lin) * ipi;
    q[4] = s.P[10] * s.etaInv * Ms * m3 * ipi;
    q[5] = s.P[11] * s.etaInv * Ms * m32 * ipi;
    q[6] = s.P[12] * s.etaInv * Ms * Ms * ipi;
    q[7] = s.P[5] / m3 * ipi;
    q[8] = s.P[6] / m32 * ipi;
    q[9] = s.P[7] / Ms * ipi;
    q[10] = s.P[8] / (Ms * m32) * ipi;
    q[11] = s.P[4] / 3.0 * ipi;
    q[12] = s.P[3] * m3 / 3.0 * ipi;

    const double lMs = log(Ms);
    q = coef + BC_PINT;
    q[0] = (C1Int + s.beta2 * s.etaInv * lMs + cst) * ipi;
    q[1] = ((s.beta1 * s.etaInv + C2Int) * Ms + lin) * ipi;
    q[2] = -s.beta3 * s.etaInv / (3.0 * Ms * Ms * Ms) * ipi;
    q[3] = s.beta2 * s.etaInv * ipi;

    q = coef + BC_PMR;
    q[0] = (C1MRD + cst) * ipi;
    q[1] = ((s.alpha1 * s.etaInv + C2MRD) * Ms + lin) * ipi;
    q[2] = -s.alpha2 * s.etaInv / Ms * ipi;
    q[3] = (4.0 / 3.0) * s.alpha3 * s.etaInv * sqrt(Ms * sqrt(Ms)) * ipi;
    q[4] = s.alpha4 * s.etaInv * ipi;
    q[5] = s.alpha5 * s.fRD / Ms;
    q[6] = Ms / s.fDM;

    bb_bin_range(net, wf, f_max_prime, coef);
    // first bin of each region, consistent with the (f < boundary) tests of the per-bin functions
    const int slots[4] = {BC_KA1, BC_KA2, BC_KP1, BC_KP2};
    const double bounds[4] = {coef[BC_FA1], coef[BC_FA2], coef[BC_FP1], coef[BC_FP2]};
    for (int i = 0; i < 4; ++i) {
        double k = ceil(bounds[i] / net.df);
        while ((k - 1.0) * net.df >= bounds[i]) k -= 1.0;
        while (k * net.df < bounds[i]) k += 1.0;
        coef[slots[i]] = k;
    }
}

// ---------------------------------------------------------------------------------------------
// per-bin evaluation (the hot path).  Inputs: f [Hz], u = f^(-1/6), lf = ln f, q34 = f^(3/4).
// Returns amplitude A (>= 0) and total phase in half turns.
// ---------------------------------------------------------------------------------------------
// amplitude without the prefactor a0 f^(-7/6)
BB_HD double bb_phenomd_amp_core(const double* c, double f, double x) {
    double a;
    if (f < c[BC_FA1]) {
        const double* k = c + BC_AINS;
        a = k[9];
        a = a * x + k[8]; a = a * x + k[7]; a = a * x + k[6]; a = a * x + k[5]; a = a * x + k[4];
        a = a * x + k[3]; a = a * x + k[2]; a = a * x + k[1]; a = a * x + k[0];
    } else if (f < c[BC_FA2]) {
        const double* k = c + BC_AINT;
        const double xs = (f - c[BC_AINT_F1]) * c[BC_AINT_INVW];
        a = k[4];
        a = a * xs + k[3]; a = a * xs + k[2]; a = a * xs + k[1]; a = a * xs + k[0];
    } else {
        const double d = f - c[BC_MR_FRD];
        a = c[BC_MR_G] * exp(-c[BC_MR_LAM] * d) / (d * d + c[BC_MR_WL2]);
    }
    return a;
}

BB_HD double bb_phenomd_amp(const double* c, double f, double u, double t, double x) {
    const double t3 = t * t * t;
    return bb_phenomd_amp_core(c, f, x) * c[BC_A0] * (u * t3);      // f^(-7/6) = u^7 = u * (u^2)^3
}

BB_HD double bb_phenomd_phase(const double* c, double f, double t, double x, double lf, double q34) {
    if (f < c[BC_FP1]) {
        const double* q = c + BC_PINS;
        double pos = q[6];
        pos = pos * x + q[5]; pos = pos * x + q[4]; pos = pos * x + q[3]; pos = pos * x + q[2];
        pos = pos * x + q[1];
        double neg = q[10] * t * t + q[9];
        neg = neg * t + q[8]; neg = neg * t + q[7];
        return q[0] + pos * x + neg * t + lf * (q[11] + q[12] * x);
    } else if (f < c[BC_FP2]) {
        const double* q = c + BC_PINT;
        const double t3 = t * t * t;
        return q[0] + q[1] * f + q[2] * (t3 * t3 * t3) + q[3] * lf;
    } else {
        const double* q = c + BC_PMR;
        return q[0] + q[1] * f + q[2] * (t * t * t) + q[3] * q34 + q[4] * atan((f - q[5]) * q[6]);
    }
}
