// Common definitions for the bilby_b200 CUDA library (sm_100a).
//
// The per-sample "prologue" math is written as host+device inline functions so that the exact same
// source can be exercised by a host-compiled unit test in the build container (tests/host_check.cpp,
// test infrastructure only - the shipped library exposes no CPU execution path).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define BB_HD __host__ __device__ __forceinline__
#else
#define BB_HD inline
#endif

#include "bb_math.cuh"

#define BB_MAX_DET 4
#define BB_NPARAM 16

// canonical per-sample parameter columns (row-major [n][BB_NPARAM] doubles); must match
// include/bilby_b200.h and bilby_b200/gw/_params.py
enum {
    BB_P_MASS_1 = 0, BB_P_MASS_2, BB_P_CHI_1, BB_P_CHI_2, BB_P_DISTANCE, BB_P_THETA_JN, BB_P_PSI,
    BB_P_PHASE, BB_P_RA, BB_P_DEC, BB_P_GEOCENT_TIME, BB_P_TIME_JITTER, BB_P_LAMBDA_1, BB_P_LAMBDA_2,
    BB_P_SPARE_0, BB_P_SPARE_1
};

// physical constants: the reference mirrors LAL's values (bilby/core/utils/constants.py:3-7)
#define BB_C_SI 299792458.0
#define BB_PARSEC_SI 3.085677581491367e+16
#define BB_MSUN_SI 1.988409870698050731911960804878414216e30
#define BB_G_SI 6.6743e-11
#define BB_PI 3.141592653589793238462643383279502884
#define BB_EULER_GAMMA 0.5772156649015328606065120900824024

// ---------------------------------------------------------------------------------------------
// per-sample coefficient record (doubles).  Written by the prologue kernel, read by the
// inner-product kernels.  All frequency dependence is expressed in Hz with the total-mass scaling
// folded into the coefficients; all phase coefficients are in units of pi ("half turns") so that
// the per-bin complex exponential is one sincospi().
// ---------------------------------------------------------------------------------------------
enum {
    BC_A0 = 0,        // amplitude prefactor multiplying f^(-7/6)
    BC_FA1, BC_FA2,   // amplitude region boundaries [Hz]
    BC_AINS,          // 10 Horner coefficients in x = f^(1/3)
    BC_AINT = BC_AINS + 10,   // 5 coefficients in xs = (f - f1) * invw
    BC_AINT_F1 = BC_AINT + 5,
    BC_AINT_INVW,
    BC_MR_FRD, BC_MR_WL2, BC_MR_G, BC_MR_LAM,
    BC_FP1, BC_FP2,   // phase region boundaries [Hz]
    BC_PINS,          // 13: 1, x, x^2, f, f x, f x^2, f^2, 1/x, 1/x^2, 1/f, f^(-5/3), ln f, x ln f
    BC_PINT = BC_PINS + 13,   // 4: 1, f, f^-3, ln f
    BC_PMR = BC_PINT + 4,     // 7: 1, f, 1/f, f^(3/4), atan coefficient, f0, 1/fdamp
    BC_KMIN = BC_PMR + 7,     // active bins [kmin, kmax) (stored as exact doubles)
    BC_KMAX,
    BC_DISTANCE,      // luminosity distance [Mpc] (distance marginalisation rescaling)
    BC_STATUS,        // 0 = ok, 1 = waveform domain error (-> likelihood sentinel)
    BC_JITTER,
    BC_DT0,           // t_c - t_start [s] (reduced-order kernels add it to the per-detector delay themselves)
    BC_KA1, BC_KA2,   // first bin of the amplitude intermediate / merger-ringdown regions (exact doubles)
    BC_KP1, BC_KP2,   // first bin of the phase intermediate / merger-ringdown regions
    BC_DET,           // per detector BC_DSTRIDE: K_re, K_im, 2*dt_det [half turns / Hz], |K|^2,
                      // Re/Im of exp(+i pi * 2 dt_det * 32 df) (phase-ramp step over one row of 32 bins)
    BC_DSTRIDE = 6,
    BC_NCOEF = BC_DET + BC_DSTRIDE * BB_MAX_DET
};
#define BB_ROW 32     // bins per row: lane l of a warp owns bins k = 32 r + l

struct BBNetwork {
    int n_det;
    int n_freq;           // bins of the full one-sided grid, N = round(T fs / 2) + 1
    int k_lo, k_hi;       // bounding bin range of the union of detector masks, inclusive
    double duration, sampling_frequency, start_time, df;
    double detector_tensor[BB_MAX_DET][9];
    double vertex[BB_MAX_DET][3];
};

struct BBWaveformConfig {
    int approximant;      // 0 IMRPhenomD, 1 TaylorF2(+tides)
    int add_jitter;       // time marginalisation with jitter: geocent_time += time_jitter (base.py:427-428)
    double f_ref, f_min, f_max;
    // frequency-sequence semantics (source.py:1068-1140, ROQ nodes / relative-binning edges): every node is
    // evaluated, f_min = first node for the IMRPhenomD domain check, no f_end check for TaylorF2
    int sequence;
    int no_time_shift;    // ROQ: the waveform carries no exp(-2 pi i f (t_c - t_start)) (roq.py:486-502)
    int fixed_antenna_time;   // 1: ROQ time marginalisation: antenna response AND delay at antenna_time (roq.py:478-481);
                              // 2: multi-banded time marginalisation: antenna response only (Interferometer.reference_time)
    double antenna_time;
};

// cubic-spline calibration grid (bilby/gw/detector/calibration.py:257-384): per detector the spline nodes are
// linspace(log10 fmin, log10 fmax, n_points)
#define BB_NCAL_MAX 32
struct BBCalGrid {
    int n_points;                    // 0 = no calibration model
    double l0[BB_MAX_DET];           // log10 of the first node
    double inv_delta[BB_MAX_DET];    // 1 / node spacing in log10 f
    int shared;                      // every detector has the same node grid: bin weights computed once per bin
};

// calibration factor C = (1 + dA) (2 + i dphi) / (2 - i dphi) = amp1 * (cr + i ci), |cr + i ci| = 1.
// rec: [n][4] = per node (amplitude, its spline coefficient, phase, its spline coefficient): the eight values one bin
// needs are two neighbouring nodes = 64 contiguous bytes (node-major since round 2: with the four arrays one after the
// other every load needed its own address, 43 IMADs per row of 32 bins in K4a)
// spline bin weights: depend only on the frequency and the node grid (calibration.py:368-376)
struct BBCalW {
    int j;
    double a, b, c, d;
};
BB_HD BBCalW bb_cal_weights(int n, double l0, double inv_delta, double lf) {
    BBCalW w;
    const double x = (lf * 0.43429448190325182765 - l0) * inv_delta;       // log10 f = ln f / ln 10
    int j = (int)x;                                                         // astype(int): truncation
    j = j < 0 ? 0 : (j > n - 2 ? n - 2 : j);
    w.j = j;
    w.b = x - (double)j;
    w.a = 1.0 - w.b;
    w.c = (w.a * w.a * w.a - w.a) * (1.0 / 6.0);
    w.d = (w.b * w.b * w.b - w.b) * (1.0 / 6.0);
    return w;
}
BB_HD void bb_cal_finish(double dA, double dP, double* amp1, double* cr, double* ci) {
    const double den = bb_rcp_pos(4.0 + dP * dP);
    *amp1 = 1.0 + dA;
    *cr = (4.0 - dP * dP) * den;
    *ci = 4.0 * dP * den;
}
BB_HD void bb_cal_apply(const double* rec, int n, const BBCalW& w, double* amp1, double* cr, double* ci) {
    const double* r = rec + 4 * w.j;
    const double dA = w.a * r[0] + w.b * r[4] + w.c * r[1] + w.d * r[5];
    const double dP = w.a * r[2] + w.b * r[6] + w.c * r[3] + w.d * r[7];
    bb_cal_finish(dA, dP, amp1, cr, ci);
}
#ifdef __CUDACC__
// the same with four 16-byte loads; rec must be 16-byte aligned (the shared-memory records of K1 and K4a are)
__device__ __forceinline__ void bb_cal_apply_v(const double* rec, int n, const BBCalW& w, double* amp1, double* cr,
                                               double* ci) {
    const double2* r = reinterpret_cast<const double2*>(rec + 4 * w.j);
    const double2 a0 = r[0], p0 = r[1], a1 = r[2], p1 = r[3];
    const double dA = w.a * a0.x + w.b * a1.x + w.c * a0.y + w.d * a1.y;
    const double dP = w.a * p0.x + w.b * p1.x + w.c * p0.y + w.d * p1.y;
    bb_cal_finish(dA, dP, amp1, cr, ci);
}
#endif
BB_HD void bb_cal_factor(const double* rec, int n, double l0, double inv_delta, double lf, double* amp1,
                         double* cr, double* ci) {
    const BBCalW w = bb_cal_weights(n, l0, inv_delta, lf);
    bb_cal_apply(rec, n, w, amp1, cr, ci);
}

struct BBQnmTable {
    const double* x;
    const double* fring;
    const double* fring_d2;
    const double* fdamp;
    const double* fdamp_d2;
    int n;
};
