// Device math layer for the per-bin loops (included by bb_common.cuh).
//
// The per-bin work of every evaluation kernel is dominated by sincospi, atan, exp and reciprocals.  The CUDA
// library versions materialise each polynomial coefficient with two UMOV immediates next to the DFMA that uses it
// (three issue slots per term) and carry full-range / IEEE slow paths.  The versions here keep their coefficients in
// __constant__ tables (one LDCU.128 per two coefficients) and are specialised for the argument ranges of this path:
//
//   bb_sincospi(x)   |x| < 2^50 half turns (NaN beyond); argument reduction x = n/2 + r by the 1.5*2^52 shift (two DADDs), minimax
//                    polynomials (7 terms) for sin(pi r), cos(pi r) on |r| <= 1/4 (oracle/tools/make_math_coeffs.py;
//                    errors 2.5e-18 / 4.7e-17 before rounding)
//   bb_atan(y)       three-way reduction at tan(pi/8), tan(3 pi/8), one reciprocal, 11-term minimax polynomial
//   bb_rcp_pos(a)    reciprocal of a normal positive number: rcp.approx.ftz.f64 (MUFU.RCP64H) + two Newton steps
//                    (relative error ~1.5e-16; not correctly rounded, no special cases)
//
// All three agree with the CUDA / libm functions to a few 1e-16 (tests/test_host_math.py, tests/test_gpu_parity.py);
// the likelihood gate is 1e-8 relative.  Host builds (tests/host_check.cpp) run the same arithmetic with the tables
// as ordinary constants.
#pragma once
#include "math_coeffs.inc"

#ifdef __CUDACC__
static __constant__ double bb_kc_sinpi_d[BB_SINPI_N] = {BB_SINPI_COEFFS};
static __constant__ double bb_kc_cospi_d[BB_COSPI_N] = {BB_COSPI_COEFFS};
static __constant__ double bb_kc_atan_d[BB_ATAN_N] = {BB_ATAN_COEFFS};
#endif
static const double bb_kc_sinpi_h[BB_SINPI_N] = {BB_SINPI_COEFFS};
static const double bb_kc_cospi_h[BB_COSPI_N] = {BB_COSPI_COEFFS};
static const double bb_kc_atan_h[BB_ATAN_N] = {BB_ATAN_COEFFS};
#ifdef __CUDA_ARCH__
#define BB_KC(name, i) name##_d[i]
#else
#define BB_KC(name, i) name##_h[i]
#endif

// 1 / a for normal a > 0
BB_HD double bb_rcp_pos(double a) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / a;
#endif
}

// sin(pi x), cos(pi x) for |x| < 2^50 half turns.  Beyond that a double carries no quarter-turn information (and
// x = inf / nan has none either): the result is NaN, formed by one select on the high word instead of a branch to the
// library's full-range path (the branch, its convergence barrier and the inlined slow path cost every per-bin loop
// four issue slots and a good part of its register budget).
BB_HD void bb_sincospi(double x, double* sn, double* cs) {
#ifdef __CUDA_ARCH__
    const bool big = (__double2hiint(x) & 0x7fffffff) >= 0x43100000;       // |x| >= 2^50, inf, nan
    const double shift = 6755399441055744.0;                              // 1.5 * 2^52
    const double t = fma(x, 2.0, shift);                                  // low word = nearest integer n of 2x
    const int n = __double2loint(t);
    double r = fma(t - shift, -0.5, x);                                   // exact, |r| <= 1/4
    r = __hiloint2double(big ? 0x7ff80000 : __double2hiint(r), __double2loint(r));
#else
    if (!(fabs(x) < 1125899906842624.0)) { *sn = *cs = NAN; return; }
    const double n2 = nearbyint(x + x);
    const int n = (int)(unsigned)(unsigned long long)(long long)n2;
    const double r = x - 0.5 * n2;
#endif
    const double u = r * r;
    static_assert(BB_SINPI_N == BB_COSPI_N, "sin / cos polynomials advance together");
    double s = BB_KC(bb_kc_sinpi, BB_SINPI_N - 1), c = BB_KC(bb_kc_cospi, BB_COSPI_N - 1);
#pragma unroll
    for (int i = BB_SINPI_N - 2; i >= 0; --i) {
        s = fma(s, u, BB_KC(bb_kc_sinpi, i));
        c = fma(c, u, BB_KC(bb_kc_cospi, i));
    }
    s *= r;
    // x = n/2 + r: rotate by n quarter turns
    const double a = (n & 1) ? c : s, b = (n & 1) ? s : c;
#ifdef __CUDA_ARCH__
    // sign flips on the high word (two integer instructions instead of a DADD and two selects each)
    *sn = __hiloint2double(__double2hiint(a) ^ ((n & 2) << 30), __double2loint(a));
    *cs = __hiloint2double(__double2hiint(b) ^ (((n + 1) & 2) << 30), __double2loint(b));
#else
    *sn = (n & 2) ? -a : a;
    *cs = ((n + 1) & 2) ? -b : b;
#endif
}

#ifdef __CUDACC__
// The same with a branch to the CUDA library for |x| >= 2^30 (round 1's form; see bb_relbin_edge_sample for its one user)
__device__ __forceinline__ void bb_sincospi_branchy(double x, double* sn, double* cs) {
    if ((__double2hiint(x) & 0x7fffffff) >= 0x41d00000) { sincospi(x, sn, cs); return; }
    bb_sincospi(x, sn, cs);
}
#endif

// atan(y), any finite y
BB_HD double bb_atan(double y) {
    const double ay = fabs(y);
    const double t1 = 0.41421356237309504880, t3 = 2.41421356237309504880;
    // |y| <= tan(pi/8): t = |y|;  <= tan(3 pi/8): t = (|y| - 1) / (|y| + 1), offset pi/4;  else t = -1/|y|, offset pi/2
    const bool mid = ay > t1, big = ay > t3;
    const double num = big ? -1.0 : (mid ? ay - 1.0 : ay);
    const double den = big ? ay : (mid ? ay + 1.0 : 1.0);
    const double t = mid ? num * bb_rcp_pos(den) : ay;
    const double off = big ? 1.57079632679489661923 : (mid ? 0.78539816339744830962 : 0.0);
    const double off_lo = big ? 6.123233995736766e-17 : (mid ? 3.061616997868383e-17 : 0.0);
    const double u = t * t;
    double p = BB_KC(bb_kc_atan, BB_ATAN_N - 1);
#pragma unroll
    for (int i = BB_ATAN_N - 2; i >= 1; --i) p = fma(p, u, BB_KC(bb_kc_atan, i));
    // atan(t) = t + t u p(u) (the leading coefficient is exactly 1)
    const double r = off + (fma(t * u, p, off_lo) + t);          // >= 0 in every branch
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(r) ^ (__double2hiint(y) & 0x80000000), __double2loint(r));   // copysign(r, y)
#else
    return y < 0.0 ? -r : r;
#endif
}
