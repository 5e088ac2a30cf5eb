// Special functions for the marginalisation epilogue.
//   bb_ln_i0        <- bilby/gw/utils.py:1006-1022  (log(i0e(x)) + |x|)
//   bb_bispev       <- bilby/core/utils/calculus.py:221-262 BoundedRectBivariateSpline, i.e. the
//                      FITPACK tensor-product cubic B-spline scipy's RectBivariateSpline evaluates,
//                      with the reference's out-of-bounds fill value (-inf)
//   bb_logsumexp_*  <- scipy.special.logsumexp(a, b=...) as used in base.py:794-820, 994-1018
#pragma once
#include "bb_common.cuh"

#ifdef __CUDACC__
#define BB_CONST_QUAL __device__ const
#else
#define BB_CONST_QUAL static const
#endif
#include "i0e_coeffs.inc"

// Clenshaw evaluation of sum_k c_k T_k(z)
BB_HD double bb_cheb(const double* c, int n, double z) {
    double b0 = 0.0, b1 = 0.0, b2 = 0.0;
    const double z2 = 2.0 * z;
    for (int k = n - 1; k >= 1; --k) {
        b2 = b1;
        b1 = b0;
        b0 = z2 * b1 - b2 + c[k];
    }
    return z * b0 - b1 + c[0];
}

// exp(-|x|) I0(|x|)
BB_HD double bb_i0e(double x, const double* ca, const double* cb) {
    x = fabs(x);
    if (x <= 8.0) return bb_cheb(ca, BB_I0E_A_N, 0.25 * x - 1.0);
    return bb_cheb(cb, BB_I0E_B_N, 16.0 / x - 1.0) / sqrt(x);
}

BB_HD double bb_ln_i0(double x, const double* ca, const double* cb) {
    return log(bb_i0e(x, ca, cb)) + fabs(x);
}

// cubic B-spline basis functions at x for knot interval l (t[l] <= x < t[l+1]); de Boor recurrence
// (FITPACK fpbspl, k = 3).  h[0..3] multiply coefficients l-3 .. l.
BB_HD void bb_bspl3(const double* t, int l, double x, double* h) {
    double hh[3];
    h[0] = 1.0;
    for (int j = 1; j <= 3; ++j) {
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 0; i < j; ++i) {
            const int li = l + i + 1, lj = li - j;
            const double f = hh[i] / (t[li] - t[lj]);
            h[i] += f * (t[li] - x);
            h[i + 1] = f * (x - t[lj]);
        }
    }
}

struct BBSpline2D {
    const double* tx;   // nx knots
    const double* ty;   // ny knots
    const double* c;    // (nx-4) * (ny-4) coefficients, row-major in x
    int nx, ny;
    double xmin, xmax, ymin, ymax;   // reference's bounding box (min/max of the grids)
};

BB_HD int bb_find_interval(const double* t, int n, double x) {
    // largest l in [3, n-5] with t[l] <= x  (FITPACK fpbisp search)
    int lo = 3, hi = n - 4;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= t[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

BB_HD double bb_bispev(const BBSpline2D& s, double x, double y) {
    if (!(x >= s.xmin) || !(x <= s.xmax) || !(y >= s.ymin) || !(y <= s.ymax)) return -INFINITY;
    const int lx = bb_find_interval(s.tx, s.nx, x);
    const int ly = bb_find_interval(s.ty, s.ny, y);
    double hx[4], hy[4];
    bb_bspl3(s.tx, lx, x, hx);
    bb_bspl3(s.ty, ly, y, hy);
    const int ncy = s.ny - 4;
    double out = 0.0;
    for (int i = 0; i < 4; ++i) {
        const double* row = s.c + (size_t)(lx - 3 + i) * ncy + (ly - 3);
        out += hx[i] * (hy[0] * row[0] + hy[1] * row[1] + hy[2] * row[2] + hy[3] * row[3]);
    }
    return out;
}
