// Detector geometry / time for the per-sample prologue.
//   greenwich_mean_sidereal_time  <- bilby/gw/time.py:114-165 (+ :57-66 julian_day, :168-192 leap seconds)
//   antenna_response (F+, Fx)     <- bilby/gw/detector/interferometer.py:267-301,
//                                    bilby/gw/geometry.py:118-186 (polarisation tensors), :261-279 (contraction)
//   time_delay_from_geocenter     <- bilby/gw/geometry.py:282-343
#pragma once
#include "bb_common.cuh"

// The GMST chain adds numbers of very different magnitude (5e8 s +- 1e-8): keep the reference's
// operation order and forbid FMA contraction so that host numpy and the device agree bit for bit.
#ifdef __CUDA_ARCH__
#define BB_ADD(a, b) __dadd_rn((a), (b))
#define BB_MUL(a, b) __dmul_rn((a), (b))
#define BB_DIV(a, b) __ddiv_rn((a), (b))
#else
#define BB_ADD(a, b) ((a) + (b))
#define BB_MUL(a, b) ((a) * (b))
#define BB_DIV(a, b) ((a) / (b))
#endif

BB_HD int bb_n_leap_seconds(double gps_int) {
    const double leap[18] = {46828800., 78364801., 109900802., 173059203., 252028804., 315187205.,
                             346723206., 393984007., 425520008., 457056009., 504489610., 551750411.,
                             599184012., 820108813., 914803214., 1025136015., 1119744016., 1167264017.};
    int n = 0;
    for (int i = 0; i < 18; ++i) n += (gps_int > leap[i]) ? 1 : 0;
    return n;
}

// GMST in radians, NOT wrapped (time.py:151-165 with equation_of_equinoxes = 0)
BB_HD double bb_gmst(double gps_time) {
    const double gps_int = floor(gps_time);
    const double frac = BB_ADD(gps_time, -gps_int);             // gps_time % 1 (exact)
    const double second = BB_ADD(gps_int, -(double)bb_n_leap_seconds(gps_int));
    // datetime(1980, 1, 6, second=...).julian_day: 367*1980 - 7*(1980+(1+9)//12)//4 + 275*1//9 + 6 = 723231
    double jd = BB_ADD(723231.0, BB_DIV(second, 86400.0));
    jd = BB_ADD(jd, 1721013.5);
    const double t_hi = BB_DIV(BB_ADD(jd, -2451545.0), 36525.0);
    const double t_lo = BB_DIV(frac, BB_MUL(36525.0, 86400.0));
    const double t = BB_ADD(t_hi, t_lo);
    // equation_of_equinoxes + (-6.2e-6 * t + 0.093104) * t**2 + 67310.54841
    double st = BB_MUL(BB_ADD(BB_MUL(-6.2e-6, t), 0.093104), BB_MUL(t, t));
    st = BB_ADD(BB_ADD(0.0, st), 67310.54841);
    st = BB_ADD(st, BB_MUL(8640184.812866, t_lo));
    st = BB_ADD(st, BB_MUL(3155760000.0, t_lo));
    st = BB_ADD(st, BB_MUL(8640184.812866, t_hi));
    st = BB_ADD(st, BB_MUL(3155760000.0, t_hi));
    // sidereal_time * 2 * np.pi / SECONDS_PER_DAY  (left to right)
    return BB_DIV(BB_MUL(BB_MUL(st, 2.0), BB_PI), 86400.0);
}

BB_HD double bb_wrap_2pi(double x) {
    // python's float % (2 pi): result has the sign of the divisor
    const double twopi = 2.0 * BB_PI;
    double r = fmod(x, twopi);
    if (r < 0.0) r += twopi;
    return r;
}

// F+, Fx for one detector tensor d[9] (row-major) given (ra, dec, psi) and the wrapped GMST
// The six sines and cosines of the sky position and polarisation angle do not depend on the detector: K0 evaluates
// them once per sample (bb_sky_trig) and every detector contracts its tensor / vertex with them.
struct BBSkyTrig {
    double cph, sph, cth, sth, cps, sps;      // phi = ra - gmst, theta = pi/2 - dec, psi
};
BB_HD BBSkyTrig bb_sky_trig(double ra, double dec, double psi, double gmst_wrapped) {
    const double phi = ra - gmst_wrapped;
    const double theta = BB_PI / 2 - dec;
    BBSkyTrig t;
    t.cph = cos(phi); t.sph = sin(phi); t.cth = cos(theta); t.sth = sin(theta);
    t.cps = cos(psi); t.sps = sin(psi);
    return t;
}
BB_HD void bb_antenna_trig(const double* d, const BBSkyTrig& t, double* fplus, double* fcross) {
    const double cph = t.cph, sph = t.sph, cth = t.cth, sth = t.sth, cps = t.cps, sps = t.sps;
    const double u[3] = {cph * cth, cth * sph, -sth};
    const double v[3] = {-sph, cph, 0.0};
    double m[3], n[3];
    for (int i = 0; i < 3; ++i) {
        m[i] = -u[i] * sps - v[i] * cps;
        n[i] = -u[i] * cps + v[i] * sps;
    }
    double fp = 0.0, fc = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            fp += d[3 * i + j] * (m[i] * m[j] - n[i] * n[j]);
            fc += d[3 * i + j] * (m[i] * n[j] + n[i] * m[j]);
        }
    *fplus = fp;
    *fcross = fc;
}
BB_HD void bb_antenna(const double* d, double ra, double dec, double psi, double gmst_wrapped,
                      double* fplus, double* fcross) {
    const BBSkyTrig t = bb_sky_trig(ra, dec, psi, gmst_wrapped);
    bb_antenna_trig(d, t, fplus, fcross);
}

// time delay from geocentre for a detector vertex [m]
BB_HD double bb_time_delay_trig(const double* vertex, double sth, double cph, double sph, double cth) {
    const double ox = sth * cph, oy = sth * sph, oz = cth;
    // omega . (0 - vertex) / c
    return (ox * (0.0 - vertex[0]) + oy * (0.0 - vertex[1]) + oz * (0.0 - vertex[2])) / BB_C_SI;
}
BB_HD double bb_time_delay(const double* vertex, double ra, double dec, double gmst_wrapped) {
    const double phi = ra - gmst_wrapped;
    const double theta = BB_PI / 2 - dec;
    const double sth = sin(theta);
    return bb_time_delay_trig(vertex, sth, cos(phi), sin(phi), cos(theta));
}

// Detector-based sky frame and detector time reference (bilby/gw/likelihood/base.py:1091-1137
// get_sky_frame_parameters; gw/utils.py:232-256 zenith_azimuth_to_ra_dec; gw/geometry.py:215-258, 346-377).
struct BBFrame {
    int sky_frame;          // 1: columns (RA, DEC) hold (azimuth, zenith) in the frame of a detector pair
    int detector_time;      // 1: column GEOCENT_TIME holds the arrival time at the reference detector
    double rotation[9];     // rotation_matrix_from_delta(vertex_1 - vertex_2), row-major
    double ref_vertex[3];   // vertex of the time-reference detector [m]
};

// in: (azimuth|ra, zenith|dec, reference time)  ->  out: (ra, dec, geocent_time)
BB_HD void bb_sky_frame(const BBFrame& fr, double a, double b, double time, double* ra, double* dec, double* tgeo) {
    double r = a, d = b;
    if (fr.sky_frame) {
        const double sz = sin(b);
        const double o[3] = {sz * cos(a), sz * sin(a), cos(b)};
        const double* R = fr.rotation;
        const double x = R[0] * o[0] + R[1] * o[1] + R[2] * o[2];
        const double y = R[3] * o[0] + R[4] * o[1] + R[5] * o[2];
        const double z = R[6] * o[0] + R[7] * o[1] + R[8] * o[2];
        const double theta = acos(z);
        const double phi = bb_wrap_2pi(atan2(y, x));
        // theta_phi_to_ra_dec with the UNWRAPPED gmst of the reference time, then ra % 2 pi
        r = bb_wrap_2pi(phi + bb_gmst(time));
        d = BB_PI / 2 - theta;
    }
    *ra = r;
    *dec = d;
    *tgeo = fr.detector_time ? time - bb_time_delay(fr.ref_vertex, r, d, bb_wrap_2pi(bb_gmst(time))) : time;
}
