"""Packing of converted parameters into the C ABI's row-major double[n][BB_NPARAM] matrix
(include/bilby_b200.h enum bb_param)."""
import numpy as np

NPARAM = 16
MASS_1, MASS_2, CHI_1, CHI_2, LUMINOSITY_DISTANCE, THETA_JN, PSI, PHASE, RA, DEC, GEOCENT_TIME, \
    TIME_JITTER, LAMBDA_1, LAMBDA_2 = range(14)

MARG_PHASE, MARG_DISTANCE, MARG_TIME = 1, 2, 4
APPROXIMANTS = {"IMRPhenomD": 0, "TaylorF2": 1}


def aligned_spin(a, tilt, xp):
    """a*cos(tilt) for the aligned-spin shortcut of bilby/gw/conversion.py:146-153."""
    return a * xp.cos(tilt if xp is np else xp.as_tensor(tilt))


def pack_rows(converted, n, xp, device=None, check_aligned=True):
    """converted: dict of scalars / length-n arrays with mass_1, mass_2, a_i, tilt_i, ... -> [n,16]."""
    if xp is np:
        rows = np.zeros((n, NPARAM), dtype=np.float64)
    else:
        rows = xp.zeros((n, NPARAM), dtype=xp.float64, device=device)

    def col(key, default=None):
        v = converted.get(key, default)
        if v is None:
            raise KeyError(f"parameter '{key}' is required")
        if xp is not np:
            v = xp.as_tensor(v, dtype=xp.float64, device=device)
        return v

    a1, a2 = col("a_1", 0.0), col("a_2", 0.0)
    t1, t2 = col("tilt_1", 0.0), col("tilt_2", 0.0)
    if check_aligned:
        for a, t in ((a1, t1), (a2, t2)):
            aa = np.asarray(a.cpu() if hasattr(a, "cpu") else a, dtype=float)
            tt = np.asarray(t.cpu() if hasattr(t, "cpu") else t, dtype=float)
            ok = (aa == 0) | (tt == 0) | (tt == np.pi)
            if not np.all(ok):
                raise ValueError("IMRPhenomD / TaylorF2 are aligned-spin models: tilts must be 0 or pi "
                                 "(the reference raises a lalsimulation error here)")
    rows[:, MASS_1] = col("mass_1")
    rows[:, MASS_2] = col("mass_2")
    rows[:, CHI_1] = aligned_spin(a1, t1, xp)
    rows[:, CHI_2] = aligned_spin(a2, t2, xp)
    rows[:, LUMINOSITY_DISTANCE] = col("luminosity_distance")
    rows[:, THETA_JN] = col("theta_jn")
    rows[:, PSI] = col("psi", 0.0)
    rows[:, PHASE] = col("phase")
    rows[:, RA] = col("ra", 0.0)
    rows[:, DEC] = col("dec", 0.0)
    rows[:, GEOCENT_TIME] = col("geocent_time", 0.0)
    rows[:, TIME_JITTER] = col("time_jitter", 0.0)
    rows[:, LAMBDA_1] = col("lambda_1", 0.0)
    rows[:, LAMBDA_2] = col("lambda_2", 0.0)
    return rows
