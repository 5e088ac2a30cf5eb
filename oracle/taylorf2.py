"""TEST INFRASTRUCTURE ONLY (oracle) - CPU restatement of TaylorF2 (+ tidal terms).

Like IMRPhenomD this arithmetic is NOT in the reference: bilby calls lalsimulation
(``bilby/gw/source.py:351-432 lal_binary_neutron_star -> :597-643``).  Restated from the published
algorithm as implemented upstream in ``LALSimInspiralTaylorF2.c`` (XLALSimInspiralTaylorF2Core) and
``LALSimInspiralPNCoefficients.c`` (XLALSimInspiralPNPhasing_F2: 3.5PN point-particle + spin terms,
tidal 5PN / 6PN [Vines+ 2011, arXiv:1101.1673], 6.5 / 7 / 7.5PN [Damour+ 2012, arXiv:1203.4352;
Henry+ 2020, arXiv:2005.13367]); amplitude at Newtonian order (pn_amplitude_order = 0).
Quadrupole-monopole parameters keep their black-hole value (bilby inserts only TidalLambda1/2,
source.py:502-549).  PARITY STATUS: lalsimulation parity UNPINNED (see oracle/phenomd.py).
"""
import numpy as np

from . import phenomd as _pd


def tidal_coefficients(m1M, m2M, lambda1, lambda2, eta):
    """v[10], v[12], v[13], v[14], v[15] of the phasing series (already times 3/(128 eta))."""
    pfaN = 3.0 / (128.0 * eta)

    def c10(x):
        return (-288.0 + 264.0 * x) * x ** 4

    def c12(x):
        return (-15895.0 / 28.0 + 4595.0 / 28.0 * x + 5715.0 / 14.0 * x * x - 325.0 / 7.0 * x ** 3) * x ** 4

    def c13(x):
        return x ** 4 * 24.0 * (12.0 - 11.0 * x) * np.pi

    def c14(x):
        return -x ** 4 * 5.0 * (193986935.0 / 571536.0 - 14415613.0 / 381024.0 * x - 57859.0 / 378.0 * x * x
                                - 209495.0 / 1512.0 * x ** 3 + 965.0 / 54.0 * x ** 4 - 4.0 * x ** 5)

    def c15(x):
        return x ** 4 * 1.0 / 28.0 * np.pi * (27719.0 - 22415.0 * x + 7598.0 * x * x - 10520.0 * x ** 3)

    out = {}
    for k, fn in ((10, c10), (12, c12), (13, c13), (14, c14), (15, c15)):
        out[k] = pfaN * (lambda1 * fn(m1M) + lambda2 * fn(m2M))
    return out


def taylorf2_series(v, pv, pvl, tidal):
    logv = np.log(v)
    ph = pv[7] * v ** 7 + (pv[6] + pvl[6] * logv) * v ** 6 + (pv[5] + pvl[5] * logv) * v ** 5 \
        + pv[4] * v ** 4 + pv[3] * v ** 3 + pv[2] * v ** 2 + pv[1] * v + pv[0]
    for k, c in tidal.items():
        ph = ph + c * v ** k
    return ph / v ** 5


def taylorf2_h(frequencies, m1, m2, chi1, chi2, lambda1, lambda2, distance_m, phi_ref, f_ref, f_min, f_max,
               delta_f=None, sequence=False):
    """htilde(f) before the inclination factors (XLALSimInspiralTaylorF2Core).  m1 >= m2 is NOT required by
    upstream; the PN coefficient routine is symmetric up to the (m1-m2) sign conventions it uses itself.
    ``sequence=True``: the frequency-sequence entry point (bilby/gw/source.py:1124-1128) - every frequency
    of the sequence is evaluated, no f_min / f_ISCO cut and no time shift."""
    if m1 <= 0 or m2 <= 0 or distance_m <= 0:
        raise _pd.WaveformDomainError("masses and distance must be positive")
    M = m1 + m2
    eta = m1 * m2 / (M * M)
    m_sec = M * _pd.MTSUN_SI
    piM = np.pi * m_sec
    f_isco = (1.0 / np.sqrt(6.0)) ** 3 / piM
    f_end = f_isco if f_max == 0 else f_max
    if not sequence and f_end <= f_min:
        raise _pd.WaveformDomainError("f_max <= f_min")
    pv, pvl = _pd.taylorf2_aligned_phasing(m1, m2, chi1, chi2)
    tidal = tidal_coefficients(m1 / M, m2 / M, lambda1, lambda2, eta)
    amp0 = -4.0 * m1 * m2 / distance_m * _pd.MRSUN_SI * _pd.MTSUN_SI * np.sqrt(np.pi / 12.0)
    frequencies = np.asarray(frequencies, dtype=float)
    out = np.zeros(len(frequencies), dtype=complex)
    if sequence:
        sel = frequencies > 0
    elif delta_f is not None:
        i_start = int(np.ceil(f_min / delta_f))
        n = int(f_end / delta_f + 1)
        sel = np.zeros(len(frequencies), dtype=bool)
        sel[i_start:min(n, len(frequencies))] = True
    else:
        sel = (frequencies >= f_min) & (frequencies <= f_end)
    f = frequencies[sel]
    v = np.cbrt(piM * f)
    ref_phasing = 0.0
    if f_ref != 0.0:
        ref_phasing = float(taylorf2_series(np.cbrt(piM * f_ref), pv, pvl, tidal))
    phasing = taylorf2_series(v, pv, pvl, tidal) - 2.0 * phi_ref - ref_phasing
    # Newtonian amplitude: sqrt(-dE/dv / flux) * v = sqrt(5 / (32 eta)) v^(-7/2)
    amp = amp0 * np.sqrt(5.0 / (32.0 * eta)) * v ** (-3.5)
    out[sel] = amp * np.exp(-1j * (phasing - np.pi / 4))
    return out


def choose_fd_waveform_taylorf2(frequencies, m1, m2, s1z, s2z, lambda1, lambda2, distance_m, inclination,
                                phi_ref, f_min, f_max, f_ref, delta_f=None, sequence=False):
    h = taylorf2_h(frequencies, m1, m2, s1z, s2z, lambda1, lambda2, distance_m, phi_ref, f_ref, f_min, f_max,
                   delta_f, sequence)
    cfac = np.cos(inclination)
    pfac = 0.5 * (1.0 + cfac * cfac)
    return pfac * h, -1j * cfac * h
