"""TEST INFRASTRUCTURE ONLY (oracle) - CPU restatement of IMRPhenomD.

The reference (bilby) does not contain this arithmetic: it calls the third-party
C library ``lalsimulation`` (lalsuite, version UNPINNED in the reference:
``gw_requirements.txt:2``) through ``bilby/gw/source.py:597-643`` and
``bilby/gw/utils.py:642-684`` (``SimInspiralChooseFDWaveform``).  lalsimulation
is not installed here and its source is not under /root/reference, so this file
restates the *published* algorithm (Husa+ 2016 arXiv:1508.07250, Khan+ 2016
arXiv:1508.07253 Table V / Appendix, as implemented upstream in
``LALSimIMRPhenomD.c`` and ``LALSimIMRPhenomD_internals.c``):

* ``final_spin_0815`` / ``e_rad_0815``         <- FinalSpin0815, EradRational0815
* ``_fit`` tables RHO/V2/GAMMA/SIGMA/BETA/ALPHA  <- rho1_fun ... alpha5Fit
* ``taylorf2_aligned_phasing``                  <- XLALSimInspiralPNPhasing_F2
                                                  (LALSimInspiralPNCoefficients.c)
* ``PhenomDCoefficients``                       <- ComputeIMRPhenomD{Amplitude,Phase}Coefficients,
                                                  ComputeIMRPhenDPhaseConnectionCoefficients,
                                                  init_{amp,phi}_ins_prefactors
* ``phenomd_h22``                               <- IMRPhenomDGenerateFD (core loop)
* ``choose_fd_waveform_phenomd``                <- XLALSimInspiralChooseFDWaveform, IMRPhenomD case
                                                  (LALSimInspiral.c: pfac/cfac polarisation assembly)

Ring-down / damping frequencies: upstream interpolates a tabulated Kerr QNM
data set (QNMData_*, not available offline); here the table is recomputed from
first principles with Leaver's method (oracle/tools/make_qnm_table.py) and
interpolated with the same natural cubic spline.

PARITY STATUS: **lalsimulation parity is UNPINNED** - the reference holds no
stored waveform vectors and lalsimulation cannot be run
here.  What is pinned: the non-LAL part of the path reproduces the reference's
known-answer tests, and the CUDA path is compared against *this* restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import os
import numpy as np
from scipy.interpolate import CubicSpline

# constants mirrored by the reference from LAL (bilby/core/utils/constants.py:3-7)
SPEED_OF_LIGHT = 299792458.0
PARSEC = 3.085677581491367e+16
SOLAR_MASS = 1.988409870698050731911960804878414216e30
GRAVITATIONAL_CONSTANT = 6.6743e-11
GM_SUN = GRAVITATIONAL_CONSTANT * SOLAR_MASS           # = 1.3271244e20 (IAU nominal)
MTSUN_SI = GM_SUN / SPEED_OF_LIGHT ** 3                # 4.925490947641267e-06 s
MRSUN_SI = GM_SUN / SPEED_OF_LIGHT ** 2                # 1476.6250380501247 m
EULER_GAMMA = 0.5772156649015328606065120900824024

F_CUT = 0.2            # Mf at which the model is cut off
AMP_FJOIN_INS = 0.014  # amplitude inspiral -> intermediate join (Mf)
PHI_FJOIN_INS = 0.018  # phase inspiral -> intermediate join (Mf)

_HERE = os.path.dirname(os.path.abspath(__file__))
_QNM = None


def qnm_splines():
    global _QNM
    if _QNM is None:
        tab = np.load(os.path.join(_HERE, "data", "qnm_l2m2n0.npz"))
        _QNM = (CubicSpline(tab["spin"], tab["fring"], bc_type="natural"),
                CubicSpline(tab["spin"], tab["fdamp"], bc_type="natural"))
    return _QNM


# ----------------------------------------------------------------------------
# final state
# ----------------------------------------------------------------------------
def final_spin_0815(eta, chi1, chi2):
    seta = np.sqrt(1.0 - 4.0 * eta)
    m1 = 0.5 * (1.0 + seta)
    m2 = 0.5 * (1.0 - seta)
    s = m1 * m1 * chi1 + m2 * m2 * chi2
    eta2 = eta * eta
    eta3 = eta2 * eta
    eta4 = eta3 * eta
    s2 = s * s
    s3 = s2 * s
    s4 = s3 * s
    return (3.4641016151377544 * eta - 4.399247300629289 * eta2
            + 9.397292189321194 * eta3 - 13.180949901606242 * eta4
            + (1 - 0.0850917821418767 * eta - 5.837029316602263 * eta2) * s
            + (0.1014665242971878 * eta - 2.0967746996832157 * eta2) * s2
            + (-1.3546806617824356 * eta + 4.108962025369336 * eta2) * s3
            + (-0.8676969352555539 * eta + 2.064046835273906 * eta2) * s4)


def e_rad_0815(eta, chi1, chi2):
    seta = np.sqrt(1.0 - 4.0 * eta)
    m1 = 0.5 * (1.0 + seta)
    m2 = 0.5 * (1.0 - seta)
    s = (m1 * m1 * chi1 + m2 * m2 * chi2) / (m1 * m1 + m2 * m2)
    eta2 = eta * eta
    eta3 = eta2 * eta
    eta4 = eta3 * eta
    return ((0.055974469826360077 * eta + 0.5809510763115132 * eta2
             - 0.9606726679372312 * eta3 + 3.352411249771192 * eta4)
            * (1. + (-0.0030302335878845507 - 2.0066110851351073 * eta
                     + 7.7050567802399215 * eta2) * s)
            / (1. + (-0.6714403054720589 - 1.4756929437702908 * eta
                     + 7.304676214885011 * eta2) * s))


def chi_pn(eta, chi1, chi2):
    seta = np.sqrt(1.0 - 4.0 * eta)
    chi_s = 0.5 * (chi1 + chi2)
    chi_a = 0.5 * (chi1 - chi2)
    return chi_s * (1.0 - eta * 76.0 / 113.0) + seta * chi_a


# ----------------------------------------------------------------------------
# phenomenological coefficient fits (Khan+ 2016, Table V).  Each row:
#   c00 + c01*eta + (c10 + c11*eta + c12*eta^2) xi + (c20 + ...) xi^2 + (c30 + ...) xi^3,
# xi = chiPN - 1.
# ----------------------------------------------------------------------------
FIT = {
    "rho1": (3931.8979897196696, -17395.758706812805,
             3132.375545898835, 343965.86092361377, -1.2162565819981997e6,
             -70698.00600428853, 1.383907177859705e6, -3.9662761890979446e6,
             -60017.52423652596, 803515.1181825735, -2.091710365941658e6),
    "rho2": (-40105.47653771657, 112253.0169706701,
             23561.696065836168, -3.476180699403351e6, 1.137593670849482e7,
             754313.1127166454, -1.308476044625268e7, 3.6444584853928134e7,
             596226.612472288, -7.4277901143564405e6, 1.8928977514040343e7),
    "rho3": (83208.35471266537, -191237.7264145924,
             -210916.2454782992, 8.71797508352568e6, -2.6914942420669552e7,
             -1.9889806527362722e6, 3.0888029960154563e7, -8.390870279256162e7,
             -1.4535031953446497e6, 1.7063528990822166e7, -4.2748659731120914e7),
    "v2": (0.8149838730507785, 2.5747553517454658,
           1.1610198035496786, -2.3627771785551537, 6.771038707057573,
           0.7570782938606834, -2.7256896890432474, 7.1140380397149965,
           0.1766934149293479, -0.7978690983168183, 2.1162391502005153),
    "gamma1": (0.006927402739328343, 0.03020474290328911,
               0.006308024337706171, -0.12074130661131138, 0.26271598905781324,
               0.0034151773647198794, -0.10779338611188374, 0.27098966966891747,
               0.0007374185938559283, -0.02749621038376281, 0.0733150789135702),
    "gamma2": (1.010344404799477, 0.0008993122007234548,
               0.283949116804459, -4.049752962958005, 13.207828172665366,
               0.10396278486805426, -7.025059158961947, 24.784892370130475,
               0.03093202475605892, -2.6924023896851663, 9.609374464684983),
    "gamma3": (1.3081615607036106, -0.005537729694807678,
               -0.06782917938621007, -0.6689834970767117, 3.403147966134083,
               -0.05296577374411866, -0.9923793203111362, 4.820681208409587,
               -0.006134139870393713, -0.38429253308696365, 1.7561754421985984),
    "sigma1": (2096.551999295543, 1463.7493168261553,
               1312.5493286098522, 18307.330017082117, -43534.1440746107,
               -833.2889543511114, 32047.31997183187, -108609.45037520859,
               452.25136398112204, 8353.439546391714, -44531.3250037322),
    "sigma2": (-10114.056472621156, -44631.01109458185,
               -6541.308761668722, -266959.23419307504, 686328.3229317984,
               3405.6372187679685, -437507.7208209015, 1.6318171307344697e6,
               -7462.648563007646, -114585.25177153319, 674402.4689098676),
    "sigma3": (22933.658273436497, 230960.00814979506,
               14961.083974183695, 1.1940181342318142e6, -3.1042239693052764e6,
               -3038.166617199259, 1.8720322849093592e6, -7.309145012085539e6,
               42738.22871475411, 467502.018616601, -3.064853498512499e6),
    "sigma4": (-14621.71522218357, -377812.8579387104,
               -9608.682631509726, -1.7108925257214056e6, 4.332924601416521e6,
               -22366.683262266528, -2.5019716386377467e6, 1.0274495902259542e7,
               -85360.30079034246, -570025.3441737515, 4.396844346849777e6),
    "beta1": (97.89747327985583, -42.659730877489224,
              153.48421037904913, -1417.0620760768954, 2752.8614143665027,
              138.7406469558649, -1433.6585075135881, 2857.7418952430758,
              41.025109467376126, -423.680737974639, 850.3594335657173),
    "beta2": (-3.282701958759534, -9.051384468245866,
              -12.415449742258042, 55.4716447709787, -106.05109938966335,
              -11.953044553690658, 76.80704618365418, -155.33172948098394,
              -3.4129261592393263, 25.572377569952536, -54.408036707740465),
    "beta3": (-0.000025156429818799565, 0.000019750256942201327,
              -0.000018370671469295915, 0.000021886317041311973, 0.00008250240316860033,
              7.157371250566708e-6, -0.000055780000112270685, 0.00019142082884072178,
              5.447166261464217e-6, -0.00003220610095021982, 0.00007974016714984341),
    "alpha1": (43.31514709695348, 638.6332679188081,
               -32.85768747216059, 2415.8938269370315, -5766.875169379177,
               -61.85459307173841, 2953.967762459948, -8986.29057591497,
               -21.571435779762044, 981.2158224673428, -3239.5664895930286),
    "alpha2": (-0.07020209449091723, -0.16269798450687084,
               -0.1872514685185499, 1.138313650449945, -2.8334196304430046,
               -0.17137955686840617, 1.7197549338119527, -4.539717148261272,
               -0.049983437357548705, 0.6062072055948309, -1.682769616644546),
    "alpha3": (9.5988072383479, -397.05438595557433,
               16.202126189517813, -1574.8286986717037, 3600.3410843831093,
               27.092429659075467, -1786.482357315139, 5152.919378666511,
               11.175710130033895, -577.7999423177481, 1808.730762932043),
    "alpha4": (-0.02989487384493607, 1.4022106448583738,
               -0.07356049468633846, 0.8337006542278661, 0.2240008282397391,
               -0.055202870001177226, 0.5667186343606578, 0.7186931973380503,
               -0.015507437354325743, 0.15750322779277187, 0.21076815715176228),
    "alpha5": (0.9974408278363099, -0.007884449714907203,
               -0.059046901195591035, 1.3958712396764088, -4.516631601676276,
               -0.05585343136869692, 1.7516580039343603, -5.990208965347804,
               -0.017945336522161195, 0.5965097794825992, -2.0608879367971804),
}
FIT_ORDER = ("rho1", "rho2", "rho3", "v2", "gamma1", "gamma2", "gamma3",
             "sigma1", "sigma2", "sigma3", "sigma4", "beta1", "beta2", "beta3",
             "alpha1", "alpha2", "alpha3", "alpha4", "alpha5")


def _fit(name, eta, xi):
    c = FIT[name]
    eta2 = eta * eta
    return (c[0] + c[1] * eta
            + (c[2] + c[3] * eta + c[4] * eta2) * xi
            + (c[5] + c[6] * eta + c[7] * eta2) * xi * xi
            + (c[8] + c[9] * eta + c[10] * eta2) * xi * xi * xi)


# ----------------------------------------------------------------------------
# TaylorF2 aligned-spin phasing coefficients (3.5PN, all spin orders)
# ----------------------------------------------------------------------------
def taylorf2_aligned_phasing(m1, m2, chi1, chi2, qm_def1=1.0, qm_def2=1.0):
    """Returns (v[0..7], vlogv[0..7]) already multiplied by 3/(128 eta).
    m1 >= m2 in any common unit."""
    M = m1 + m2
    m1M = m1 / M
    m2M = m2 / M
    eta = m1 * m2 / (M * M)
    d = (m1 - m2) / M
    pi = np.pi
    pfaN = 3.0 / (128.0 * eta)
    v = np.zeros(8)
    vl = np.zeros(8)
    v[0] = 1.0
    v[1] = 0.0
    v[2] = 5.0 * (743.0 / 84.0 + 11.0 * eta) / 9.0
    v[3] = -16.0 * pi
    v[4] = 5.0 * (3058.673 / 7.056 + 5429.0 / 7.0 * eta + 617.0 * eta * eta) / 72.0
    v[5] = 5.0 / 9.0 * (7729.0 / 84.0 - 13.0 * eta) * pi
    vl[5] = 5.0 / 3.0 * (7729.0 / 84.0 - 13.0 * eta) * pi
    v[6] = ((11583.231236531 / 4.694215680 - 640.0 / 3.0 * pi * pi - 6848.0 / 21.0 * EULER_GAMMA)
            + eta * (-15737.765635 / 3.048192 + 2255. / 12. * pi * pi)
            + eta * eta * 76055.0 / 1728.0
            - eta * eta * eta * 127825.0 / 1296.0)
    v[6] += (-6848.0 / 21.0) * np.log(4.0)
    vl[6] = -6848.0 / 21.0
    v[7] = pi * (77096675. / 254016. + 378515. / 1512. * eta - 74045. / 756. * eta * eta)

    chi1sq = chi1 * chi1
    chi2sq = chi2 * chi2
    chi1dotchi2 = chi1 * chi2
    SL = m1M * m1M * chi1 + m2M * m2M * chi2
    dSigmaL = d * (m2M * chi2 - m1M * chi1)

    pn_sigma = eta * (721. / 48. * chi1 * chi2 - 247. / 48. * chi1dotchi2)
    pn_sigma += (720. * qm_def1 - 1.) / 96.0 * m1M * m1M * chi1 * chi1
    pn_sigma += (720. * qm_def2 - 1.) / 96.0 * m2M * m2M * chi2 * chi2
    pn_sigma -= (240. * qm_def1 - 7.) / 96.0 * m1M * m1M * chi1sq
    pn_sigma -= (240. * qm_def2 - 7.) / 96.0 * m2M * m2M * chi2sq

    pn_ss3 = (326.75 / 1.12 + 557.5 / 1.8 * eta) * eta * chi1 * chi2
    pn_ss3 += ((4703.5 / 8.4 + 2935. / 6. * m1M - 120. * m1M * m1M) * qm_def1
               + (-4108.25 / 6.72 - 108.5 / 1.2 * m1M + 125.5 / 3.6 * m1M * m1M)) * m1M * m1M * chi1sq
    pn_ss3 += ((4703.5 / 8.4 + 2935. / 6. * m2M - 120. * m2M * m2M) * qm_def2
               + (-4108.25 / 6.72 - 108.5 / 1.2 * m2M + 125.5 / 3.6 * m2M * m2M)) * m2M * m2M * chi2sq

    pn_gamma = (554345. / 1134. + 110. * eta / 9.) * SL + (13915. / 84. - 10. * eta / 3.) * dSigmaL

    v[7] += ((-8980424995. / 762048. + 6586595. * eta / 756. - 305. * eta * eta / 36.) * SL
             - (170978035. / 48384. - 2876425. * eta / 672. - 4735. * eta * eta / 144.) * dSigmaL)
    v[6] += pi * (3760. * SL + 1490. * dSigmaL) / 3. + pn_ss3
    v[5] += -1. * pn_gamma
    vl[5] += -3. * pn_gamma
    v[4] += -10. * pn_sigma
    v[3] += 188. * SL / 3. + 25. * dSigmaL
    return v * pfaN, vl * pfaN


def subtract_3pn_ss(m1, m2, chi1, chi2):
    M = m1 + m2
    eta = m1 * m2 / (M * M)
    m1M = m1 / M
    m2M = m2 / M
    pn_ss3 = (326.75 / 1.12 + 557.5 / 1.8 * eta) * eta * chi1 * chi2
    pn_ss3 += ((4703.5 / 8.4 + 2935. / 6. * m1M - 120. * m1M * m1M)
               + (-4108.25 / 6.72 - 108.5 / 1.2 * m1M + 125.5 / 3.6 * m1M * m1M)) * m1M * m1M * chi1 * chi1
    pn_ss3 += ((4703.5 / 8.4 + 2935. / 6. * m2M - 120. * m2M * m2M)
               + (-4108.25 / 6.72 - 108.5 / 1.2 * m2M + 125.5 / 3.6 * m2M * m2M)) * m2M * m2M * chi2 * chi2
    return pn_ss3


# ----------------------------------------------------------------------------
class PhenomDCoefficients:
    """All per-binary constants of the model (masses in solar masses, m1 >= m2)."""

    def __init__(self, m1, m2, chi1, chi2):
        if m2 > m1:
            m1, m2 = m2, m1
            chi1, chi2 = chi2, chi1
        self.m1, self.m2, self.chi1, self.chi2 = m1, m2, chi1, chi2
        M = m1 + m2
        self.M = M
        eta = m1 * m2 / (M * M)
        if eta > 0.25:
            eta = 0.25
        self.eta = eta
        self.seta = np.sqrt(1.0 - 4.0 * eta)
        self.chi = chi_pn(eta, chi1, chi2)
        xi = self.chi - 1.0
        self.finspin = final_spin_0815(eta, chi1, chi2)
        ring, damp = qnm_splines()
        erad = e_rad_0815(eta, chi1, chi2)
        # the QNM table spans |a| <= 0.9999; clamp (upstream's GSL spline errors outside its table)
        a_f = min(max(self.finspin, ring.x[0]), ring.x[-1])
        self.fRD = float(ring(a_f)) / (1.0 - erad)
        self.fDM = float(damp(a_f)) / (1.0 - erad)
        for name in FIT_ORDER:
            setattr(self, name, _fit(name, eta, xi))
        # ---- amplitude inspiral prefactors (init_amp_ins_prefactors)
        pi = np.pi
        chi12 = chi1 * chi1
        chi22 = chi2 * chi2
        eta2 = eta * eta
        eta3 = eta2 * eta
        Seta = self.seta
        self.amp0 = np.sqrt(2.0 * eta / 3.0) * pi ** (-1.0 / 6.0)
        self.A = np.zeros(10)  # coefficient of Mf^(k/3), k = 0..9
        self.A[0] = 1.0
        self.A[2] = ((-969 + 1804 * eta) * pi ** (2.0 / 3.0)) / 672.
        self.A[3] = ((chi1 * (81 * (1 + Seta) - 44 * eta) + chi2 * (81 - 81 * Seta - 44 * eta)) * pi) / 48.
        self.A[4] = ((-27312085.0 - 10287648 * chi22 - 10287648 * chi12 * (1 + Seta) + 10287648 * chi22 * Seta
                      + 24 * (-1975055 + 857304 * chi12 - 994896 * chi1 * chi2 + 857304 * chi22) * eta
                      + 35371056 * eta2) * pi ** (4.0 / 3.0)) / 8.128512e6
        self.A[5] = (pi ** (5.0 / 3.0) * (chi2 * (-285197 * (-1 + Seta) + 4 * (-91902 + 1579 * Seta) * eta - 35632 * eta2)
                                          + chi1 * (285197 * (1 + Seta) - 4 * (91902 + 1579 * Seta) * eta - 35632 * eta2)
                                          + 42840 * (-1.0 + 4 * eta) * pi)) / 32256.
        self.A[6] = - (pi * pi * (-336 * (-3248849057.0 + 2943675504 * chi12 - 3339284256 * chi1 * chi2
                                         + 2943675504 * chi22) * eta2
                                  - 324322727232 * eta3
                                  - 7 * (-177520268561 + 107414046432 * chi22 + 107414046432 * chi12 * (1 + Seta)
                                         - 107414046432 * chi22 * Seta
                                         + 11087290368 * (chi1 + chi2 + chi1 * Seta - chi2 * Seta) * pi)
                                  + 12 * eta * (-545384828789 - 176491177632 * chi1 * chi2 + 202603761360 * chi22
                                                + 77616 * chi12 * (2610335 + 995766 * Seta)
                                                - 77287373856 * chi22 * Seta
                                                + 5841690624 * (chi1 + chi2) * pi + 21384760320 * pi * pi))
                       ) / 6.0085960704e10
        self.A[7] = self.rho1
        self.A[8] = self.rho2
        self.A[9] = self.rho3
        # ---- amplitude peak and intermediate collocation
        g2, g3 = self.gamma2, self.gamma3
        if not (g2 > 1):
            self.fmaxCalc = abs(self.fRD + (self.fDM * (-1 + np.sqrt(1 - g2 * g2)) * g3) / g2)
        else:
            self.fmaxCalc = abs(self.fRD + (-self.fDM * g3) / g2)
        f1 = AMP_FJOIN_INS
        f3 = self.fmaxCalc
        f2 = 0.5 * (f1 + f3)
        v1 = self.amp_ins(f1)
        v3 = self.amp_mrd(f3)
        d1 = self.damp_ins(f1)
        d3 = self.damp_mrd(f3)
        self.f1, self.f2, self.f3 = f1, f2, f3
        # A(f) = sum_k delta_k f^k with A(f1)=v1, A(f2)=v2, A(f3)=v3, A'(f1)=d1, A'(f3)=d3.
        # Solved in the shifted/scaled variable x=(f-f1)/(f3-f1) (same polynomial, better conditioned
        # than upstream's closed form in raw powers of f).
        w = f3 - f1
        mat = np.array([[1, 0, 0, 0, 0],
                        [1, 0.5, 0.25, 0.125, 0.0625],
                        [1, 1, 1, 1, 1],
                        [0, 1, 0, 0, 0],
                        [0, 1, 2, 3, 4]], dtype=float)
        rhs = np.array([v1, self.v2, v3, d1 * w, d3 * w])
        self.delta_x = np.linalg.solve(mat, rhs)   # coefficients in x
        self.amp_int_w = w
        # ---- PN phasing
        pv, pvl = taylorf2_aligned_phasing(m1, m2, chi1, chi2)
        pv[6] -= subtract_3pn_ss(m1, m2, chi1, chi2) * pv[0]
        self.pn_v, self.pn_vlogv = pv, pvl
        # phase-inspiral prefactors in powers of Mf (init_phi_ins_prefactors)
        p13 = pi ** (1.0 / 3.0)
        P = {}
        P["initial_phasing"] = pv[5] - pi / 4.0
        P["two_thirds"] = pv[7] * p13 * p13
        P["third"] = pv[6] * p13
        P["third_with_logv"] = pvl[6] * p13
        P["logv"] = pvl[5]
        P["minus_third"] = pv[4] / p13
        P["minus_two_thirds"] = pv[3] / (p13 * p13)
        P["minus_one"] = pv[2] / pi
        P["minus_four_thirds"] = pv[1] / (p13 ** 4)
        P["minus_five_thirds"] = pv[0] / (p13 ** 5)
        P["one"] = self.sigma1
        P["four_thirds"] = self.sigma2 * 0.75
        P["five_thirds"] = self.sigma3 * 0.6
        P["two"] = self.sigma4 * 0.5
        self.P = P
        # ---- phase connection (ComputeIMRPhenDPhaseConnectionCoefficients)
        self.fInsJoin = PHI_FJOIN_INS
        self.fMRDJoin = 0.5 * self.fRD
        self.C1Int = 0.0
        self.C2Int = 0.0
        self.C1MRD = 0.0
        self.C2MRD = 0.0
        fi = self.fInsJoin
        self.C2Int = self.dphi_ins(fi) - self.dphi_int(fi)
        self.C1Int = self.phi_ins(fi) - self.phi_int(fi)   # C1Int = 0 and C2Int term included in phi_int
        # (phi_int adds C1Int + C2Int*f with C1Int still zero here)
        fm = self.fMRDJoin
        self.C2MRD = self.dphi_int(fm) + self.C2Int - self.dphi_mrd(fm)
        self.C1MRD = self.phi_int(fm) - self.phi_mrd(fm)   # C1MRD = 0 inside phi_mrd here

    # ----- amplitude pieces (all WITHOUT the amp0 * Mf^(-7/6) prefactor)
    def amp_ins(self, f):
        x = np.cbrt(f)
        out = 0.0
        for k in range(9, -1, -1):
            out = out * x + self.A[k]
        return out

    def damp_ins(self, f):
        x = np.cbrt(f)
        out = 0.0
        for k in range(9, 0, -1):
            out = out * x + self.A[k] * (k / 3.0)
        # d/df sum A_k f^(k/3) = sum (k/3) A_k f^(k/3 - 1) = x^-2 * sum_{k>=1} (k/3) A_k x^(k-1)
        return out / (x * x)

    def amp_mrd(self, f):
        fdg3 = self.fDM * self.gamma3
        d = f - self.fRD
        return np.exp(-d * self.gamma2 / fdg3) * (fdg3 * self.gamma1) / (d * d + fdg3 * fdg3)

    def damp_mrd(self, f):
        fdg3 = self.fDM * self.gamma3
        d = f - self.fRD
        den = d * d + fdg3 * fdg3
        e = np.exp(-d * self.gamma2 / fdg3)
        return e * fdg3 * self.gamma1 * (-self.gamma2 / fdg3 / den - 2.0 * d / (den * den))

    def amp_int(self, f):
        x = (f - self.f1) / self.amp_int_w
        c = self.delta_x
        return c[0] + x * (c[1] + x * (c[2] + x * (c[3] + x * c[4])))

    # ----- phase pieces
    def phi_ins(self, f):
        P = self.P
        x = np.cbrt(f)
        v = x * np.pi ** (1.0 / 3.0)
        logv = np.log(v)
        ph = P["initial_phasing"]
        ph = ph + P["two_thirds"] * x * x
        ph = ph + P["third"] * x
        ph = ph + P["third_with_logv"] * logv * x
        ph = ph + P["logv"] * logv
        ph = ph + P["minus_third"] / x
        ph = ph + P["minus_two_thirds"] / (x * x)
        ph = ph + P["minus_one"] / f
        ph = ph + P["minus_four_thirds"] / (f * x)
        ph = ph + P["minus_five_thirds"] / (f * x * x)
        ph = ph + (P["one"] * f + P["four_thirds"] * f * x + P["five_thirds"] * f * x * x
                   + P["two"] * f * f) / self.eta
        return ph

    def dphi_ins(self, f):
        pv, pvl = self.pn_v, self.pn_vlogv
        pi = np.pi
        v = np.cbrt(pi * f)
        logv = np.log(v)
        v2 = v * v
        v3 = v2 * v
        v4 = v3 * v
        v5 = v4 * v
        v6 = v5 * v
        v7 = v6 * v
        v8 = v7 * v
        d = 2.0 * pv[7] * v7
        d += (pv[6] + pvl[6] * (1.0 + logv)) * v6
        d += pvl[5] * v5
        d += -1.0 * pv[4] * v4
        d += -2.0 * pv[3] * v3
        d += -3.0 * pv[2] * v2
        d += -4.0 * pv[1] * v
        d += -5.0 * pv[0]
        d /= v8 * 3.0 / pi
        x = np.cbrt(f)
        d += (self.sigma1 + self.sigma2 * x + self.sigma3 * x * x + self.sigma4 * f) / self.eta
        return d

    def phi_int(self, f):
        return ((self.beta1 * f - self.beta3 / (3.0 * f ** 3) + self.beta2 * np.log(f)) / self.eta
                + self.C1Int + self.C2Int * f)

    def dphi_int(self, f):
        """derivative of the beta part only (upstream DPhiIntAnsatz)."""
        return (self.beta1 + self.beta3 / f ** 4 + self.beta2 / f) / self.eta

    def phi_mrd(self, f):
        return ((-(self.alpha2 / f) + (4.0 / 3.0) * (self.alpha3 * f ** 0.75) + self.alpha1 * f
                 + self.alpha4 * np.arctan((f - self.alpha5 * self.fRD) / self.fDM)) / self.eta
                + self.C1MRD + self.C2MRD * f)

    def dphi_mrd(self, f):
        """derivative of the alpha part only (upstream DPhiMRD)."""
        x = (f - self.alpha5 * self.fRD) / self.fDM
        return (self.alpha1 + self.alpha2 / (f * f) + self.alpha3 / f ** 0.25
                + self.alpha4 / (self.fDM * (1.0 + x * x))) / self.eta

    # ----- full piecewise functions of Mf (arrays)
    def amplitude(self, Mf):
        Mf = np.asarray(Mf, dtype=float)
        pre = self.amp0 * Mf ** (-7.0 / 6.0)
        out = np.where(Mf < AMP_FJOIN_INS, self.amp_ins(Mf),
                       np.where(Mf >= self.fmaxCalc, self.amp_mrd(Mf), self.amp_int(Mf)))
        return pre * out

    def phase(self, Mf):
        Mf = np.asarray(Mf, dtype=float)
        return np.where(Mf < self.fInsJoin, self.phi_ins(Mf),
                        np.where(Mf >= self.fMRDJoin, self.phi_mrd(Mf), self.phi_int(Mf)))


class WaveformDomainError(Exception):
    """Restates lalsimulation's XLAL_EDOM ('Input domain error'): the reference maps it to
    ``None`` -> likelihood sentinel (bilby/gw/source.py:644-662)."""


def phenomd_h22(frequencies, m1, m2, chi1, chi2, distance_m, phi_ref, f_ref, f_min, f_max, delta_f=None,
                sequence=False):
    """h22-like strain htilde(f) (before the inclination factors), complex128 array on
    ``frequencies`` (Hz, uniform grid starting at 0 when ``delta_f`` is given - then upstream's
    index rule  i in [int(f_min/df), int(f_max'/df))  decides which bins are filled).

    ``sequence=True`` restates the frequency-sequence entry point (SimInspiralChooseFDWaveformSequence ->
    IMRPhenomDFrequencySequence, called from bilby/gw/source.py:1124-1128): f_min is the first frequency
    of the sequence, f_max is ignored and EVERY frequency of the sequence is evaluated with the ansatz
    (no f_min / f_CUT zeroing); the (fCut <= f_min) domain check still applies.

    Restates IMRPhenomDGenerateFD."""
    if m1 <= 0 or m2 <= 0 or distance_m <= 0:
        raise WaveformDomainError("masses and distance must be positive")
    if abs(chi1) > 1.0 or abs(chi2) > 1.0:
        raise WaveformDomainError("Spins outside the range [-1,1] are not supported")
    c = PhenomDCoefficients(m1, m2, chi1, chi2)
    M = m1 + m2
    M_sec = M * MTSUN_SI
    f_cut = F_CUT / M_sec
    frequencies = np.asarray(frequencies, dtype=float)
    if sequence:
        f_min, f_max = float(frequencies[0]), 0.0
    if f_ref == 0.0:
        f_ref = f_min
    f_max_prime = f_cut if f_max == 0 else min(f_max, f_cut)
    if f_max_prime <= f_min:
        raise WaveformDomainError("(fCut = %g Hz) <= f_min = %g" % (f_cut, f_min))
    amp0 = 2.0 * np.sqrt(5.0 / (64.0 * np.pi)) * M * MRSUN_SI * M * MTSUN_SI / distance_m
    frequencies = np.asarray(frequencies, dtype=float)
    out = np.zeros(len(frequencies), dtype=complex)
    if sequence:
        sel = np.ones(len(frequencies), dtype=bool)
    elif delta_f is not None:
        ind_min = int(f_min / delta_f)
        ind_max = int(f_max_prime / delta_f)
        sel = np.zeros(len(frequencies), dtype=bool)
        sel[ind_min:min(ind_max, len(frequencies))] = True
    else:
        sel = (frequencies >= f_min) & (frequencies <= f_max_prime)
    t0 = c.dphi_mrd(c.fmaxCalc)
    MfRef = M_sec * f_ref
    phi_precalc = 2.0 * phi_ref + float(c.phase(np.array([MfRef]))[0])
    Mf = M_sec * frequencies[sel]
    amp = c.amplitude(Mf)
    phi = c.phase(Mf) - (t0 * (Mf - MfRef) + phi_precalc)
    out[sel] = amp0 * amp * np.exp(-1j * phi)
    return out


def choose_fd_waveform_phenomd(frequencies, m1, m2, s1z, s2z, distance_m, inclination, phi_ref,
                               f_min, f_max, f_ref, delta_f=None, sequence=False):
    """(hplus, hcross) as SimInspiralChooseFDWaveform[Sequence] assembles them for IMRPhenomD."""
    h = phenomd_h22(frequencies, m1, m2, s1z, s2z, distance_m, phi_ref, f_ref, f_min, f_max, delta_f, sequence)
    cfac = np.cos(inclination)
    pfac = 0.5 * (1.0 + cfac * cfac)
    return pfac * h, -1j * cfac * h
