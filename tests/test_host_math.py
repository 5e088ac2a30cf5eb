"""The library's host+device inline math (bilby_b200/csrc/*.cuh) compiled with g++ (tests/host_check.cpp,
test infrastructure) against the oracle: per-sample prologue + per-bin amplitude/phase (as detector-frame
strain), GMST, ln I0 and the FITPACK bicubic evaluation."""
import os
import subprocess

import numpy as np
import pytest

from oracle import cbc_likelihood as ocl

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def host_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("hc") / "host_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-DBB_HAVE_TAYLORF2", "-o", exe,
                           os.path.join(ROOT, "tests", "host_check.cpp")])

    def run(mode, arr):
        d = os.path.dirname(exe)
        np.asarray(arr, dtype=np.float64).tofile(os.path.join(d, "in.bin"))
        subprocess.check_call([exe, mode, os.path.join(d, "in.bin"), os.path.join(d, "out.bin")])
        return np.fromfile(os.path.join(d, "out.bin"))
    return run


def test_phenomd_prologue_and_bins_vs_oracle(host_check):
    g = np.load(os.path.join(GOLDEN, "bbh_4s_noise_H1L1V1.npz"))
    st = float(g["start_time"])
    ifos = [ocl.OracleInterferometer(n, 2048.0, 4.0, st) for n in ("H1", "L1", "V1")]
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_") and k != "param_time_jitter"}
    n = len(draws["chirp_mass"])
    conv = [ocl.convert_to_lal_binary_black_hole_parameters({k: draws[k][i] for k in draws}) for i in range(n)]
    params = np.zeros((n, 16))
    for i, c in enumerate(conv):
        params[i, :11] = [c["mass_1"], c["mass_2"], c["a_1"] * np.cos(c["tilt_1"]), c["a_2"] * np.cos(c["tilt_2"]),
                          c["luminosity_distance"], c["theta_jn"], c["psi"], c["phase"], c["ra"], c["dec"],
                          c["geocent_time"]]
    nf = len(ifos[0].frequency_array)
    hdr = [n, 3, nf, 4.0, 2048.0, st, 0, 50.0, 20.0, 1024.0, 80, 4096]
    blob = np.concatenate([hdr] + [i.detector_tensor.ravel() for i in ifos] + [i.vertex for i in ifos]
                          + [params.ravel()])
    raw = host_check("wave", blob)
    NC, DET, DS = int(raw[0]), int(raw[1]), int(raw[2])
    out = raw[4:].reshape(n, -1)
    assert out.shape[1] == NC + 2 * nf
    f = ifos[0].frequency_array
    worst = 0.0
    for i in range(0, n, 3):
        coef, ap = out[i, :NC], out[i, NC:].reshape(nf, 2)
        p = conv[i]
        pols = ocl.lal_binary_black_hole(f, *[p[k] for k in ocl.SOURCE_ARGS], **dict(
            waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0))
        for d, ifo in enumerate(ifos):
            ref = ifo.get_detector_response(pols, p)
            K = coef[DET + DS * d] + 1j * coef[DET + DS * d + 1]
            mine = K * ap[:, 0] * np.exp(-1j * np.pi * (ap[:, 1] + coef[DET + DS * d + 2] * f)) * ifo.frequency_mask
            worst = max(worst, np.max(np.abs(mine - ref)) / np.max(np.abs(ref)))
    assert worst < 1e-10, worst


def test_gmst_bit_exact(host_check):
    t = np.concatenate([[1126259642.413], np.random.default_rng(1).uniform(1.0e9, 1.4e9, 500)])
    got = host_check("gmst", t)
    ref = np.array([ocl.greenwich_mean_sidereal_time(x) for x in t])
    assert np.array_equal(got, ref)


def test_ln_i0(host_check):
    from scipy.special import i0e
    x = np.concatenate([np.linspace(-10, 10, 1001), np.logspace(-8, 10, 500)])
    got = host_check("lni0", x)
    ref = np.log(i0e(x)) + np.abs(x)
    assert np.max(np.abs(got - ref) / np.maximum(1e-3, np.abs(ref))) < 2e-12


def test_bicubic_spline_vs_scipy(host_check):
    from scipy.interpolate import RectBivariateSpline
    rng = np.random.default_rng(3)
    x = np.logspace(-5, 10, 80)
    y = np.logspace(-5, 10, 40)
    z = np.log1p(np.outer(x, 1 / np.sqrt(y))) + rng.normal(0, 0.01, (80, 40))
    spl = RectBivariateSpline(x, y, z, kx=3, ky=3, s=0)
    tx, ty, c = spl.tck
    xq = 10 ** rng.uniform(-5.5, 10.5, 4000)
    yq = 10 ** rng.uniform(-5.5, 10.5, 4000)
    xq[:3] = [x[0], x[-1], x[17]]
    yq[:3] = [y[0], y[-1], y[9]]
    blob = np.concatenate([[len(tx), len(ty), len(xq), x.min(), x.max(), y.min(), y.max()], tx, ty, c, xq, yq])
    got = host_check("bispev", blob)
    ref = spl(xq, yq, grid=False)
    bad = (xq < x.min()) | (xq > x.max()) | (yq < y.min()) | (yq > y.max())
    assert np.all(got[bad] == -np.inf)
    assert np.max(np.abs(got[~bad] - ref[~bad]) / np.maximum(1.0, np.abs(ref[~bad]))) < 1e-12


def test_taylorf2_prologue_and_bins_vs_oracle(host_check):
    """TaylorF2 + tides on the 128 s / 4096 Hz BNS grid (BASELINE.json configs[3])."""
    st = 1126259642.413 - 126.0
    ifos = [ocl.OracleInterferometer(n, 4096.0, 128.0, st) for n in ("H1", "L1", "V1")]
    rng = np.random.default_rng(11)
    n = 6
    mc = rng.uniform(1.15, 1.25, n)
    q = rng.uniform(0.5, 1.0, n)
    total = mc * (1 + q) ** 1.2 / q ** 0.6
    m1 = total / (1 + q)
    m2 = m1 * q
    params = np.zeros((n, 16))
    params[:, 0], params[:, 1] = m1, m2
    params[:, 2] = rng.uniform(-0.05, 0.05, n)
    params[:, 3] = rng.uniform(-0.05, 0.05, n)
    params[:, 4] = rng.uniform(10, 500, n)
    params[:, 5] = np.arccos(rng.uniform(-1, 1, n))
    params[:, 6] = rng.uniform(0, np.pi, n)
    params[:, 7] = rng.uniform(0, 2 * np.pi, n)
    params[:, 8] = rng.uniform(0, 2 * np.pi, n)
    params[:, 9] = np.arcsin(rng.uniform(-1, 1, n))
    params[:, 10] = rng.uniform(st + 125.9, st + 126.1, n)
    params[:, 12] = rng.uniform(0, 5000, n)
    params[:, 13] = rng.uniform(0, 5000, n)
    nf = len(ifos[0].frequency_array)
    k_lo, k_hi = int(20 * 128), nf - 1
    hdr = [n, 3, nf, 128.0, 4096.0, st, 1, 50.0, 20.0, 2048.0, k_lo, k_hi]
    blob = np.concatenate([hdr] + [i.detector_tensor.ravel() for i in ifos] + [i.vertex for i in ifos]
                          + [params.ravel()])
    raw = host_check("wave", blob)
    NC, DET, DS = int(raw[0]), int(raw[1]), int(raw[2])
    out = raw[4:].reshape(n, -1)
    f = ifos[0].frequency_array
    worst = 0.0
    for i in range(n):
        coef, ap = out[i, :NC], out[i, NC:].reshape(nf, 2)
        p = dict(mass_1=m1[i], mass_2=m2[i], luminosity_distance=params[i, 4], a_1=abs(params[i, 2]),
                 tilt_1=0.0 if params[i, 2] >= 0 else np.pi, phi_12=0.0, a_2=abs(params[i, 3]),
                 tilt_2=0.0 if params[i, 3] >= 0 else np.pi, phi_jl=0.0, theta_jn=params[i, 5], phase=params[i, 7],
                 lambda_1=params[i, 12], lambda_2=params[i, 13])
        pols = ocl.lal_binary_neutron_star(f, **p, waveform_approximant="TaylorF2", reference_frequency=50.0,
                                           minimum_frequency=20.0)
        ext = dict(ra=params[i, 8], dec=params[i, 9], psi=params[i, 6], geocent_time=params[i, 10])
        for d, ifo in enumerate(ifos):
            ref = ifo.get_detector_response(pols, ext)
            K = coef[DET + DS * d] + 1j * coef[DET + DS * d + 1]
            mine = K * ap[:, 0] * np.exp(-1j * np.pi * (ap[:, 1] + coef[DET + DS * d + 2] * f)) * ifo.frequency_mask
            worst = max(worst, np.max(np.abs(mine - ref)) / np.max(np.abs(ref)))
    # phases reach ~1e6 rad (126 s time shift at 2 kHz): double rounding of the argument alone is ~1e-10
    assert worst < 5e-9, worst
