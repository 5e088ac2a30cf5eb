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
NC = 72


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
    out = host_check("wave", blob).reshape(n, -1)
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
            K = coef[NC - 16 + 4 * d] + 1j * coef[NC - 16 + 4 * d + 1]
            mine = K * ap[:, 0] * np.exp(-1j * np.pi * (ap[:, 1] + coef[NC - 16 + 4 * d + 2] * f)) * ifo.frequency_mask
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
    assert np.max(np.abs(got - ref) / np.maximum(1e-3, np.abs(ref))) < 1e-13


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
