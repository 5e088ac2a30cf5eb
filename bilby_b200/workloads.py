"""Synthetic benchmark workloads: injection parameters and prior draws.
Host-side input generation only; shared by bench.py's GPU and CPU arms so both see identical inputs."""
import numpy as np

# examples/gw_examples/injection_examples/fast_tutorial.py:31-47 with aligned spins
INJECTION = dict(mass_1=36.0, mass_2=29.0, chi_1=0.4, chi_2=0.3, luminosity_distance=2000.0,
                 theta_jn=0.4, psi=2.659, phase=1.3, geocent_time=1126259642.413, ra=1.375, dec=-1.2108)


def draw_bbh_prior(n, rng, t_inj=INJECTION["geocent_time"]):
    """chirp_mass U(25,35), mass_ratio U(0.125,1) with m1,m2 in [5,100], chi_i U(-0.99,0.99),
    d_L PowerLaw(2,100,5000), cos theta_jn U(-1,1), psi U(0,pi), phase U(0,2pi), ra U(0,2pi),
    sin dec U(-1,1), t_c U(t_inj +- 0.1)."""
    out = {}
    mc = np.empty(n)
    q = np.empty(n)
    filled = 0
    while filled < n:
        m = rng.uniform(25, 35, n)
        qq = rng.uniform(0.125, 1, n)
        total = m * (1 + qq) ** 1.2 / qq ** 0.6
        m1 = total / (1 + qq)
        m2 = m1 * qq
        ok = (m1 >= 5) & (m1 <= 100) & (m2 >= 5) & (m2 <= 100)
        k = min(n - filled, int(ok.sum()))
        mc[filled:filled + k] = m[ok][:k]
        q[filled:filled + k] = qq[ok][:k]
        filled += k
    out["chirp_mass"] = mc
    out["mass_ratio"] = q
    out["chi_1"] = rng.uniform(-0.99, 0.99, n)
    out["chi_2"] = rng.uniform(-0.99, 0.99, n)
    u = rng.uniform(0, 1, n)
    out["luminosity_distance"] = (100.0 ** 3 + u * (5000.0 ** 3 - 100.0 ** 3)) ** (1 / 3)
    out["theta_jn"] = np.arccos(rng.uniform(-1, 1, n))
    out["psi"] = rng.uniform(0, np.pi, n)
    out["phase"] = rng.uniform(0, 2 * np.pi, n)
    out["ra"] = rng.uniform(0, 2 * np.pi, n)
    out["dec"] = np.arcsin(rng.uniform(-1, 1, n))
    out["geocent_time"] = rng.uniform(t_inj - 0.1, t_inj + 0.1, n)
    return out
