"""BUILD TOOL - golden vectors for the marginalised-parameter reconstruction and the per-detector SNRs, from the
UNMODIFIED reference (bilby imported from /root/reference with the stand-ins of oracle/standins):

    PYTHONPATH=oracle/standins:/root/reference python oracle/tools/make_golden_recon.py

Writes tests/golden/recon_4s_H1L1V1.npz:
  * GravitationalWaveTransient.generate_posterior_sample_from_marginalized_likelihood (base.py:502-773) for six
    marginalisation modes; the reference's global generator is re-seeded with 1000 + i before sample i, and the
    unit-interval draws it makes are stored next to the results (numpy default_rng(seed).uniform(0, 1) replayed);
  * bilby.gw.conversion.compute_snrs (conversion.py:2215-2288): complex matched-filter SNR and optimal SNR per
    detector.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import bilby  # noqa: E402
from bilby.core.prior import Uniform, PowerLaw, PriorDict  # noqa: E402
from bilby.core.utils import random as brandom  # noqa: E402
from oracle import cbc_likelihood as ocl  # noqa: E402
from make_golden import build  # noqa: E402

bilby.core.utils.logger.setLevel("ERROR")

MODES = {
    "phase": dict(phase_marginalization=True),
    "distance": dict(distance_marginalization=True),
    "distance_phase": dict(distance_marginalization=True, phase_marginalization=True),
    "time": dict(time_marginalization=True),
    "time_phase": dict(time_marginalization=True, phase_marginalization=True),
    "time_distance_phase": dict(time_marginalization=True, distance_marginalization=True, phase_marginalization=True),
}


def main():
    names = ["H1", "L1", "V1"]
    inj, start_time, wfg, ifos = build(4.0, 2048.0, names, noise_seed=88170235)
    n0 = 12
    draws = ocl.draw_bbh_prior(n0, np.random.default_rng(20261017))
    # posterior-like points: the injection and small perturbations of it (peaked posteriors)
    rng = np.random.default_rng(5)
    base = dict(chirp_mass=28.0956, mass_ratio=29.0 / 36.0, chi_1=0.4, chi_2=0.3, luminosity_distance=2000.0,
                theta_jn=0.4, psi=2.659, phase=1.3, ra=1.375, dec=-1.2108, geocent_time=inj["geocent_time"])
    extra = [dict(base)]
    for _ in range(7):
        e = dict(base)
        e["chirp_mass"] += rng.normal(0, 0.05)
        e["mass_ratio"] = min(1.0, e["mass_ratio"] + rng.normal(0, 0.02))
        e["geocent_time"] += rng.normal(0, 2e-3)
        e["luminosity_distance"] *= np.exp(rng.normal(0, 0.3))
        e["theta_jn"] += rng.normal(0, 0.2)
        e["ra"] += rng.normal(0, 0.05)
        e["dec"] += rng.normal(0, 0.05)
        extra.append(e)
    for k in draws:
        draws[k] = np.concatenate([draws[k], [e[k] for e in extra]])
    n = len(draws["chirp_mass"])
    draws["time_jitter"] = np.random.default_rng(7).uniform(-1 / 2048.0, 1 / 2048.0, n)
    res = dict(start_time=start_time, duration=4.0, sampling_frequency=2048.0, detectors=np.array(names),
               noise_seed=88170235)
    for k in draws:
        res["param_" + k] = draws[k]
    t_inj = inj["geocent_time"]
    for mode, kw in MODES.items():
        pri = {}
        if kw.get("phase_marginalization"):
            pri["phase"] = Uniform(0, 2 * np.pi, "phase")
        if kw.get("distance_marginalization"):
            pri["luminosity_distance"] = PowerLaw(2, 100.0, 5000.0, "luminosity_distance")
            kw = dict(kw, distance_marginalization_lookup_table="/tmp/golden_%s_lookup.npz"
                      % ("dp" if kw.get("phase_marginalization") else "d"))
        if kw.get("time_marginalization"):
            pri["geocent_time"] = Uniform(t_inj - 0.1, t_inj + 0.1, "geocent_time")
            kw = dict(kw, jitter_time=True)
        like = bilby.gw.likelihood.GravitationalWaveTransient(ifos, wfg, priors=PriorDict(pri), **kw)
        n_u = sum(bool(kw.get(k)) for k in ("time_marginalization", "distance_marginalization", "phase_marginalization"))
        out = np.zeros((n, 3))
        uni = np.full((n, 3), np.nan)
        for i in range(n):
            p = {k: float(draws[k][i]) for k in draws}
            if not kw.get("time_marginalization"):
                p.pop("time_jitter")
            brandom.seed(1000 + i)
            new = like.generate_posterior_sample_from_marginalized_likelihood(p)
            out[i] = [new["geocent_time"], new["luminosity_distance"], new["phase"]]
            replay = np.random.default_rng(1000 + i)
            drawn = [replay.uniform(0, 1) for _ in range(n_u)]
            j = 0
            for c, key in enumerate(("time_marginalization", "distance_marginalization", "phase_marginalization")):
                if kw.get(key):
                    uni[i, c] = drawn[j]
                    j += 1
        res["recon_" + mode] = out
        res["uniforms_" + mode] = uni
        if kw.get("distance_marginalization"):
            res["ref_dist"] = like._ref_dist
        print(mode, out[n0])
    # per-detector SNRs (plain likelihood)
    like = bilby.gw.likelihood.GravitationalWaveTransient(ifos, wfg)
    mf = np.zeros((n, len(names)), dtype=complex)
    opt = np.zeros((n, len(names)))
    for i in range(n):
        p = {k: float(draws[k][i]) for k in draws if k != "time_jitter"}
        bilby.gw.conversion.compute_snrs(p, like)
        for j, name in enumerate(names):
            mf[i, j] = p[f"{name}_matched_filter_snr"]
            opt[i, j] = p[f"{name}_optimal_snr"]
    res["matched_filter_snr"] = mf
    res["optimal_snr"] = opt
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "recon_4s_H1L1V1.npz"), **res)
    print("written", n, "samples")
    main_8s(draws, n0)


def main_8s(draws, n0):
    """8 s, zero-noise H1+L1, time + phase: duration != 4 s exposes the normalisation of the reference's 16384 Hz
    transform (base.py:626 has no 4/T factor, unlike calculate_snrs :325-330)."""
    names = ["H1", "L1"]
    inj, start_time, wfg, ifos = build(8.0, 2048.0, names, noise_seed=None)
    idx = [0, 3, n0, n0 + 1, n0 + 4, n0 + 6]
    t_inj = inj["geocent_time"]
    pri = dict(geocent_time=Uniform(t_inj - 0.1, t_inj + 0.1, "geocent_time"), phase=Uniform(0, 2 * np.pi, "phase"))
    like = bilby.gw.likelihood.GravitationalWaveTransient(ifos, wfg, priors=PriorDict(pri), time_marginalization=True,
                                                          phase_marginalization=True, jitter_time=True)
    res = dict(start_time=start_time, duration=8.0, sampling_frequency=2048.0, detectors=np.array(names))
    for k in draws:
        res["param_" + k] = draws[k][idx]
    out = np.zeros((len(idx), 3))
    uni = np.full((len(idx), 3), np.nan)
    for m, i in enumerate(idx):
        p = {k: float(draws[k][i]) for k in draws}
        brandom.seed(2000 + i)
        new = like.generate_posterior_sample_from_marginalized_likelihood(p)
        out[m] = [new["geocent_time"], new["luminosity_distance"], new["phase"]]
        replay = np.random.default_rng(2000 + i)
        uni[m, 0] = replay.uniform(0, 1)
        uni[m, 2] = replay.uniform(0, 1)
    res["recon_time_phase"] = out
    res["uniforms_time_phase"] = uni
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "recon_8s_zero_H1L1.npz"), **res)
    print("8 s:", out)


if __name__ == "__main__":
    main()
