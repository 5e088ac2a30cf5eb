"""BUILD TOOL - golden vectors for the calibration-marginalised likelihood from the UNMODIFIED reference
(base.py:305-346, 860-877, 1037-1051; calibration.py:503-591):

    PYTHONPATH=oracle/standins:/root/reference python oracle/tools/make_golden_calmarg.py

Writes tests/golden/calmarg_4s_H1L1V1.npz: the spline-node draws of the response curves the reference generated
(40 curves per detector, 10 nodes, Gaussian priors; the curves themselves are rebuilt from them by the consumers) and
log_likelihood_ratio for {calibration, calibration + phase, calibration + distance + phase} marginalisation.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import bilby  # noqa: E402
from bilby.core.prior import Uniform, PowerLaw, PriorDict, Gaussian  # noqa: E402
from bilby.core.utils import random as brandom  # noqa: E402
from oracle import cbc_likelihood as ocl  # noqa: E402
from make_golden import build  # noqa: E402

bilby.core.utils.logger.setLevel("ERROR")
N_CURVES, N_POINTS = 40, 10


def main():
    names = ["H1", "L1", "V1"]
    inj, start_time, wfg, ifos = build(4.0, 2048.0, names, noise_seed=88170235)
    g = np.load(os.path.join(ROOT, "tests", "golden", "recon_4s_H1L1V1.npz"))
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_") and k != "param_time_jitter"}
    n = len(draws["chirp_mass"])
    res = dict(start_time=start_time, duration=4.0, sampling_frequency=2048.0, detectors=np.array(names),
               noise_seed=88170235, n_curves=N_CURVES, n_points=N_POINTS)
    for k in draws:
        res["param_" + k] = draws[k]

    t_inj = inj["geocent_time"]
    jitter = np.random.default_rng(7).uniform(-1 / 2048.0, 1 / 2048.0, n)
    res["param_time_jitter"] = jitter

    def priors(phase=False, distance=False, time=False):
        pri = {}
        if time:
            pri["geocent_time"] = Uniform(t_inj - 0.1, t_inj + 0.1, "geocent_time")
        for name in names:
            for i in range(N_POINTS):
                pri[f"recalib_{name}_amplitude_{i}"] = Gaussian(0.0, 0.05, f"recalib_{name}_amplitude_{i}")
                pri[f"recalib_{name}_phase_{i}"] = Gaussian(0.0, 0.05, f"recalib_{name}_phase_{i}")
        if phase:
            pri["phase"] = Uniform(0, 2 * np.pi, "phase")
        if distance:
            pri["luminosity_distance"] = PowerLaw(2, 100.0, 5000.0, "luminosity_distance")
        return PriorDict(pri)

    first = None
    for mode, kw in (("cal", {}), ("cal_phase", dict(phase_marginalization=True)),
                     ("cal_distance_phase", dict(phase_marginalization=True, distance_marginalization=True,
                                                 distance_marginalization_lookup_table="/tmp/golden_dp_lookup.npz")),
                     # time + calibration (base.py:305-323): one FFT per response curve; with distance marginalisation
                     # the reference itself raises a shape mismatch, so only these two exist
                     ("cal_time", dict(time_marginalization=True, jitter_time=True)),
                     ("cal_time_phase", dict(time_marginalization=True, jitter_time=True, phase_marginalization=True))):
        brandom.seed(424242)
        for ifo in ifos:     # build_calibration_lookup resets the model to the identity (calibration.py:552)
            ifo.calibration_model = bilby.gw.detector.calibration.CubicSpline(
                prefix=f"recalib_{ifo.name}_", minimum_frequency=ifo.minimum_frequency,
                maximum_frequency=ifo.maximum_frequency, n_points=N_POINTS)
        like = bilby.gw.likelihood.GravitationalWaveTransient(
            ifos, wfg, calibration_marginalization=True, number_of_response_curves=N_CURVES,
            priors=priors(kw.get("phase_marginalization", False), kw.get("distance_marginalization", False),
                          kw.get("time_marginalization", False)), **kw)
        if first is None:
            first = like
            for name in names:
                frame = like.calibration_parameter_draws[name]
                amp = np.stack([frame[f"recalib_{name}_amplitude_{i}"].to_numpy() for i in range(N_POINTS)], axis=1)
                pha = np.stack([frame[f"recalib_{name}_phase_{i}"].to_numpy() for i in range(N_POINTS)], axis=1)
                res[f"curve_nodes_{name}"] = np.stack([amp, pha], axis=1)          # [n_curves, 2, n_points]
                # three curves kept whole to pin the curve builders
                res[f"curve_samples_{name}"] = like.calibration_draws[name][[0, 17, 39]][:, ::16]
        else:
            for name in names:
                assert np.array_equal(like.calibration_draws[name], first.calibration_draws[name])
        vals = np.zeros(n)
        for i in range(n):
            p = {k: float(draws[k][i]) for k in draws}
            for name in names:
                for j in range(N_POINTS):
                    p[f"recalib_{name}_amplitude_{j}"] = 0.0
                    p[f"recalib_{name}_phase_{j}"] = 0.0
            if kw.get("time_marginalization"):
                p["geocent_time"] = float(start_time)
                p["time_jitter"] = float(jitter[i])
            vals[i] = like.log_likelihood_ratio(p)
        res["lnl_" + mode] = vals
        print(mode, vals[:4], vals[12])
        if kw.get("time_marginalization"):
            continue          # the reconstruction below is pinned for the modes without time marginalisation
        # marginalised-parameter reconstruction (base.py:502-578): recalib_index by rng.choice over the response
        # curves' posterior, then distance / phase with the chosen curve applied (base.py:289-290); the unit-interval
        # draws of the reference's generator are replayed and stored
        n_u = 1 + bool(kw.get("distance_marginalization")) + bool(kw.get("phase_marginalization"))
        out = np.zeros((n, 3))
        uni = np.full((n, 3), np.nan)
        for i in range(n):
            p = {k: float(draws[k][i]) for k in draws}
            for name in names:
                for j in range(N_POINTS):
                    p[f"recalib_{name}_amplitude_{j}"] = 0.0
                    p[f"recalib_{name}_phase_{j}"] = 0.0
            brandom.seed(3000 + i)
            new = like.generate_posterior_sample_from_marginalized_likelihood(p)
            out[i] = [new["recalib_index"], new["luminosity_distance"], new["phase"]]
            replay = np.random.default_rng(3000 + i)
            drawn = [replay.uniform(0, 1) for _ in range(n_u)]
            uni[i, 0] = drawn[0]
            j = 1
            for c, key in ((1, "distance_marginalization"), (2, "phase_marginalization")):
                if kw.get(key):
                    uni[i, c] = drawn[j]
                    j += 1
        res["recon_" + mode] = out
        res["uniforms_" + mode] = uni
        print(mode, "recon", out[:3])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "calmarg_4s_H1L1V1.npz"), **res)


if __name__ == "__main__":
    main()
