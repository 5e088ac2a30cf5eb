"""BUILD TOOL - golden vectors for the device-resident sampling front end (unit cube -> sampled parameters ->
source-model parameters -> parameter rows) from the UNMODIFIED reference:
  bilby.core.prior.PriorDict.rescale (core/prior/dict.py:647-666) with the analytic priors of core/prior/analytical.py,
  bilby.gw.conversion.convert_to_lal_binary_black_hole_parameters (:182-283) / ..._neutron_star_parameters (:286-348).

    PYTHONPATH=oracle/standins:/root/reference python oracle/tools/make_golden_prior.py

Writes tests/golden/prior_transform.npz: per case the prior table (kind, key, a, b, c), the fixed parameters, the
unit-cube draws, the reference's sampled parameters and the 14 row columns the kernels read."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

import bilby  # noqa: E402
from bilby.core.prior import Uniform, PowerLaw, LogUniform, Sine, Cosine, Gaussian, DeltaFunction  # noqa: E402
from bilby.gw import conversion  # noqa: E402

bilby.core.utils.logger.setLevel("ERROR")

ROW_KEYS = ["mass_1", "mass_2", "chi_1", "chi_2", "luminosity_distance", "theta_jn", "psi", "phase", "ra", "dec",
            "geocent_time", "time_jitter", "lambda_1", "lambda_2"]
T0 = 1126259642.413

CASES = {
    # the usual BBH set-up: chirp mass and mass ratio, aligned spins, power-law distance
    "bbh_mc_q": (False, dict(
        chirp_mass=Uniform(25.0, 35.0), mass_ratio=Uniform(0.125, 1.0), chi_1=Uniform(-0.9, 0.9),
        chi_2=Uniform(-0.9, 0.9), luminosity_distance=PowerLaw(2.0, 100.0, 5000.0), theta_jn=Sine(),
        psi=Uniform(0.0, np.pi), phase=Uniform(0.0, 2 * np.pi), ra=Uniform(0.0, 2 * np.pi), dec=Cosine(),
        geocent_time=Uniform(T0 - 0.1, T0 + 0.1))),
    # total mass + symmetric mass ratio, spin magnitude with cos tilt = +-1, cos theta_jn, delta_phase, log-uniform
    # distance, Gaussian spin, marginalised time (geocent_time fixed to the segment start, time_jitter sampled)
    "bbh_mtot_eta": (False, dict(
        total_mass=Uniform(40.0, 120.0), symmetric_mass_ratio=Uniform(0.1, 0.25), a_1=Uniform(0.0, 0.8),
        cos_tilt_1=DeltaFunction(-1.0), chi_2=Gaussian(0.0, 0.2), luminosity_distance=LogUniform(50.0, 4000.0),
        cos_theta_jn=Uniform(-1.0, 1.0), psi=Uniform(0.0, np.pi), delta_phase=Uniform(0.0, 2 * np.pi),
        ra=Uniform(0.0, 2 * np.pi), dec=Cosine(-0.5, 1.0), geocent_time=DeltaFunction(T0 - 2.0),
        time_jitter=Uniform(-1.0 / 4096, 1.0 / 4096))),
    # component mass + total mass, chirp mass + total mass variants
    "bbh_m1_mtot": (False, dict(
        mass_1=Uniform(20.0, 60.0), total_mass=Uniform(70.0, 90.0), chi_1=DeltaFunction(0.3), a_2=DeltaFunction(0.0),
        luminosity_distance=PowerLaw(-0.5, 10.0, 2000.0), theta_jn=Sine(0.2, 2.5), psi=DeltaFunction(0.7),
        phase=Uniform(0.0, 2 * np.pi), ra=DeltaFunction(1.375), dec=DeltaFunction(-1.2108),
        geocent_time=Gaussian(T0, 0.01))),
    "bbh_mc_mtot": (False, dict(
        chirp_mass=Uniform(20.0, 21.0), total_mass=Uniform(50.0, 60.0), chi_1=Uniform(-0.5, 0.5), chi_2=DeltaFunction(0.0),
        luminosity_distance=Uniform(100.0, 1000.0), theta_jn=Sine(), psi=Uniform(0.0, np.pi),
        phase=Uniform(0.0, 2 * np.pi), ra=Uniform(0.0, 2 * np.pi), dec=Cosine(), geocent_time=Uniform(T0 - 0.1, T0 + 0.1))),
    # BNS: lambda_tilde + delta_lambda_tilde
    "bns_lt_dlt": (True, dict(
        chirp_mass=Uniform(1.18, 1.22), mass_ratio=Uniform(0.5, 1.0), chi_1=Uniform(-0.05, 0.05),
        chi_2=Uniform(-0.05, 0.05), luminosity_distance=PowerLaw(2.0, 10.0, 500.0), theta_jn=Sine(),
        psi=Uniform(0.0, np.pi), phase=Uniform(0.0, 2 * np.pi), ra=Uniform(0.0, 2 * np.pi), dec=Cosine(),
        geocent_time=Uniform(T0 - 0.1, T0 + 0.1), lambda_tilde=Uniform(0.0, 1000.0),
        delta_lambda_tilde=Uniform(-500.0, 500.0))),
    # BNS: lambda_tilde alone; mass_2 + mass_ratio
    "bns_lt": (True, dict(
        mass_2=Uniform(1.0, 1.4), mass_ratio=Uniform(0.6, 1.0), chi_1=DeltaFunction(0.0), chi_2=DeltaFunction(0.0),
        luminosity_distance=Uniform(10.0, 300.0), theta_jn=Sine(), psi=Uniform(0.0, np.pi),
        phase=Uniform(0.0, 2 * np.pi), ra=Uniform(0.0, 2 * np.pi), dec=Cosine(),
        geocent_time=Uniform(T0 - 0.1, T0 + 0.1), lambda_tilde=Uniform(10.0, 3000.0))),
    # BNS: component tides, and lambda_1 alone (lambda_2 follows the mass ratio)
    "bns_l1_l2": (True, dict(
        mass_1=Uniform(1.3, 1.8), mass_2=Uniform(1.0, 1.3), chi_1=Uniform(-0.05, 0.05), chi_2=Uniform(-0.05, 0.05),
        luminosity_distance=PowerLaw(2.0, 10.0, 500.0), theta_jn=Sine(), psi=Uniform(0.0, np.pi),
        phase=Uniform(0.0, 2 * np.pi), ra=Uniform(0.0, 2 * np.pi), dec=Cosine(),
        geocent_time=Uniform(T0 - 0.1, T0 + 0.1), lambda_1=Uniform(0.0, 5000.0), lambda_2=Uniform(0.0, 5000.0))),
    "bns_l1": (True, dict(
        mass_1=Uniform(1.3, 1.8), mass_ratio=Uniform(0.7, 1.0), chi_1=DeltaFunction(0.01), chi_2=DeltaFunction(-0.02),
        luminosity_distance=DeltaFunction(100.0), theta_jn=Sine(), psi=Uniform(0.0, np.pi),
        phase=Uniform(0.0, 2 * np.pi), ra=Uniform(0.0, 2 * np.pi), dec=Cosine(),
        geocent_time=Uniform(T0 - 0.1, T0 + 0.1), lambda_1=Uniform(0.0, 3000.0))),
}

KIND = {"DeltaFunction": 0, "Uniform": 1, "PowerLaw": 2, "LogUniform": 2, "Sine": 3, "Cosine": 4, "Gaussian": 5}


def spec(p):
    name = type(p).__name__
    if name == "DeltaFunction":
        return KIND[name], float(p.peak), 0.0, 0.0
    if name in ("PowerLaw", "LogUniform"):
        return KIND[name], float(p.minimum), float(p.maximum), float(p.alpha)
    if name == "Gaussian":
        return KIND[name], float(p.mu), float(p.sigma), 0.0
    return KIND[name], float(p.minimum), float(p.maximum), 0.0


def main():
    n = 257
    res = dict(row_keys=np.array(ROW_KEYS), case_names=np.array(list(CASES)))
    for ci, (name, (bns, pri)) in enumerate(CASES.items()):
        priors = bilby.core.prior.PriorDict(dict(pri))
        keys = [k for k in priors if not isinstance(priors[k], DeltaFunction)]
        fixed = {k: float(priors[k].peak) for k in priors if isinstance(priors[k], DeltaFunction)}
        rng = np.random.default_rng(4000 + ci)
        u = rng.uniform(0.0, 1.0, (n, len(keys)))
        u[0] = 0.0            # the ends of the unit interval
        u[1] = 1.0
        if "Gaussian" in [type(priors[k]).__name__ for k in keys]:
            u[:2] = np.clip(u[:2], 1e-12, 1 - 1e-12)
        theta = np.array([priors.rescale(keys, u[i]) for i in range(n)], dtype=float)
        rows = np.zeros((n, len(ROW_KEYS)))
        convert = (conversion.convert_to_lal_binary_neutron_star_parameters if bns
                   else conversion.convert_to_lal_binary_black_hole_parameters)
        for i in range(n):
            p = dict(fixed)
            p.update({k: float(theta[i, j]) for j, k in enumerate(keys)})
            c, _ = convert(p)
            for slot, rk in enumerate(ROW_KEYS):
                if rk in ("chi_1", "chi_2"):
                    idx = rk[-1]
                    rows[i, slot] = c.get(f"a_{idx}", 0.0) * np.cos(c.get(f"tilt_{idx}", 0.0))
                else:
                    rows[i, slot] = c.get(rk, 0.0)
        res[f"{name}_bns"] = np.array(bns)
        res[f"{name}_keys"] = np.array(keys)
        res[f"{name}_spec"] = np.array([spec(priors[k]) for k in keys], dtype=float)
        res[f"{name}_fixed_keys"] = np.array(list(fixed))
        res[f"{name}_fixed_values"] = np.array([fixed[k] for k in fixed], dtype=float)
        res[f"{name}_unit"] = u
        res[f"{name}_theta"] = theta
        res[f"{name}_rows"] = rows
        print(name, keys, rows[2, :6])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "prior_transform.npz"), **res)


if __name__ == "__main__":
    main()
