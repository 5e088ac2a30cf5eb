"""Sampled -> source-model parameters, vectorised (numpy arrays or torch tensors).

Mirrors the subset of bilby/gw/conversion.py that is on the hot path:
``convert_to_lal_binary_black_hole_parameters`` :182-283, ``generate_component_masses`` :1826-1985,
mass helpers :849-870, :905-937, :969-989, tidal maps :1187-1264,
``convert_to_lal_binary_neutron_star_parameters`` :286-348 (EOS branches out of scope).
Cosmological (redshift / *_source) branches need astropy in the reference and are out of scope.
"""
import numpy as np


def _xp(parameters):
    try:
        import torch
        if any(isinstance(v, torch.Tensor) for v in parameters.values()):
            return torch
    except ImportError:  # pragma: no cover
        pass
    return np


def symmetric_mass_ratio_to_mass_ratio(symmetric_mass_ratio):
    temp = (1 / symmetric_mass_ratio / 2 - 1)
    return temp - (temp ** 2 - 1) ** 0.5


def chirp_mass_and_mass_ratio_to_total_mass(chirp_mass, mass_ratio):
    return chirp_mass * (1 + mass_ratio) ** 1.2 / mass_ratio ** 0.6


def total_mass_and_mass_ratio_to_component_masses(mass_ratio, total_mass):
    mass_1 = total_mass / (1 + mass_ratio)
    mass_2 = mass_1 * mass_ratio
    return mass_1, mass_2


def chirp_mass_and_total_mass_to_symmetric_mass_ratio(chirp_mass, total_mass):
    return (chirp_mass / total_mass) ** (5 / 3)


def component_masses_to_symmetric_mass_ratio(mass_1, mass_2):
    eta = (mass_1 * mass_2) / (mass_1 + mass_2) ** 2
    if isinstance(eta, np.ndarray) or np.isscalar(eta):
        return np.minimum(eta, 1 / 4)
    return eta.clamp(max=0.25)


def generate_component_masses(sample, require_add=False):
    """conversion.py:1826-1985 (non-source keys; mass_x + chirp_mass cubic inversions not supported)."""
    out = dict(sample)
    keys = sample.keys()
    if "mass_1" in keys:
        if "mass_2" in keys:
            return out
        if "total_mass" in keys:
            out["mass_2"] = out["total_mass"] - out["mass_1"]
            return out
        if "mass_ratio" not in keys:
            if "symmetric_mass_ratio" in keys:
                out["mass_ratio"] = symmetric_mass_ratio_to_mass_ratio(out["symmetric_mass_ratio"])
            elif require_add:
                raise KeyError("Insufficient mass parameters in input dictionary")
            else:
                return out
        out["mass_2"] = out["mass_ratio"] * out["mass_1"]
        return out
    if "mass_2" in keys:
        if "total_mass" in keys:
            out["mass_1"] = out["total_mass"] - out["mass_2"]
            return out
        if "mass_ratio" not in keys:
            if "symmetric_mass_ratio" in keys:
                out["mass_ratio"] = symmetric_mass_ratio_to_mass_ratio(out["symmetric_mass_ratio"])
            elif require_add:
                raise KeyError("Insufficient mass parameters in input dictionary")
            else:
                return out
        out["mass_1"] = 1 / out["mass_ratio"] * out["mass_2"]
        return out
    if "total_mass" in keys:
        if "mass_ratio" in keys:
            pass
        elif "symmetric_mass_ratio" in keys:
            out["mass_ratio"] = symmetric_mass_ratio_to_mass_ratio(out["symmetric_mass_ratio"])
        elif "chirp_mass" in keys:
            out["symmetric_mass_ratio"] = chirp_mass_and_total_mass_to_symmetric_mass_ratio(
                out["chirp_mass"], out["total_mass"])
            out["mass_ratio"] = symmetric_mass_ratio_to_mass_ratio(out["symmetric_mass_ratio"])
        elif require_add:
            raise KeyError("Insufficient mass parameters in input dictionary")
        else:
            return out
    elif "chirp_mass" in keys:
        if "mass_ratio" in keys:
            pass
        elif "symmetric_mass_ratio" in keys:
            out["mass_ratio"] = symmetric_mass_ratio_to_mass_ratio(sample["symmetric_mass_ratio"])
        elif require_add:
            raise KeyError("Insufficient mass parameters in input dictionary")
        else:
            return out
        out["total_mass"] = chirp_mass_and_mass_ratio_to_total_mass(out["chirp_mass"], out["mass_ratio"])
    if "total_mass" not in out or "mass_ratio" not in out:
        if require_add:
            raise KeyError("Insufficient mass parameters in input dictionary")
        return out
    out["mass_1"], out["mass_2"] = total_mass_and_mass_ratio_to_component_masses(
        total_mass=out["total_mass"], mass_ratio=out["mass_ratio"])
    return out


def convert_to_lal_binary_black_hole_parameters(parameters):
    """conversion.py:182-283.  Returns (converted_parameters, added_keys) like the reference."""
    converted = dict(parameters)
    original_keys = list(converted.keys())
    xp = _xp(parameters)
    for key in original_keys:
        if key in ("redshift", "comoving_distance") or key.endswith("_source"):
            raise NotImplementedError(
                "cosmological / source-frame parameters need astropy in the reference and are out of scope")
    converted = generate_component_masses(converted, require_add=False)
    for idx in ("1", "2"):
        key = f"chi_{idx}"
        if key in original_keys:
            if f"chi_{idx}_in_plane" in original_keys:
                converted[f"a_{idx}"] = (converted[key] ** 2 + converted[f"chi_{idx}_in_plane"] ** 2) ** 0.5
                converted[f"cos_tilt_{idx}"] = converted[key] / converted[f"a_{idx}"]
            elif f"a_{idx}" not in original_keys:
                converted[f"a_{idx}"] = abs(converted[key])
                converted[f"cos_tilt_{idx}"] = xp.sign(xp.as_tensor(converted[key]) if xp is not np
                                                       else np.asarray(converted[key]))
            else:
                a = converted[f"a_{idx}"]
                with np.errstate(invalid="ignore", divide="ignore"):
                    ct = converted[key] / a
                if xp is np:
                    ct = np.where(np.asarray(a) == 0, 1.0, ct)
                else:
                    ct = xp.where(a == 0, xp.ones_like(ct), ct)
                converted[f"cos_tilt_{idx}"] = ct
    for key in ("phi_jl", "phi_12"):
        if key not in converted:
            converted[key] = 0.0
    for angle in ("tilt_1", "tilt_2", "theta_jn"):
        cos_angle = "cos_" + angle
        if cos_angle in converted:
            v = converted[cos_angle]
            converted[angle] = xp.arccos(v if xp is np else xp.as_tensor(v))
    if "delta_phase" in converted:
        tj = converted["theta_jn"]
        converted["phase"] = xp.remainder(
            converted["delta_phase"] - xp.sign(xp.cos(tj if xp is np else xp.as_tensor(tj))) * converted["psi"],
            2 * np.pi) if xp is not np else np.mod(
            converted["delta_phase"] - np.sign(np.cos(tj)) * converted["psi"], 2 * np.pi)
    added_keys = [key for key in converted if key not in original_keys]
    return converted, added_keys


def lambda_tilde_delta_lambda_tilde_to_lambda_1_lambda_2(lambda_tilde, delta_lambda_tilde, mass_1, mass_2):
    """conversion.py:1187-1231."""
    eta = component_masses_to_symmetric_mass_ratio(mass_1, mass_2)
    c1 = (1 + 7 * eta - 31 * eta ** 2)
    c2 = (1 - 4 * eta) ** 0.5 * (1 + 9 * eta - 11 * eta ** 2)
    c3 = (1 - 4 * eta) ** 0.5 * (1 - 13272 / 1319 * eta + 8944 / 1319 * eta ** 2)
    c4 = (1 - 15910 / 1319 * eta + 32850 / 1319 * eta ** 2 + 3380 / 1319 * eta ** 3)
    lambda_1 = ((13 * lambda_tilde / 8 * (c3 - c4) - 2 * delta_lambda_tilde * (c1 - c2))
                / ((c1 + c2) * (c3 - c4) - (c1 - c2) * (c3 + c4)))
    lambda_2 = ((13 * lambda_tilde / 8 * (c3 + c4) - 2 * delta_lambda_tilde * (c1 + c2))
                / ((c1 - c2) * (c3 + c4) - (c1 + c2) * (c3 - c4)))
    return lambda_1, lambda_2


def lambda_tilde_to_lambda_1_lambda_2(lambda_tilde, mass_1, mass_2):
    """conversion.py:1234-1264."""
    eta = component_masses_to_symmetric_mass_ratio(mass_1, mass_2)
    q = mass_2 / mass_1
    lambda_1 = 13 / 8 * lambda_tilde / (
        (1 + 7 * eta - 31 * eta ** 2) * (1 + q ** -5)
        + (1 - 4 * eta) ** 0.5 * (1 + 9 * eta - 11 * eta ** 2) * (1 - q ** -5))
    lambda_2 = lambda_1 / q ** 5
    return lambda_1, lambda_2


def convert_to_lal_binary_neutron_star_parameters(parameters):
    """conversion.py:286-348 (tidal branches; equation-of-state branches :349-557 out of scope)."""
    converted = dict(parameters)
    original_keys = list(converted.keys())
    converted, added_keys = convert_to_lal_binary_black_hole_parameters(converted)
    if any(k.startswith("eos_") for k in converted) or "lambda_symmetric" in converted:
        raise NotImplementedError("equation-of-state / lambda_symmetric parameterisations are out of scope")
    if not any(k in converted for k in ("lambda_1", "lambda_2", "lambda_tilde", "delta_lambda_tilde")):
        converted["lambda_1"] = 0
        converted["lambda_2"] = 0
        return converted, added_keys + ["lambda_1", "lambda_2"]
    if "delta_lambda_tilde" in converted:
        converted["lambda_1"], converted["lambda_2"] = lambda_tilde_delta_lambda_tilde_to_lambda_1_lambda_2(
            converted["lambda_tilde"], parameters["delta_lambda_tilde"], converted["mass_1"], converted["mass_2"])
    elif "lambda_tilde" in converted:
        converted["lambda_1"], converted["lambda_2"] = lambda_tilde_to_lambda_1_lambda_2(
            converted["lambda_tilde"], converted["mass_1"], converted["mass_2"])
    if "lambda_2" not in converted and "lambda_1" in converted:
        converted["lambda_2"] = converted["lambda_1"] * converted["mass_1"] ** 5 / converted["mass_2"] ** 5
    elif "lambda_2" in converted and converted["lambda_2"] is None:
        converted["lambda_2"] = converted["lambda_1"] * converted["mass_1"] ** 5 / converted["mass_2"] ** 5
    added_keys = [key for key in converted if key not in original_keys]
    return converted, added_keys


# ---- post-processing over posterior rows, batched on the device --------------------
def _columns(samples):
    """dict of arrays or pandas DataFrame -> (dict of float64 arrays for the numeric columns, is_frame)."""
    try:
        from pandas import DataFrame
    except ImportError:  # pragma: no cover
        DataFrame = ()
    if isinstance(samples, DataFrame):
        cols = {k: samples[k].to_numpy(dtype=np.float64) for k in samples.columns
                if np.issubdtype(samples[k].dtype, np.number) and not np.issubdtype(samples[k].dtype, np.complexfloating)}
        return cols, True
    return {k: np.atleast_1d(np.asarray(v, dtype=np.float64)) for k, v in samples.items()
            if not np.iscomplexobj(v)}, False


def compute_snrs(sample, likelihood, npool=1):
    """bilby/gw/conversion.py:2215-2271: adds ``<IFO>_matched_filter_snr`` (complex) and ``<IFO>_optimal_snr`` to a
    parameter dict or to every row of a DataFrame.  The reference maps rows over a process pool (``npool``, ignored
    here); this evaluates all rows in one pass of K0 + K1."""
    if likelihood is None:
        return
    cols, is_frame = _columns(sample)
    snrs = likelihood.compute_snrs_batch(cols)
    for key, val in snrs.items():
        if is_frame:
            sample[key] = val
        else:
            sample[key] = val if np.ndim(next(iter(sample.values()))) else val[0]


def generate_posterior_samples_from_marginalized_likelihood(samples, likelihood, npool=1, block=10, use_cache=True,
                                                            rng=None):
    """bilby/gw/conversion.py:2400-2492 (the cache / pool arguments are accepted and ignored): a dict passes through
    unchanged like in the reference; a DataFrame gets new ``geocent_time`` / ``luminosity_distance`` / ``phase``
    columns from one batched call of ``bb_reconstruct_marginalized_device``."""
    if len(getattr(likelihood, "_marginalized_parameters", [])) == 0 or isinstance(samples, dict):
        return samples
    cols, is_frame = _columns(samples)
    if not is_frame:
        raise ValueError("Unable to handle input samples of type {}".format(type(samples)))
    new = likelihood.generate_posterior_samples_from_marginalized_likelihood_batch(cols, rng=rng)
    for key in ("geocent_time", "luminosity_distance", "phase"):
        if key in new:
            samples[key] = new[key]
    return samples
