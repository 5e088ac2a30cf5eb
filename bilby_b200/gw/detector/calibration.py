"""Calibration models (host-side description; evaluation happens inside the fused kernel).
Mirrors bilby/gw/detector/calibration.py: Recalibrate :209-254 (identity), CubicSpline :257-384."""
import numpy as np


class Recalibrate:
    name = "none"

    def __init__(self, prefix="recalib_"):
        self.prefix = prefix
        self.params = dict()

    def __repr__(self):
        return self.__class__.__name__ + "(prefix='{}')".format(self.prefix)

    def get_calibration_factor(self, frequency_array, **params):
        return np.ones_like(frequency_array)


class CubicSpline(Recalibrate):
    name = "cubic_spline"

    def __init__(self, prefix, minimum_frequency, maximum_frequency, n_points):
        super().__init__(prefix=prefix)
        if n_points < 4:
            raise ValueError("Cubic spline calibration requires at least 4 spline nodes.")
        self.n_points = n_points
        self.minimum_frequency = minimum_frequency
        self.maximum_frequency = maximum_frequency
        self._log_spline_points = np.linspace(np.log10(minimum_frequency), np.log10(maximum_frequency), n_points)

    @property
    def log_spline_points(self):
        return self._log_spline_points

    @property
    def delta_log_spline_points(self):
        return self._log_spline_points[1] - self._log_spline_points[0]

    @property
    def nodes_to_spline_coefficients(self):
        """calibration.py:302-325 (LIGO-T2300140 Eq. 9)."""
        n = self.n_points
        tmp1 = np.zeros((n, n))
        tmp1[0, 0], tmp1[0, 1], tmp1[0, 2] = -1, 2, -1
        tmp1[-1, -3], tmp1[-1, -2], tmp1[-1, -1] = -1, 2, -1
        tmp2 = np.zeros((n, n))
        for i in range(1, n - 1):
            tmp1[i, i - 1], tmp1[i, i], tmp1[i, i + 1] = 1 / 6, 2 / 3, 1 / 6
            tmp2[i, i - 1], tmp2[i, i], tmp2[i, i + 1] = 1, -2, 1
        return np.linalg.solve(tmp1, tmp2)

    def get_calibration_factor(self, frequency_array, prefix="recalib_", **params):
        """calibration.py:349-384 on the host (set-up only: fiducial waveforms of the relative-binning likelihood;
        the per-sample factor is evaluated inside the CUDA kernels)."""
        f = np.asarray(frequency_array, dtype=float)
        with np.errstate(divide="ignore"):
            x = np.nan_to_num(np.log10(f) - self.log_spline_points[0], neginf=0.0) / self.delta_log_spline_points
        prev = np.clip(x.astype(int), 0, self.n_points - 2)
        b = x - prev
        a = 1 - b
        c = (a ** 3 - a) / 6
        d = (b ** 3 - b) / 6
        out = []
        for kind in ("amplitude", "phase"):
            p = np.array([params[f"{prefix}{kind}_{ii}"] for ii in range(self.n_points)], dtype=float)
            sc = self.nodes_to_spline_coefficients.dot(p)
            out.append(a * p[prev] + b * p[prev + 1] + c * sc[prev] + d * sc[prev + 1])
        da, dp = out
        return np.nan_to_num((1 + da) * (2 + 1j * dp) / (2 - 1j * dp))

    def __repr__(self):
        return (f"{self.__class__.__name__}(prefix='{self.prefix}', minimum_frequency={self.minimum_frequency}, "
                f"maximum_frequency={self.maximum_frequency}, n_points={self.n_points})")


def curves_from_spline_and_prior(parameters, label, n_points, frequency_array, n_curves):
    """bilby/gw/detector/calibration.py:578-591: ``parameters`` is a dict of arrays (or DataFrame) holding
    ``recalib_<label>_{amplitude,phase}_<i>`` draws; returns complex curves [n_curves, len(frequency_array)]."""
    spline = CubicSpline(prefix=f"recalib_{label}_", minimum_frequency=frequency_array[0],
                         maximum_frequency=frequency_array[-1], n_points=n_points)
    curves = []
    for ii in range(n_curves):
        row = {k: np.asarray(parameters[k])[ii] for k in parameters.keys()}
        curves.append(spline.get_calibration_factor(frequency_array, prefix=spline.prefix, **row))
    return curves


def build_calibration_lookup(interferometers, lookup_files=None, priors=None, number_of_response_curves=1000,
                             starting_index=0, rng=None):
    """bilby/gw/detector/calibration.py:503-575.  HDF5 look-up files need h5py (absent here): ``lookup_files`` may
    instead map a detector name to an array of response curves [n_curves, n_masked_bins] or to a dict / DataFrame
    of spline-node draws.  Without an entry the curves are drawn from the ``recalib_<IFO>_*`` priors of a
    CubicSpline model, as the reference does; like the reference the detector's model is then reset to the
    identity (:552)."""
    if lookup_files is None and priors is None:
        raise ValueError("One of calibration_lookup_table or priors must be specified for "
                         "building calibration marginalization lookup table.")
    lookup_files = dict() if lookup_files is None else lookup_files
    draws, parameters = dict(), dict()
    for interferometer in interferometers:
        name = interferometer.name
        frequencies = interferometer.frequency_array[interferometer.frequency_mask]
        entry = lookup_files.get(name)
        if isinstance(entry, str):
            raise NotImplementedError("HDF5 calibration files need h5py, which is not available in this image")
        idxs = np.arange(number_of_response_curves, dtype=int) + starting_index
        if isinstance(entry, np.ndarray):
            draws[name] = np.asarray(entry)[idxs]
            parameters[name] = None
        else:
            if entry is not None:
                pars = {k: np.asarray(entry[k])[idxs] for k in entry.keys()}
            else:
                if priors is None:
                    raise ValueError("Priors must be passed to generate calibration response curves for cubic spline.")
                if rng is None:
                    from ...core.utils import random
                    rng = random.rng
                pars = {k: np.asarray(priors[k].sample(number_of_response_curves, rng=rng))
                        for k in priors.keys() if "recalib" in k and name in k and hasattr(priors[k], "sample")}
            n_points = getattr(interferometer.calibration_model, "n_points", None)
            if n_points is None:
                n_points = len([k for k in pars if "amplitude" in k])
            draws[name] = np.array(curves_from_spline_and_prior(pars, name, n_points, frequencies,
                                                                number_of_response_curves))
            parameters[name] = pars
        interferometer.calibration_model = Recalibrate()
    return draws, parameters
