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

    def __repr__(self):
        return (f"{self.__class__.__name__}(prefix='{self.prefix}', minimum_frequency={self.minimum_frequency}, "
                f"maximum_frequency={self.maximum_frequency}, n_points={self.n_points})")
