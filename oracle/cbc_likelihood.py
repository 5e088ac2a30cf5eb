"""TEST INFRASTRUCTURE ONLY (oracle) - numpy restatement of the reference's compact-binary
likelihood hot path (bilby, /root/reference).  Every function cites the reference lines it
follows.  This module is validated against the UNMODIFIED reference in the build container
(oracle/tools/make_golden.py -> tests/golden/*.npz, tests/test_oracle_vs_golden.py) because
the reference itself (Python) cannot travel to the GPU box.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (bilby_b200) never does.
"""
import os
import numpy as np
from scipy.interpolate import interp1d, RectBivariateSpline
from scipy.special import i0e, logsumexp

from . import phenomd as _pd

SPEED_OF_LIGHT = 299792458.0

# --------------------------------------------------------------------------------------
# time  (bilby/gw/time.py:57-66, 69-93, 96-192)
# --------------------------------------------------------------------------------------
LEAP_SECONDS = np.array([
    46828800, 78364801, 109900802, 173059203, 252028804, 315187205, 346723206, 393984007,
    425520008, 457056009, 504489610, 551750411, 599184012, 820108813, 914803214, 1025136015,
    1119744016, 1167264017])


def greenwich_mean_sidereal_time(gps_time):
    """bilby/gw/time.py:114-165 (equation of equinoxes = 0)."""
    gps_int = gps_time // 1
    n_leap = np.sum(gps_int > LEAP_SECONDS)
    second = gps_int - n_leap
    # datetime(1980, 1, 6).julian_day with .second = second  (time.py:57-66)
    julian_day = (367 * 1980 - 7 * (1980 + (1 + 9) // 12) // 4 + 275 * 1 // 9 + 6
                  + second / 86400.0 + 1721013.5)
    t_hi = (julian_day - 2451545.0) / 36525.0
    t_lo = (gps_time % 1) / (36525.0 * 86400.0)
    t = t_hi + t_lo
    sidereal_time = gps_time * 0 + (-6.2e-6 * t + 0.093104) * t ** 2 + 67310.54841
    sidereal_time += 8640184.812866 * t_lo
    sidereal_time += 3155760000.0 * t_lo
    sidereal_time += 8640184.812866 * t_hi
    sidereal_time += 3155760000.0 * t_hi
    return sidereal_time * 2 * np.pi / 86400.0


# --------------------------------------------------------------------------------------
# geometry  (bilby/gw/geometry.py:51-115, 118-186, 261-343; bilby/gw/utils.py:59-88)
# --------------------------------------------------------------------------------------
def calculate_arm(arm_tilt, arm_azimuth, longitude, latitude):
    e_long = np.array([-np.sin(longitude), np.cos(longitude), 0.0])
    e_lat = np.array([-np.sin(latitude) * np.cos(longitude),
                      -np.sin(latitude) * np.sin(longitude), np.cos(latitude)])
    e_h = np.array([np.cos(latitude) * np.cos(longitude),
                    np.cos(latitude) * np.sin(longitude), np.sin(latitude)])
    return (np.cos(arm_tilt) * np.cos(arm_azimuth) * e_long
            + np.cos(arm_tilt) * np.sin(arm_azimuth) * e_lat
            + np.sin(arm_tilt) * e_h)


def detector_tensor(x, y):
    return (np.outer(x, x) - np.outer(y, y)) / 2


def vertex_position_geocentric(latitude, longitude, elevation):
    a = 6378137
    b = 6356752.314
    radius = a ** 2 * (a ** 2 * np.cos(latitude) ** 2 + b ** 2 * np.sin(latitude) ** 2) ** (-0.5)
    x = (radius + elevation) * np.cos(latitude) * np.cos(longitude)
    y = (radius + elevation) * np.cos(latitude) * np.sin(longitude)
    z = ((b / a) ** 2 * radius + elevation) * np.sin(latitude)
    return np.array([x, y, z])


def polarization_tensors(ra, dec, time, psi):
    """plus and cross tensors, geometry.py:143-176."""
    gmst = greenwich_mean_sidereal_time(time) % (2 * np.pi)
    phi = ra - gmst
    theta = np.pi / 2 - dec
    u = np.array([np.cos(phi) * np.cos(theta), np.cos(theta) * np.sin(phi), -np.sin(theta)])
    v = np.array([-np.sin(phi), np.cos(phi), 0.0])
    m = -u * np.sin(psi) - v * np.cos(psi)
    n = -u * np.cos(psi) + v * np.sin(psi)
    return np.outer(m, m) - np.outer(n, n), np.outer(m, n) + np.outer(n, m)


def time_delay_from_geocenter(vertex, ra, dec, time):
    """geometry.py:282-343 (detector2 = 0)."""
    gmst = greenwich_mean_sidereal_time(time) % (2 * np.pi)
    phi = ra - gmst
    theta = np.pi / 2 - dec
    omega = np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])
    delta_d = np.zeros(3) - vertex
    return omega @ delta_d / SPEED_OF_LIGHT


# detector files bilby/gw/detector/detectors/{H1,L1,V1}.interferometer (physical site constants)
DETECTORS = {
    "H1": dict(latitude=46 + 27. / 60 + 18.528 / 3600, longitude=-(119 + 24. / 60 + 27.5657 / 3600),
               elevation=142.554, xarm_azimuth=125.9994, yarm_azimuth=215.9994,
               xarm_tilt=-6.195e-4, yarm_tilt=1.25e-5, curve="aLIGO_O4_high_asd.txt"),
    "L1": dict(latitude=30 + 33. / 60 + 46.4196 / 3600, longitude=-(90 + 46. / 60 + 27.2654 / 3600),
               elevation=-6.574, xarm_azimuth=197.7165, yarm_azimuth=287.7165,
               xarm_tilt=-3.121e-4, yarm_tilt=-6.107e-4, curve="aLIGO_O4_high_asd.txt"),
    "V1": dict(latitude=43 + 37. / 60 + 53.0921 / 3600, longitude=10 + 30. / 60 + 16.1878 / 3600,
               elevation=51.884, xarm_azimuth=70.5674, yarm_azimuth=160.5674,
               xarm_tilt=0.0, yarm_tilt=0.0, curve="AdV_psd.txt"),
}

_CURVES = None


def load_curve(name):
    """(frequency, psd) of a packed noise curve (psd.py:340-349: asd files are squared)."""
    global _CURVES
    if _CURVES is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bilby_b200",
                            "data", "noise_curves.npz")
        _CURVES = dict(np.load(path))
    f = _CURVES[name + ":frequency"]
    val = _CURVES[name + ":value"]
    if name.endswith("_asd.txt"):
        val = val ** 2
    return f, val


def create_frequency_series(sampling_frequency, duration):
    """bilby/core/utils/series.py:115-134."""
    n_samples = np.round(duration * sampling_frequency)
    n_freq = int(np.round(n_samples / 2) + 1)
    return np.linspace(0, sampling_frequency / 2, num=n_freq)


class OracleInterferometer:
    """The per-detector state the hot path reads (interferometer.py:303-368, 551-564, 607-640;
    strain_data.py:142-159; psd.py:236-258)."""

    def __init__(self, name, sampling_frequency, duration, start_time,
                 minimum_frequency=20.0, maximum_frequency=None, spec=None):
        spec = dict(DETECTORS[name]) if spec is None else dict(spec)
        self.name = name
        self.duration = float(duration)
        self.sampling_frequency = float(sampling_frequency)
        self.start_time = float(start_time)
        lat = np.deg2rad(spec["latitude"]) if False else spec["latitude"] * np.pi / 180
        lon = spec["longitude"] * np.pi / 180
        self.x = calculate_arm(spec["xarm_tilt"], spec["xarm_azimuth"] * np.pi / 180, lon, lat)
        self.y = calculate_arm(spec["yarm_tilt"], spec["yarm_azimuth"] * np.pi / 180, lon, lat)
        self.detector_tensor = detector_tensor(self.x, self.y)
        self.vertex = vertex_position_geocentric(lat, lon, spec["elevation"])
        self.frequency_array = create_frequency_series(sampling_frequency, duration)
        self.minimum_frequency = minimum_frequency
        nyq = sampling_frequency / 2
        self.maximum_frequency = nyq if maximum_frequency is None else min(maximum_frequency, nyq)
        self.frequency_mask = ((self.frequency_array >= self.minimum_frequency)
                               & (self.frequency_array <= self.maximum_frequency))
        cf, cpsd = load_curve(spec["curve"])
        self.power_spectral_density_array = interp1d(cf, cpsd, bounds_error=False,
                                                     fill_value=np.inf)(self.frequency_array)
        self.frequency_domain_strain = np.zeros(len(self.frequency_array), dtype=complex)
        self.calibration = None   # or OracleCubicSpline
        self.reference_time = None   # interferometer.py:336-339: antenna response at this time when set

    def set_gaussian_noise(self, rng):
        """psd.py:350-376 + series.py:161-198 semantics, but drawing from a caller-supplied
        numpy Generator (the reference uses its global bilby.core.utils.random.rng)."""
        n = len(self.frequency_array)
        norm1 = 0.5 * self.duration ** 0.5
        re1, im1 = rng.normal(0, norm1, (2, n))
        white = re1 + 1j * im1
        white[0] = 0
        if np.mod(np.round(self.duration * self.sampling_frequency), 2) == 0:
            white[-1] = 0
        with np.errstate(invalid="ignore"):
            out = white * self.power_spectral_density_array ** 0.5
        out[~np.isfinite(out)] = 0
        self.frequency_domain_strain = out * self.frequency_mask

    def antenna_response(self, ra, dec, time, psi):
        plus, cross = polarization_tensors(ra, dec, time, psi)
        return (np.einsum("ij,ij->", self.detector_tensor, plus),
                np.einsum("ij,ij->", self.detector_tensor, cross))

    def get_detector_response(self, pols, parameters, frequencies=None):
        """interferometer.py:303-368."""
        if frequencies is None:
            frequencies = self.frequency_array
            mask = self.frequency_mask
        else:
            mask = np.ones(len(frequencies), dtype=bool)
        antenna_time = parameters["geocent_time"] if self.reference_time is None else self.reference_time
        fp, fc = self.antenna_response(parameters["ra"], parameters["dec"], antenna_time, parameters["psi"])
        signal = pols["plus"] * mask * fp + pols["cross"] * mask * fc
        time_shift = time_delay_from_geocenter(self.vertex, parameters["ra"], parameters["dec"],
                                               parameters["geocent_time"])
        dt_geocent = parameters["geocent_time"] - self.start_time
        dt = dt_geocent + time_shift
        signal = signal * np.exp(-1j * 2 * np.pi * dt * frequencies)
        if self.calibration is not None:
            signal = signal * self.calibration.get_calibration_factor(
                frequencies, prefix=f"recalib_{self.name}_", **parameters)
        return signal

    def inner_product(self, signal):
        """interferometer.py:624-640 -> gw/utils.py:118-138."""
        m = self.frequency_mask
        return 4 / self.duration * np.sum(
            signal[m].conj() * self.frequency_domain_strain[m] / self.power_spectral_density_array[m])

    def optimal_snr_squared(self, signal):
        m = self.frequency_mask
        return 4 / self.duration * np.sum(
            signal[m].conj() * signal[m] / self.power_spectral_density_array[m])


class OracleCubicSpline:
    """calibration.py:257-384."""

    def __init__(self, prefix, minimum_frequency, maximum_frequency, n_points):
        self.prefix = prefix
        self.n_points = n_points
        self.log_spline_points = np.linspace(np.log10(minimum_frequency), np.log10(maximum_frequency),
                                             n_points)
        self.delta = self.log_spline_points[1] - self.log_spline_points[0]
        n = n_points
        tmp1 = np.zeros((n, n))
        tmp1[0, 0], tmp1[0, 1], tmp1[0, 2] = -1, 2, -1
        tmp1[-1, -3], tmp1[-1, -2], tmp1[-1, -1] = -1, 2, -1
        tmp2 = np.zeros((n, n))
        for i in range(1, n - 1):
            tmp1[i, i - 1], tmp1[i, i], tmp1[i, i + 1] = 1 / 6, 2 / 3, 1 / 6
            tmp2[i, i - 1], tmp2[i, i], tmp2[i, i + 1] = 1, -2, 1
        self.nodes_to_spline_coefficients = np.linalg.solve(tmp1, tmp2)

    def bin_weights(self, frequency_array):
        with np.errstate(divide="ignore"):
            x = np.nan_to_num(np.log10(frequency_array) - self.log_spline_points[0], neginf=0.0) / self.delta
        prev = np.clip(x.astype(int), 0, self.n_points - 2)
        b = x - prev
        a = 1 - b
        c = (a ** 3 - a) / 6
        d = (b ** 3 - b) / 6
        return prev, a, b, c, d

    def get_calibration_factor(self, frequency_array, prefix=None, **params):
        prefix = self.prefix if prefix is None else prefix
        prev, a, b, c, d = self.bin_weights(frequency_array)
        out = []
        for kind in ("amplitude", "phase"):
            p = np.array([params[f"{prefix}{kind}_{ii}"] for ii in range(self.n_points)])
            sc = self.nodes_to_spline_coefficients.dot(p)
            out.append(a * p[prev] + b * p[prev + 1] + c * sc[prev] + d * sc[prev + 1])
        da, dp = out
        return np.nan_to_num((1 + da) * (2 + 1j * dp) / (2 - 1j * dp))


def curves_from_spline_nodes(name, nodes, frequency_array, n_points):
    """calibration.py:578-591 curves_from_spline_and_prior: nodes [n_curves, 2 (amplitude, phase), n_points] ->
    complex response curves [n_curves, len(frequency_array)]; the spline spans frequency_array[0] .. [-1]."""
    spline = OracleCubicSpline(f"recalib_{name}_", frequency_array[0], frequency_array[-1], n_points)
    curves = []
    for c in range(len(nodes)):
        params = {}
        for i in range(n_points):
            params[f"recalib_{name}_amplitude_{i}"] = nodes[c, 0, i]
            params[f"recalib_{name}_phase_{i}"] = nodes[c, 1, i]
        curves.append(spline.get_calibration_factor(frequency_array, **params))
    return np.array(curves)


# --------------------------------------------------------------------------------------
# source model (bilby/gw/source.py:269-348, 552-690; conversion.py:146-153)
# --------------------------------------------------------------------------------------
def lal_binary_black_hole(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12,
                          a_2, tilt_2, phi_jl, theta_jn, phase, **kwargs):
    """bilby-signature source model backed by the restated IMRPhenomD (oracle/phenomd.py)."""
    wa = dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0,
              maximum_frequency=frequency_array[-1], catch_waveform_errors=False)
    wa.update(kwargs)
    if wa["waveform_approximant"] != "IMRPhenomD":
        raise ValueError("oracle restates IMRPhenomD only for lal_binary_black_hole")
    for a, tilt in ((a_1, tilt_1), (a_2, tilt_2)):
        if not (a == 0 or tilt in (0, np.pi)):
            raise ValueError("IMRPhenomD is an aligned-spin model")
    s1z = a_1 * np.cos(tilt_1)
    s2z = a_2 * np.cos(tilt_2)
    delta_f = frequency_array[1] - frequency_array[0]
    bounds = (frequency_array >= wa["minimum_frequency"]) * (frequency_array <= wa["maximum_frequency"])
    try:
        hp, hc = _pd.choose_fd_waveform_phenomd(
            np.arange(len(frequency_array)) * delta_f,
            mass_1, mass_2, s1z, s2z,
            luminosity_distance * 1e6 * _pd.PARSEC, theta_jn, phase,
            wa["minimum_frequency"], wa["maximum_frequency"], wa["reference_frequency"], delta_f)
    except _pd.WaveformDomainError:
        if wa["catch_waveform_errors"]:
            return None
        raise
    return dict(plus=hp * bounds, cross=hc * bounds)


def lal_binary_neutron_star(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12,
                            a_2, tilt_2, phi_jl, theta_jn, phase, lambda_1, lambda_2, **kwargs):
    """bilby-signature BNS source model (source.py:351-432) backed by the restated TaylorF2 (+tides)."""
    from . import taylorf2 as _tf2
    wa = dict(waveform_approximant="TaylorF2", reference_frequency=50.0, minimum_frequency=20.0,
              maximum_frequency=frequency_array[-1], catch_waveform_errors=False)
    wa.update(kwargs)
    if wa["waveform_approximant"] != "TaylorF2":
        raise ValueError("oracle restates TaylorF2 only for lal_binary_neutron_star")
    s1z = a_1 * np.cos(tilt_1)
    s2z = a_2 * np.cos(tilt_2)
    delta_f = frequency_array[1] - frequency_array[0]
    bounds = (frequency_array >= wa["minimum_frequency"]) * (frequency_array <= wa["maximum_frequency"])
    try:
        hp, hc = _tf2.choose_fd_waveform_taylorf2(
            np.arange(len(frequency_array)) * delta_f, mass_1, mass_2, s1z, s2z, lambda_1, lambda_2,
            luminosity_distance * 1e6 * _pd.PARSEC, theta_jn, phase,
            wa["minimum_frequency"], wa["maximum_frequency"], wa["reference_frequency"], delta_f)
    except _pd.WaveformDomainError:
        if wa["catch_waveform_errors"]:
            return None
        raise
    return dict(plus=hp * bounds, cross=hc * bounds)


def convert_to_lal_binary_black_hole_parameters(parameters):
    """Subset of bilby/gw/conversion.py:182-283 used by the benchmark priors:
    (chirp_mass, mass_ratio) -> (mass_1, mass_2); chi_i -> (a_i, cos_tilt_i); cos_theta_jn;
    missing precession angles -> 0."""
    p = dict(parameters)
    if "mass_1" not in p or "mass_2" not in p:
        if "chirp_mass" in p and "mass_ratio" in p:
            q = p["mass_ratio"]
            total = p["chirp_mass"] * (1 + q) ** 1.2 / q ** 0.6      # conversion.py:969-989
            p["mass_1"] = total / (1 + q)                           # conversion.py:849-870
            p["mass_2"] = p["mass_1"] * q
        elif "total_mass" in p and "mass_ratio" in p:
            q = p["mass_ratio"]
            p["mass_1"] = p["total_mass"] / (1 + q)
            p["mass_2"] = p["mass_1"] * q
    for idx in ("1", "2"):
        key = f"chi_{idx}"
        if key in p:
            p[f"a_{idx}"] = abs(p[key])
            p[f"cos_tilt_{idx}"] = np.sign(p[key])
        elif f"a_{idx}" not in p:
            p[f"a_{idx}"] = 0.0
            p[f"cos_tilt_{idx}"] = 1.0
    for angle in ("tilt_1", "tilt_2", "theta_jn"):
        if f"cos_{angle}" in p and angle not in p:
            p[angle] = np.arccos(p[f"cos_{angle}"])
    for angle in ("tilt_1", "tilt_2", "phi_12", "phi_jl"):
        p.setdefault(angle, 0.0)
    if "delta_phase" in p:
        p["phase"] = np.mod(p["delta_phase"] - np.sign(np.cos(p["theta_jn"])) * p["psi"], 2 * np.pi)
    return p


SOURCE_ARGS = ("mass_1", "mass_2", "luminosity_distance", "a_1", "tilt_1", "phi_12", "a_2", "tilt_2",
               "phi_jl", "theta_jn", "phase")


# --------------------------------------------------------------------------------------
# likelihood (bilby/gw/likelihood/base.py)
# --------------------------------------------------------------------------------------
def ln_i0(value):
    """bilby/gw/utils.py:1006-1022."""
    return np.log(i0e(value)) + np.abs(value)


def interped_sample(xx, yy, u):
    """Interped(xx, yy).rescale(u): core/prior/interpolated.py:12-60, 88-94, 161-176 (grid replaced by a linspace of
    the same length, density re-interpolated linearly, trapezoid normalisation, cumulative trapezoid with the last
    element forced to one, linear inverse interpolation)."""
    from scipy.integrate import cumulative_trapezoid
    xx = np.asarray(xx, dtype=float)
    yy = np.asarray(yy, dtype=float)
    all_interpolated = interp1d(x=xx, y=yy, bounds_error=False, fill_value=0)
    grid = np.linspace(float(min(xx)), float(max(xx)), len(xx))
    dens = all_interpolated(grid)
    dens = dens / np.trapezoid(dens, grid)
    cdf = cumulative_trapezoid(dens, grid, initial=0)
    cdf[-1] = 1
    return float(interp1d(x=cdf, y=grid, bounds_error=True)(u))


class OracleUniform:
    def __init__(self, minimum, maximum):
        self.minimum, self.maximum = minimum, maximum

    def prob(self, val):
        """core/prior/analytical.py:231-242 (inclusive at both ends)."""
        val = np.asarray(val)
        return ((val >= self.minimum) & (val <= self.maximum)) / (self.maximum - self.minimum)

    def rescale(self, u):
        return self.minimum + u * (self.maximum - self.minimum)


class OraclePowerLaw:
    """core/prior/analytical.py:107-147."""

    def __init__(self, alpha, minimum, maximum):
        self.alpha, self.minimum, self.maximum = alpha, minimum, maximum

    def rescale(self, val):
        if self.alpha == -1:
            return self.minimum * np.exp(val * np.log(self.maximum / self.minimum))
        return (self.minimum ** (1 + self.alpha)
                + val * (self.maximum ** (1 + self.alpha) - self.minimum ** (1 + self.alpha))) ** (
            1. / (1 + self.alpha))

    def prob(self, val):
        val = np.asarray(val, dtype=float)
        inside = (val >= self.minimum) & (val <= self.maximum)
        if self.alpha == -1:
            return np.nan_to_num(1 / val / np.log(self.maximum / self.minimum)) * inside
        return np.nan_to_num(val ** self.alpha * (1 + self.alpha)
                             / (self.maximum ** (1 + self.alpha) - self.minimum ** (1 + self.alpha))) * inside


def _lookup_rows(rows, ref_dist, distance_array, x_ref, y_ref, phase_marginalization, prior_term):
    scaling = ref_dist / distance_array
    d_full = np.outer(x_ref, scaling)
    if phase_marginalization:
        d_full = ln_i0(abs(d_full))
    out = np.zeros((len(rows), len(x_ref)))
    for n, ii in enumerate(rows):
        out[n] = logsumexp(d_full - (y_ref[ii] * scaling ** 2) / 2, b=prior_term, axis=1)
    return out


class OracleLikelihood:
    """Restates GravitationalWaveTransient (bilby/gw/likelihood/base.py:150-229, 260-354, 419-477,
    775-820, 879-914, 994-1035) for sky reference frame, geocenter time reference, no calibration
    marginalisation."""

    def __init__(self, interferometers, source_model=lal_binary_black_hole, waveform_arguments=None,
                 parameter_conversion=convert_to_lal_binary_black_hole_parameters,
                 time_marginalization=False, distance_marginalization=False, phase_marginalization=False,
                 distance_prior=None, time_prior=None, jitter_time=True, lookup_table=None, table_processes=1,
                 calibration_draws=None):
        self.ifos = list(interferometers)
        self.duration = self.ifos[0].duration
        self.sampling_frequency = self.ifos[0].sampling_frequency
        self.start_time = self.ifos[0].start_time
        self.frequency_array = self.ifos[0].frequency_array
        self.source_model = source_model
        self.waveform_arguments = dict(waveform_arguments or {})
        self.parameter_conversion = parameter_conversion
        self.time_marginalization = time_marginalization
        self.distance_marginalization = distance_marginalization
        self.phase_marginalization = phase_marginalization
        self.jitter_time = jitter_time and time_marginalization
        self.time_prior = time_prior
        # calibration marginalisation (base.py:1037-1051): {detector name: complex [n_curves, n_masked_bins]}
        self.calibration_marginalization = calibration_draws is not None
        if self.calibration_marginalization:
            if time_marginalization and distance_marginalization:
                # the reference itself fails here (base.py:775-784 broadcasts [n_curves, n_times] against [n_curves])
                raise ValueError("time + calibration + distance marginalisation: shape mismatch in the reference")
            self.calibration_draws = {k: np.asarray(v) for k, v in calibration_draws.items()}
            self.calibration_abs_draws = {k: np.abs(v) ** 2 for k, v in self.calibration_draws.items()}
            self.number_of_response_curves = len(next(iter(self.calibration_draws.values())))
        if time_marginalization:
            self._delta_tc = 2 / self.sampling_frequency
            self._times = self.start_time + np.linspace(
                0, self.duration, int(self.duration / 2 * self.sampling_frequency + 1))[1:]
        if distance_marginalization:
            self.distance_prior = distance_prior
            self._distance_array = np.linspace(distance_prior.minimum, distance_prior.maximum, int(1e4))
            self.distance_prior_array = np.array([distance_prior.prob(d) for d in self._distance_array])
            self._ref_dist = distance_prior.rescale(0.5)
            if lookup_table is None:
                lookup_table = self.create_lookup_table(processes=table_processes)
            self._dist_margd_loglikelihood_array = lookup_table
            self._interp = RectBivariateSpline(self._d_inner_h_ref_array, self._optimal_snr_squared_ref_array,
                                               lookup_table.T, kx=3, ky=3, s=0)

    # ---- distance marginalisation set-up (base.py:894-914, 994-1018)
    @property
    def _optimal_snr_squared_ref_array(self):
        return np.logspace(-5, 10, 400)

    @property
    def _d_inner_h_ref_array(self):
        if self.phase_marginalization:
            return np.logspace(-5, 10, 800)
        return np.hstack((-np.logspace(3, -3, 400), np.logspace(-3, 10, 400)))

    def create_lookup_table(self, rows=None, processes=1):
        """base.py:994-1018; ``rows`` restricts to a subset of optimal-SNR rows (tests); ``processes`` > 1
        spreads the rows over a multiprocessing pool (same arithmetic per entry)."""
        table = np.zeros((400, 800))
        idx = list(range(400)) if rows is None else list(rows)
        prior_term = self.distance_prior_array * (self._distance_array[1] - self._distance_array[0])
        args = (self._ref_dist, self._distance_array, self._d_inner_h_ref_array,
                self._optimal_snr_squared_ref_array, self.phase_marginalization, prior_term)
        if processes > 1:
            import multiprocessing
            chunks = [idx[i::processes] for i in range(processes)]
            with multiprocessing.get_context("fork").Pool(processes) as pool:
                parts = pool.starmap(_lookup_rows, [(c,) + args for c in chunks])
            for c, part in zip(chunks, parts):
                table[c] = part
        else:
            table[idx] = _lookup_rows(idx, *args)
        log_norm = logsumexp(0 / self._distance_array, b=prior_term)
        table -= log_norm
        return table

    def interp_dist(self, d_inner_h_ref, h_inner_h_ref):
        """calculus.py:221-262 BoundedRectBivariateSpline with fill_value=-inf."""
        x = np.atleast_1d(np.asarray(d_inner_h_ref, dtype=float))
        y = np.broadcast_to(np.asarray(h_inner_h_ref, dtype=float), x.shape)
        res = self._interp(x, y, grid=False)
        xa, ya = self._d_inner_h_ref_array, self._optimal_snr_squared_ref_array
        bad = (x < xa.min()) | (x > xa.max()) | (y < ya.min()) | (y > ya.max())
        res = np.where(bad, -np.inf, res)
        return res

    # ---- per-sample pieces
    def polarizations(self, parameters):
        """waveform_generator.py:178-209, 260-269."""
        p = self.parameter_conversion(parameters)
        args = {k: p[k] for k in SOURCE_ARGS if k in p}
        if "neutron" in getattr(self.source_model, "__name__", ""):
            for k in ("lambda_1", "lambda_2"):
                args[k] = p.get(k, 0.0)
        return self.source_model(self.frequency_array, **args, **self.waveform_arguments)

    def calculate_snrs(self, pols, ifo, parameters):
        """base.py:260-354."""
        signal = ifo.get_detector_response(pols, parameters)
        d_inner_h = ifo.inner_product(signal)
        hh = ifo.optimal_snr_squared(signal).real
        d_inner_h_array = None
        if self.time_marginalization:
            d_inner_h_array = 4 / self.duration * np.fft.fft(
                signal[0:-1] * ifo.frequency_domain_strain.conj()[0:-1]
                / ifo.power_spectral_density_array[0:-1])
        return d_inner_h, hh, d_inner_h_array

    def log_likelihood_ratio(self, parameters, return_snrs=False):
        """base.py:419-477."""
        parameters = dict(parameters)
        pols = self.polarizations(parameters)
        if pols is None:
            return np.nan_to_num(-np.inf)
        if self.time_marginalization and self.jitter_time:
            parameters["geocent_time"] = parameters["geocent_time"] + parameters["time_jitter"]
        d_inner_h = 0j
        hh = 0.0
        arr = None
        per_det = []
        for ifo in self.ifos:
            d, h, a = self.calculate_snrs(pols, ifo, parameters)
            d_inner_h += d
            hh += h
            per_det.append((d, h))
            if a is not None:
                arr = a if arr is None else arr + a
        if return_snrs:
            return per_det
        if self.calibration_marginalization:
            return float(np.real(self.calibration_marginalized_likelihood(pols, parameters)))
        if self.time_marginalization:
            log_l = self.time_marginalized_likelihood(arr, hh, parameters)
        elif self.distance_marginalization:
            log_l = self.distance_marginalized_likelihood(d_inner_h, hh, parameters)
        elif self.phase_marginalization:
            log_l = ln_i0(abs(d_inner_h)) - hh / 2
        else:
            log_l = np.real(d_inner_h) - hh / 2
        return float(np.real(log_l))

    def calibration_log_likelihoods(self, pols, parameters):
        """base.py:333-346 (per-detector arrays over the response curves; note conj(d) * h, the conjugate of
        inner_product's convention), :109-148 (arrays add over detectors), :822-858: the point likelihood per curve.
        With time marginalisation (base.py:305-323, 860-866): one transform of the calibrated integrand per curve,
        [n_curves, N - 1], then the time-marginalised likelihood per curve (base.py:794-820, 786-792)."""
        d_arr, hh_arr = 0, 0
        for ifo in self.ifos:
            signal = ifo.get_detector_response(pols, parameters)
            m = ifo.frequency_mask
            norm = 4 / self.duration
            integrand = norm * ifo.frequency_domain_strain.conj() * signal / ifo.power_spectral_density_array
            if self.time_marginalization:
                tiled = np.tile(integrand, (self.number_of_response_curves, 1)).T          # [N, n_curves]
                tiled[m] *= self.calibration_draws[ifo.name].T
                d_arr = d_arr + np.fft.fft(tiled[0:-1], axis=0).T
            else:
                d_arr = d_arr + np.dot(integrand[m], self.calibration_draws[ifo.name].T)
            hh_integrand = norm * np.abs(signal) ** 2 / ifo.power_spectral_density_array
            hh_arr = hh_arr + np.dot(hh_integrand[m], self.calibration_abs_draws[ifo.name].T)
        if self.time_marginalization:
            times = self._times
            if self.jitter_time:
                times = times + parameters["time_jitter"]
            tmask = (times >= self.time_prior.minimum) & (times <= self.time_prior.maximum)
            arr = d_arr[:, tmask]
            time_prior_array = self.time_prior.prob(times[tmask]) * self._delta_tc
            if self.phase_marginalization:
                log_l = ln_i0(abs(arr)) - hh_arr[:, np.newaxis] / 2
            else:
                log_l = arr.real - hh_arr[:, np.newaxis] / 2
            return logsumexp(log_l, b=time_prior_array, axis=-1)
        if self.distance_marginalization:
            return self.distance_marginalized_likelihood(d_arr, hh_arr, parameters)
        if self.phase_marginalization:
            return ln_i0(abs(d_arr)) - hh_arr / 2
        return np.real(d_arr - hh_arr / 2)

    def calibration_marginalized_likelihood(self, pols, parameters):
        """base.py:860-877."""
        return logsumexp(self.calibration_log_likelihoods(pols, parameters)) - np.log(self.number_of_response_curves)

    def generate_calibration_sample(self, pols, parameters, u):
        """base.py:544-578: index of a response curve drawn from the curves' posterior.  `u` is the unit-interval
        draw numpy's Generator.choice(n, p=post) makes: cdf = cumsum(p) / cumsum(p)[-1], searchsorted(u, 'right')."""
        parameters.pop("recalib_index", None)
        log_like = self.calibration_log_likelihoods(pols, parameters)
        post = np.exp(log_like - max(log_like))
        post /= np.sum(post)
        cdf = post.cumsum()
        cdf /= cdf[-1]
        return int(cdf.searchsorted(u, side="right"))

    def distance_marginalized_likelihood(self, d_inner_h, hh, parameters):
        """base.py:775-784, 879-885."""
        hh_ref = hh * parameters["luminosity_distance"] ** 2 / self._ref_dist ** 2.
        d_ref = d_inner_h * parameters["luminosity_distance"] / self._ref_dist
        d_ref = abs(d_ref) if self.phase_marginalization else np.real(d_ref)
        out = self.interp_dist(d_ref, hh_ref)
        return out if np.ndim(d_inner_h) else out[0]

    def time_marginalized_likelihood(self, d_inner_h_tc_array, hh, parameters):
        """base.py:794-820."""
        times = self._times
        if self.jitter_time:
            times = times + parameters["time_jitter"]
        tmask = (times >= self.time_prior.minimum) & (times <= self.time_prior.maximum)
        times = times[tmask]
        arr = d_inner_h_tc_array[tmask]
        time_prior_array = self.time_prior.prob(times) * self._delta_tc
        if self.distance_marginalization:
            log_l = self.distance_marginalized_likelihood(arr, hh, parameters)
        elif self.phase_marginalization:
            log_l = ln_i0(abs(arr)) - hh / 2
        else:
            log_l = arr.real - hh / 2
        return logsumexp(log_l, b=time_prior_array, axis=-1)

    # ---- marginalised-parameter reconstruction (base.py:502-773); `uniforms` = the unit-interval draws that
    #      Interped.sample() (core/prior/base.py:143-164) would make, in the order time, distance, phase
    def generate_posterior_sample_from_marginalized_likelihood(self, parameters, uniforms):
        """base.py:502-541."""
        parameters = dict(parameters)
        if not (self.time_marginalization or self.distance_marginalization or self.phase_marginalization
                or self.calibration_marginalization):
            return parameters
        pols = {k: v.copy() for k, v in self.polarizations(parameters).items()}
        uniforms = list(uniforms)
        if self.calibration_marginalization:        # base.py:526-529; its draw is uniforms[3]
            parameters["recalib_index"] = self.generate_calibration_sample(pols, parameters, uniforms[3])
        if self.time_marginalization:
            parameters["geocent_time"] = self.generate_time_sample(pols, parameters, uniforms[0])
        if self.distance_marginalization:
            parameters["luminosity_distance"] = self.generate_distance_sample(pols, parameters, uniforms[1])
        if self.phase_marginalization:
            parameters["phase"] = self.generate_phase_sample(pols, parameters, uniforms[2])
        return parameters

    def _inner_products(self, pols, parameters):
        """base.py:715-725."""
        d_inner_h, hh = 0j, 0.0
        for ifo in self.ifos:
            signal = ifo.get_detector_response(pols, parameters)
            if "recalib_index" in parameters:        # base.py:289-290
                signal[ifo.frequency_mask] *= self.calibration_draws[ifo.name][int(parameters["recalib_index"])]
            d_inner_h += ifo.inner_product(signal)
            hh += ifo.optimal_snr_squared(signal)
        return d_inner_h, hh

    def generate_time_sample(self, pols, parameters, u):
        """base.py:578-658: the 16384 Hz zero-padded FFT of h conj(d) / S (NO 4/T factor in the reference), point
        likelihood per time, prior, > max/1000 cut, Interped sample."""
        if self.jitter_time:
            parameters["geocent_time"] = parameters["geocent_time"] + parameters["time_jitter"]
        n_time_steps = int(self.duration * 16384)
        times = np.linspace(parameters["geocent_time"] - self.start_time,
                            self.duration + (parameters["geocent_time"] - self.start_time) - 1 / 16384,
                            num=n_time_steps)                         # core/utils/series.py:91-112
        times = times % self.duration
        times += self.start_time
        prior = self.time_prior
        in_prior = (times >= prior.minimum) & (times < prior.maximum)
        times = times[in_prior]
        d_inner_h = np.zeros(len(times), dtype=complex)
        psd = np.ones(n_time_steps)
        signal_long = np.zeros(n_time_steps, dtype=complex)
        data = np.zeros(n_time_steps, dtype=complex)
        hh = np.zeros(1)
        for ifo in self.ifos:
            n_ifo = len(ifo.frequency_domain_strain)
            mask = ifo.frequency_mask
            signal = ifo.get_detector_response(pols, parameters)
            signal_long[:n_ifo] = signal
            data[:n_ifo] = np.conj(ifo.frequency_domain_strain)
            psd[:n_ifo][mask] = ifo.power_spectral_density_array[mask]
            d_inner_h += np.fft.fft(signal_long * data / psd)[in_prior]
            hh += ifo.optimal_snr_squared(signal).real
        if self.distance_marginalization:
            time_log_like = self.distance_marginalized_likelihood(d_inner_h, hh, parameters)
        elif self.phase_marginalization:
            time_log_like = ln_i0(abs(d_inner_h)) - hh.real / 2
        else:
            time_log_like = d_inner_h.real - hh.real / 2
        time_post = np.exp(time_log_like - max(time_log_like)) * prior.prob(times)
        keep = time_post > max(time_post) / 1000
        if sum(keep) < 3:
            keep[1:-1] = keep[1:-1] | keep[2:] | keep[:-2]
        return interped_sample(times[keep], time_post[keep], u)

    def generate_distance_sample(self, pols, parameters, u):
        """base.py:660-713 (incl. the in-place rescaling of the polarisations, :711 -> _rescale_signal :1020-1025)."""
        d_inner_h, hh = self._inner_products(pols, parameters)
        d_dist = d_inner_h * parameters["luminosity_distance"] / self._distance_array
        hh_dist = hh * parameters["luminosity_distance"] ** 2 / self._distance_array ** 2
        if self.phase_marginalization:
            log_like = ln_i0(abs(d_dist)) - hh_dist.real / 2
        else:
            log_like = d_dist.real - hh_dist.real / 2
        post = np.exp(log_like - max(log_like)) * self.distance_prior_array
        new_distance = interped_sample(self._distance_array, post, u)
        for mode in pols:
            pols[mode] *= self._ref_dist / new_distance
        return new_distance

    def generate_phase_sample(self, pols, parameters, u):
        """base.py:746-773."""
        d_inner_h, hh = self._inner_products(pols, parameters)
        phases = np.linspace(0, 2 * np.pi, 101)
        phasor = np.exp(-2j * phases)
        log_post = d_inner_h * phasor - hh / 2
        post = np.exp(log_post.real - max(log_post.real))
        return interped_sample(phases, post, u)

    def noise_log_likelihood(self):
        """base.py:402-417."""
        log_l = 0.0
        for ifo in self.ifos:
            m = ifo.frequency_mask
            log_l -= abs(4 / self.duration * np.sum(
                ifo.frequency_domain_strain[m].conj() * ifo.frequency_domain_strain[m]
                / ifo.power_spectral_density_array[m]) / 2)
        return log_l


# --------------------------------------------------------------------------------------
# benchmark set-up shared by tests and bench.py's CPU leg
# --------------------------------------------------------------------------------------
INJECTION = dict(mass_1=36.0, mass_2=29.0, chi_1=0.4, chi_2=0.3, luminosity_distance=2000.0,
                 theta_jn=0.4, psi=2.659, phase=1.3, geocent_time=1126259642.413, ra=1.375, dec=-1.2108)


def draw_bbh_prior(n, rng, t_inj=INJECTION["geocent_time"]):
    """Prior draws of SURVEY.md section 8d (dict of arrays, sampled parameterisation)."""
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


# --------------------------------------------------------------------------------------
# detector-based sky frame / detector time reference (base.py:1091-1137)
# --------------------------------------------------------------------------------------
def rotation_matrix_from_delta(delta_x):
    """bilby/gw/geometry.py:215-258."""
    delta_x = np.asarray(delta_x, dtype=float)
    delta_x = delta_x / (delta_x ** 2).sum() ** 0.5
    alpha = np.arctan2(-delta_x[1] * delta_x[2], delta_x[0])
    beta = np.arccos(delta_x[2])
    gamma = np.arctan2(delta_x[1], delta_x[0])
    r1 = np.array([[np.cos(alpha), -np.sin(alpha), 0], [np.sin(alpha), np.cos(alpha), 0], [0, 0, 1]])
    r2 = np.array([[np.cos(beta), 0, np.sin(beta)], [0, 1, 0], [-np.sin(beta), 0, np.cos(beta)]])
    r3 = np.array([[np.cos(gamma), -np.sin(gamma), 0], [np.sin(gamma), np.cos(gamma), 0], [0, 0, 1]])
    return r3 @ r2 @ r1


def zenith_azimuth_to_ra_dec(zenith, azimuth, time, vertex_1, vertex_2):
    """bilby/gw/utils.py:232-256 with geometry.py:346-377 (zenith_azimuth_to_theta_phi for a detector pair,
    delta_x = vertex_1 - vertex_2, detector/networks.py:487-490) and theta_phi_to_ra_dec."""
    omega_prime = np.array([np.sin(zenith) * np.cos(azimuth), np.sin(zenith) * np.sin(azimuth), np.cos(zenith)])
    omega = rotation_matrix_from_delta(np.asarray(vertex_1) - np.asarray(vertex_2)) @ omega_prime
    theta = np.arccos(omega[2])
    phi = np.arctan2(omega[1], omega[0]) % (2 * np.pi)
    gmst = greenwich_mean_sidereal_time(time)
    ra = (phi + gmst) % (2 * np.pi)
    dec = np.pi / 2 - theta
    return ra, dec


def get_sky_frame_parameters(parameters, frame_vertices=None, time_reference_vertex=None, time_key="geocent_time"):
    """base.py:1091-1137.  frame_vertices = (vertex_1, vertex_2) of the reference detector pair or None for the
    sky frame; time_reference_vertex = vertex of the time-reference detector or None for the geocentre."""
    time = parameters[time_key]
    if frame_vertices is not None:
        ra, dec = zenith_azimuth_to_ra_dec(parameters["zenith"], parameters["azimuth"], time, *frame_vertices)
    else:
        ra, dec = parameters["ra"], parameters["dec"]
    if time_reference_vertex is not None:
        geocent_time = time - time_delay_from_geocenter(time_reference_vertex, ra, dec, time)
    else:
        geocent_time = parameters["geocent_time"]
    return dict(ra=ra, dec=dec, geocent_time=geocent_time)
