"""GravitationalWaveTransient on B200: the reference's likelihood API over the fused CUDA path.

Mirrors bilby/gw/likelihood/base.py:
  constructor + prior side effects            :150-229 (geocent_time -> start_time, time_jitter prior,
                                               phase -> 0, luminosity_distance -> d_ref)
  _check_set_duration_and_sampling_frequency  :239-258
  noise_log_likelihood                        :402-417
  log_likelihood_ratio                        :419-446  (scalar call == batch of one)
  distance marginalisation set-up             :894-934, 994-1018 (table built on the device)
  time marginalisation set-up                 :1027-1035
  get_sky_frame_parameters                    :1091-1137 (sky frame / geocentre time reference)
plus the batched entry point the reference lacks: ``log_likelihood_ratio_batch``.
"""
import copy
import ctypes
import os

import numpy as np

from .. import _lib
from ..core.likelihood import Likelihood
from ..core.prior import Prior, Uniform
from ..core.utils import logger
from . import _params
from .detector import InterferometerList
from .detector.calibration import CubicSpline


class DeviceNetwork:
    """Device-resident detector network: one bb_handle + helpers (interferometers -> tiles)."""

    def __init__(self, interferometers, device=None):
        import torch
        self.torch = torch
        self.ifos = list(interferometers)
        self.handle = _lib.Handle(device)
        self.lib = self.handle.lib
        self.device = torch.device("cuda", self.handle.device)
        self.version = None
        self.upload()

    @property
    def ptr(self):
        return self.handle.ptr

    def upload(self):
        ifos = self.ifos
        n_det = len(ifos)
        f = ifos[0].frequency_array
        self.n_freq = n_freq = len(f)
        self.n_det = n_det
        tens = np.ascontiguousarray(np.stack([ifo.detector_tensor.ravel() for ifo in ifos]), dtype=np.float64)
        vert = np.ascontiguousarray(np.stack([ifo.vertex for ifo in ifos]), dtype=np.float64)
        strain = np.zeros((n_det, n_freq, 2), dtype=np.float64)
        psd = np.zeros((n_det, n_freq), dtype=np.float64)
        mask = np.zeros((n_det, n_freq), dtype=np.uint8)
        for i, ifo in enumerate(ifos):
            s = ifo.strain_data._frequency_domain_strain
            if s is None:
                s = np.zeros(n_freq, dtype=complex)
            strain[i, :, 0] = s.real
            strain[i, :, 1] = s.imag
            psd[i] = ifo.power_spectral_density_array
            mask[i] = ifo.frequency_mask
        _lib.check(self.lib.bb_set_network(
            self.ptr, n_det, n_freq, float(ifos[0].duration), float(ifos[0].sampling_frequency),
            float(ifos[0].start_time), tens.ctypes.data, vert.ctypes.data, strain.ctypes.data, psd.ctypes.data,
            mask.ctypes.data))

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _rows(self, **kw):
        rows = np.zeros((1, _params.NPARAM))
        rows[0, _params.MASS_1] = rows[0, _params.MASS_2] = 1.0
        rows[0, _params.LUMINOSITY_DISTANCE] = 1.0
        for k, v in kw.items():
            rows[0, getattr(_params, k.upper())] = v
        return self.torch.from_numpy(rows).to(self.device)

    def antenna(self, ra, dec, time, psi):
        rows = self._rows(ra=ra, dec=dec, geocent_time=time, psi=psi)
        out = self.torch.empty((1, self.n_det, 3), dtype=self.torch.float64, device=self.device)
        _lib.check(self.lib.bb_antenna_response_device(self.ptr, rows.data_ptr(), 1, out.data_ptr(), self._stream()))
        return out[0].cpu().numpy()

    def _complex_to_dev(self, arr):
        arr = np.asarray(arr, dtype=complex)
        buf = np.empty((len(arr), 2))
        buf[:, 0] = arr.real
        buf[:, 1] = arr.imag
        return self.torch.from_numpy(buf).to(self.device)

    def project(self, pols, parameters, det):
        unknown = set(pols) - {"plus", "cross"}
        if unknown:
            raise NotImplementedError(f"polarisation modes {sorted(unknown)} are outside the hot path")
        zero = np.zeros(self.n_freq, dtype=complex)
        plus = self._complex_to_dev(pols.get("plus", zero))
        cross = self._complex_to_dev(pols.get("cross", zero))
        ifo = self.ifos[det]
        if isinstance(ifo.calibration_model, CubicSpline):
            raise NotImplementedError("calibration of injected signals is not supported")
        rows = self._rows(ra=parameters["ra"], dec=parameters["dec"], psi=parameters["psi"],
                          geocent_time=parameters["geocent_time"])
        if ifo.reference_time is not None:
            raise NotImplementedError("Interferometer.reference_time is not supported on the device path")
        out = self.torch.empty((self.n_freq, 2), dtype=self.torch.float64, device=self.device)
        _lib.check(self.lib.bb_project_polarizations_device(self.ptr, det, plus.data_ptr(), cross.data_ptr(),
                                                            rows.data_ptr(), out.data_ptr(), self._stream()))
        o = out.cpu().numpy()
        return o[:, 0] + 1j * o[:, 1]

    def inner_product_arrays(self, det, a, b):
        a_d = self._complex_to_dev(a)
        b_d = self._complex_to_dev(b) if b is not None else None
        out = self.torch.empty(2, dtype=self.torch.float64, device=self.device)
        _lib.check(self.lib.bb_noise_weighted_inner_product_device(
            self.ptr, det, a_d.data_ptr(), b_d.data_ptr() if b_d is not None else None, out.data_ptr(),
            self._stream()))
        o = out.cpu().numpy()
        return complex(o[0], o[1])


class GravitationalWaveTransient(Likelihood):
    def __init__(self, interferometers, waveform_generator, time_marginalization=False,
                 distance_marginalization=False, phase_marginalization=False, calibration_marginalization=False,
                 priors=None, distance_marginalization_lookup_table=None, calibration_lookup_table=None,
                 number_of_response_curves=1000, starting_index=0, jitter_time=True, reference_frame="sky",
                 time_reference="geocenter", device=None):
        super().__init__()
        self.waveform_generator = waveform_generator
        self.interferometers = InterferometerList(interferometers)
        self.time_marginalization = time_marginalization
        self.distance_marginalization = distance_marginalization
        self.phase_marginalization = phase_marginalization
        self.calibration_marginalization = calibration_marginalization
        self.priors = priors
        self._check_set_duration_and_sampling_frequency_of_waveform_generator()
        self._noise_log_likelihood_value = None
        self.jitter_time = jitter_time
        self.reference_frame = reference_frame
        if "geocent" not in time_reference:            # base.py:171-181
            from .detector.networks import get_empty_interferometer
            self.time_reference = time_reference
            self.reference_ifo = get_empty_interferometer(self.time_reference)
            if self.time_marginalization:
                logger.info("Cannot marginalise over non-geocenter time.")
                self.time_marginalization = False
                self.jitter_time = False
        else:
            self.time_reference = "geocent"
            self.reference_ifo = None
        self._device_index = device
        self._net = None
        self._net_versions = None
        self._cal_points = 0
        # base.py:166, 183-229: self.priors is a COPY that keeps the Prior objects; the side effects below land on the
        # caller's dict (the one the sampler receives), exactly like the reference.

        if self.time_marginalization:
            self._check_marginalized_prior_is_set(key="geocent_time")
            self._check_time_prior_is_uniform()
            self._setup_time_marginalization()
            priors["geocent_time"] = float(self.interferometers.start_time)
            if self.jitter_time:
                priors["time_jitter"] = Uniform(minimum=-self._delta_tc / 2, maximum=self._delta_tc / 2,
                                                boundary="periodic", name="time_jitter", latex_label="$t_j$")
            self._marginalized_parameters.append("geocent_time")
        elif self.jitter_time:
            self.jitter_time = False

        if self.phase_marginalization:
            self._check_marginalized_prior_is_set(key="phase")
            priors["phase"] = float(0)
            self._marginalized_parameters.append("phase")

        if self.distance_marginalization:
            self._lookup_table_filename = None
            self._check_marginalized_prior_is_set(key="luminosity_distance")
            self._distance_array = np.linspace(self.priors["luminosity_distance"].minimum,
                                               self.priors["luminosity_distance"].maximum, int(1e4))
            self.distance_prior_array = np.array([self.priors["luminosity_distance"].prob(distance)
                                                  for distance in self._distance_array])
            self._ref_dist = self.priors["luminosity_distance"].rescale(0.5)
            self._setup_distance_marginalization(distance_marginalization_lookup_table)
            for key in ["redshift", "comoving_distance"]:
                if key in priors:
                    del priors[key]
            priors["luminosity_distance"] = float(self._ref_dist)
            self._marginalized_parameters.append("luminosity_distance")

        if self.calibration_marginalization:           # base.py:225-229
            if self.time_marginalization and self.distance_marginalization:
                # the reference itself fails here: distance_marginalized_likelihood (base.py:775-784) broadcasts the
                # [n_curves, n_times] array against the [n_curves] optimal SNRs
                raise ValueError("time + calibration + distance marginalisation is not defined by the reference "
                                 "(shape mismatch in base.py:775-784)")
            self.number_of_response_curves = number_of_response_curves
            self.starting_index = starting_index
            self._setup_calibration_marginalization(calibration_lookup_table, priors)
            self._marginalized_parameters.append("recalib_index")

    def _setup_calibration_marginalization(self, calibration_lookup_table, priors=None):
        """base.py:1037-1051."""
        from .detector import calibration
        from ..core.prior import DeltaFunction
        self.calibration_draws, self.calibration_parameter_draws = calibration.build_calibration_lookup(
            interferometers=self.interferometers, lookup_files=calibration_lookup_table, priors=priors,
            number_of_response_curves=self.number_of_response_curves, starting_index=self.starting_index)
        for name, parameters in self.calibration_parameter_draws.items():
            if parameters is not None and priors is not None:
                for key in set(parameters.keys()).intersection(priors.keys()):
                    priors[key] = DeltaFunction(0.0)
        self.calibration_abs_draws = {name: np.abs(c) ** 2 for name, c in self.calibration_draws.items()}

    def __repr__(self):
        return (f"{self.__class__.__name__}(interferometers={self.interferometers},\n\twaveform_generator="
                f"{self.waveform_generator},\n\ttime_marginalization={self.time_marginalization}, "
                f"distance_marginalization={self.distance_marginalization}, phase_marginalization="
                f"{self.phase_marginalization}, calibration_marginalization={self.calibration_marginalization}, "
                f"priors={self.priors})")

    # ---- set-up --------------------------------------------------------------------------------
    @property
    def priors(self):
        return self._prior

    @priors.setter
    def priors(self, priors):
        if priors is not None:
            self._prior = priors.copy()
        elif any([self.time_marginalization, self.phase_marginalization, self.distance_marginalization]):
            raise ValueError("You can't use a marginalized likelihood without specifying a priors")
        else:
            self._prior = None

    def _check_set_duration_and_sampling_frequency_of_waveform_generator(self):
        for attribute in ["duration", "sampling_frequency", "start_time"]:
            setattr(self.waveform_generator, attribute, getattr(self.interferometers, attribute))

    def _check_marginalized_prior_is_set(self, key):
        """base.py:356-388 (cosmological / default-BBH-prior branches need astropy: out of scope)."""
        if key in self.priors and getattr(self.priors[key], "is_fixed", False):
            raise ValueError("Cannot use marginalized likelihood for {}: prior is fixed".format(key))
        if key not in self.priors or not isinstance(self.priors[key], Prior):
            if key == "geocent_time":
                logger.warning("Prior not provided for geocent time, using the full segment.")
                self.priors[key] = Uniform(self.interferometers.start_time,
                                           self.interferometers.start_time + self.interferometers.duration)
            else:
                raise ValueError(f"Prior not provided for {key}: the reference would fall back to BBHPriorDict "
                                 "(astropy); supply the prior explicitly")

    def _check_time_prior_is_uniform(self):
        """The device applies the weight prior.prob(t) * delta_tc of base.py:794-806 as delta_tc / (max - min) inside
        [minimum, maximum]: exact for a Uniform prior only, so anything else is refused instead of mis-weighted."""
        prior = self.priors["geocent_time"]
        if not isinstance(prior, Uniform) or not (np.isfinite(prior.minimum) and np.isfinite(prior.maximum)):
            raise NotImplementedError(
                "time marginalisation on the device needs a Uniform geocent_time prior with finite bounds "
                f"(got {prior!r}); the reference weights by prior.prob(times) for any prior (base.py:794-806)")

    def _setup_time_marginalization(self):
        self._delta_tc = 2 / self.waveform_generator.sampling_frequency
        self._times = self.interferometers.start_time + np.linspace(
            0, self.interferometers.duration,
            int(self.interferometers.duration / 2 * self.waveform_generator.sampling_frequency + 1))[1:]
        self.time_prior_array = self.priors["geocent_time"].prob(self._times) * self._delta_tc

    @property
    def _delta_distance(self):
        return self._distance_array[1] - self._distance_array[0]

    @property
    def _optimal_snr_squared_ref_array(self):
        return np.logspace(-5, 10, self._dist_margd_loglikelihood_array.shape[0])

    @property
    def _d_inner_h_ref_array(self):
        if self.phase_marginalization:
            return np.logspace(-5, 10, self._dist_margd_loglikelihood_array.shape[1])
        n_negative = self._dist_margd_loglikelihood_array.shape[1] // 2
        n_positive = self._dist_margd_loglikelihood_array.shape[1] - n_negative
        return np.hstack((-np.logspace(3, -3, n_negative), np.logspace(-3, 10, n_positive)))

    @property
    def cached_lookup_table_filename(self):
        if self._lookup_table_filename is None:
            self._lookup_table_filename = ".distance_marginalization_lookup.npz"
        return self._lookup_table_filename

    @cached_lookup_table_filename.setter
    def cached_lookup_table_filename(self, filename):
        if isinstance(filename, str) and filename[-4:] != ".npz":
            filename += ".npz"
        self._lookup_table_filename = filename

    def _setup_distance_marginalization(self, lookup_table=None):
        """base.py:916-934: same .npz cache format as the reference (:979-992)."""
        table = None
        if isinstance(lookup_table, str) or lookup_table is None:
            self.cached_lookup_table_filename = lookup_table
            lookup_table = self.load_lookup_table(self.cached_lookup_table_filename)
        if isinstance(lookup_table, dict) and self._test_cached_lookup_table(lookup_table)[0]:
            table = lookup_table["lookup_table"]
        if table is None:
            self._create_lookup_table()
        else:
            self._dist_margd_loglikelihood_array = np.asarray(table)
        from scipy.interpolate import RectBivariateSpline
        x, y = self._d_inner_h_ref_array, self._optimal_snr_squared_ref_array
        spl = RectBivariateSpline(x, y, self._dist_margd_loglikelihood_array.T, kx=3, ky=3, s=0)
        tx, ty, c = spl.tck
        self._dist_spline = dict(tx=np.ascontiguousarray(tx), ty=np.ascontiguousarray(ty),
                                 c=np.ascontiguousarray(c), bbox=(x.min(), x.max(), y.min(), y.max()))

    def load_lookup_table(self, filename):
        if filename is not None and os.path.exists(filename):
            loaded = dict(np.load(filename))
            match, failure = self._test_cached_lookup_table(loaded)
            if match:
                return loaded
            logger.info("Loaded distance marginalisation lookup table does not match for {}.".format(failure))
        return None

    def cache_lookup_table(self):
        np.savez(self.cached_lookup_table_filename, distance_array=self._distance_array,
                 prior_array=self.distance_prior_array, lookup_table=self._dist_margd_loglikelihood_array,
                 reference_distance=self._ref_dist, phase_marginalization=self.phase_marginalization)

    def _test_cached_lookup_table(self, loaded_file):
        pairs = dict(distance_array=self._distance_array, prior_array=self.distance_prior_array,
                     reference_distance=self._ref_dist, phase_marginalization=self.phase_marginalization)
        for key in pairs:
            if key not in loaded_file:
                return False, key
            if not np.allclose(np.atleast_1d(loaded_file[key]), np.atleast_1d(pairs[key]), rtol=1e-15):
                return False, key
        return True, None

    def _create_lookup_table(self):
        """base.py:994-1018 on the device (bb_build_distance_table)."""
        self._dist_margd_loglikelihood_array = np.zeros((400, 800))
        x = np.ascontiguousarray(self._d_inner_h_ref_array)
        y = np.ascontiguousarray(self._optimal_snr_squared_ref_array)
        dist = np.ascontiguousarray(self._distance_array)
        prior = np.ascontiguousarray(self.distance_prior_array)
        table = np.zeros((400, 800))
        h = _lib.Handle(self._device_index)
        _lib.check(h.lib.bb_build_distance_table(h.ptr, x.ctypes.data, len(x), y.ctypes.data, len(y),
                                                 dist.ctypes.data, prior.ctypes.data, len(dist),
                                                 float(self._ref_dist), int(self.phase_marginalization),
                                                 table.ctypes.data))
        self._dist_margd_loglikelihood_array = table
        if self._lookup_table_filename is not None or True:
            try:
                self.cache_lookup_table()
            except OSError:  # pragma: no cover
                pass

    # ---- device state --------------------------------------------------------------------------
    def _versions(self):
        return tuple(ifo._data_version for ifo in self.interferometers)

    @property
    def device_network(self):
        if self._net is None or self._net_versions != self._versions():
            self._net = DeviceNetwork(self.interferometers, self._device_index)
            self._net_versions = self._versions()
            self._recon_grid_set = False
            self._configure()
        return self._net

    def _configure(self):
        net = self._net
        approx, f_ref, f_min, f_max = self.waveform_generator.approximant_config()
        _lib.check(net.lib.bb_set_waveform(net.ptr, approx, f_ref, f_min, f_max))
        # calibration model (interferometer.py:364 -> calibration.py:349-384)
        models = [ifo.calibration_model for ifo in self.interferometers]
        self._cal_points = 0
        if any(isinstance(m, CubicSpline) for m in models):
            if not all(isinstance(m, CubicSpline) for m in models) or len({m.n_points for m in models}) != 1:
                raise NotImplementedError("all interferometers must use CubicSpline models with equal n_points")
            npts = models[0].n_points
            lo = np.array([m.log_spline_points[0] for m in models], dtype=np.float64)
            hi = np.array([m.log_spline_points[-1] for m in models], dtype=np.float64)
            mat = np.ascontiguousarray(models[0].nodes_to_spline_coefficients, dtype=np.float64)
            _lib.check(net.lib.bb_set_calibration(net.ptr, npts, lo.ctypes.data, hi.ctypes.data, mat.ctypes.data))
            self._cal_points = npts
        else:
            _lib.check(net.lib.bb_set_calibration(net.ptr, 0, None, None, None))
        flags = 0
        if self.phase_marginalization:
            flags |= _params.MARG_PHASE
        if self.distance_marginalization:
            flags |= _params.MARG_DISTANCE
        if self.time_marginalization:
            flags |= _params.MARG_TIME
        tmin = tmax = 0.0
        if self.time_marginalization:
            tmin, tmax = float(self.priors["geocent_time"].minimum), float(self.priors["geocent_time"].maximum)
        if self.distance_marginalization:
            s = self._dist_spline
            _lib.check(net.lib.bb_set_marginalization(
                net.ptr, flags, float(self._ref_dist), s["tx"].ctypes.data, len(s["tx"]), s["ty"].ctypes.data,
                len(s["ty"]), s["c"].ctypes.data, *[float(v) for v in s["bbox"]], tmin, tmax,
                int(bool(self.jitter_time))))
        else:
            _lib.check(net.lib.bb_set_marginalization(net.ptr, flags, 1.0, None, 0, None, 0, None, 0.0, 0.0, 0.0,
                                                      0.0, tmin, tmax, int(bool(self.jitter_time))))

        if self.calibration_marginalization:
            # response curves on the full frequency grid, [n_det, n_curves, n_freq] complex
            ifos = self.interferometers
            n_freq = len(ifos[0].frequency_array)
            full = np.zeros((len(ifos), self.number_of_response_curves, n_freq), dtype=np.complex128)
            for d, ifo in enumerate(ifos):
                full[d][:, ifo.frequency_mask] = self.calibration_draws[ifo.name]
            buf = np.ascontiguousarray(np.stack([full.real, full.imag], axis=-1))
            _lib.check(net.lib.bb_set_calibration_marginalization(net.ptr, self.number_of_response_curves,
                                                                  buf.ctypes.data))
        else:
            _lib.check(net.lib.bb_set_calibration_marginalization(net.ptr, 0, None))

    # ---- reference frame (base.py:1063-1137) ------------------------------------------------------
    @property
    def reference_frame(self):
        return self._reference_frame

    @reference_frame.setter
    def reference_frame(self, frame):
        if isinstance(frame, str) and frame == "sky":
            self._reference_frame = frame
        elif isinstance(frame, InterferometerList):
            self._reference_frame = InterferometerList(list(frame)[:2])
        elif isinstance(frame, list):
            self._reference_frame = InterferometerList(frame[:2])
        elif isinstance(frame, str):
            self._reference_frame = InterferometerList([frame[:2], frame[2:4]])
        else:
            raise ValueError("Unable to parse reference frame {}".format(frame))

    @property
    def _reference_frame_str(self):
        if isinstance(self.reference_frame, str):
            return self.reference_frame
        return "".join([ifo.name for ifo in self.reference_frame])

    @staticmethod
    def _rotation_matrix_from_delta(delta_x):
        """bilby/gw/geometry.py:215-258 (set-up: one 3x3 matrix per likelihood)."""
        d = np.asarray(delta_x, dtype=float)
        d = d / np.sqrt((d ** 2).sum())
        alpha, beta, gamma = np.arctan2(-d[1] * d[2], d[0]), np.arccos(d[2]), np.arctan2(d[1], d[0])

        def rz(a):
            return np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
        ry = np.array([[np.cos(beta), 0.0, np.sin(beta)], [0.0, 1.0, 0.0], [-np.sin(beta), 0.0, np.cos(beta)]])
        return rz(gamma) @ ry @ rz(alpha)

    def _set_device_frame(self, use_sky_frame):
        """Tell the handle how to read the (RA, DEC, GEOCENT_TIME) columns (bb_set_reference_frame)."""
        net = self.device_network
        rot = vert = None
        if use_sky_frame:
            f = self.reference_frame
            rot = np.ascontiguousarray(self._rotation_matrix_from_delta(f[0].vertex - f[1].vertex), dtype=np.float64)
        if self.reference_ifo is not None:
            vert = np.ascontiguousarray(self.reference_ifo.vertex, dtype=np.float64)
        state = (None if rot is None else rot.tobytes(), None if vert is None else vert.tobytes())
        if getattr(net, "_frame_state", (None, None)) != state:
            _lib.check(net.lib.bb_set_reference_frame(net.ptr, None if rot is None else rot.ctypes.data,
                                                      None if vert is None else vert.ctypes.data))
            net._frame_state = state

    def _frame_columns(self, parameters):
        """Parameters as the device reads them: (ra|azimuth, dec|zenith, geocent_time|{IFO}_time).  Mirrors the
        fall-backs of base.py:1107-1136."""
        time_key = f"{self.time_reference}_time"
        out = dict(parameters)
        use_frame = False
        if self.reference_frame != "sky":
            if "zenith" in parameters and "azimuth" in parameters:
                out["ra"], out["dec"] = parameters["azimuth"], parameters["zenith"]
                use_frame = True
            elif "ra" in parameters and "dec" in parameters:
                logger.warning("Cannot convert from zenith/azimuth to ra/dec falling back to provided ra/dec")
            else:
                raise KeyError("zenith")
        if self.reference_ifo is not None:
            if time_key not in parameters:
                raise KeyError(time_key)
            out["geocent_time"] = parameters[time_key]
        self._set_device_frame(use_frame)
        return out

    # ---- evaluation ----------------------------------------------------------------------------
    def get_sky_frame_parameters(self, parameters):
        """base.py:1091-1137: {ra, dec, geocent_time} of one parameter dict (converted on the device)."""
        if self.reference_frame == "sky" and self.reference_ifo is None:
            return dict(ra=parameters["ra"], dec=parameters["dec"], geocent_time=parameters["geocent_time"])
        cols = self._frame_columns(parameters)
        net = self.device_network
        rows = net._rows(ra=float(cols["ra"]), dec=float(cols["dec"]), geocent_time=float(cols["geocent_time"]))
        out = net.torch.empty((1, 3), dtype=net.torch.float64, device=net.device)
        _lib.check(net.lib.bb_sky_frame_parameters_device(net.ptr, rows.data_ptr(), 1, out.data_ptr(), net._stream()))
        ra, dec, tg = out[0].cpu().numpy()
        return dict(ra=float(ra), dec=float(dec), geocent_time=float(tg))

    def _rows_from_parameters(self, parameters, n, xp, device=None):
        converted = self.waveform_generator.convert(self._frame_columns(parameters))
        return _params.pack_rows(converted, n, xp, device=device)

    def _cal_from_parameters(self, parameters, n, xp, device=None):
        """[n, n_det, 2, n_points] calibration parameters recalib_{IFO}_{amplitude,phase}_{i}
        (calibration.py:248-251 prefix convention), or None without a calibration model."""
        self.device_network
        npts = self._cal_points
        if not npts:
            return None
        n_det = len(self.interferometers)
        if xp is np:
            out = np.zeros((n, n_det, 2, npts), dtype=np.float64)
        else:
            out = xp.zeros((n, n_det, 2, npts), dtype=xp.float64, device=device)
        for d, ifo in enumerate(self.interferometers):
            for k, kind in enumerate(("amplitude", "phase")):
                for i in range(npts):
                    v = parameters[f"recalib_{ifo.name}_{kind}_{i}"]
                    out[:, d, k, i] = v if xp is np else xp.as_tensor(v, dtype=xp.float64, device=device)
        return out

    def log_likelihood_ratio(self, parameters):
        """base.py:419-446: one parameter dict in, one float out (a batch of one through the same kernels)."""
        parameters = copy.deepcopy(parameters)
        rows = np.ascontiguousarray(self._rows_from_parameters(parameters, 1, np))
        return float(self.log_likelihood_ratio_rows_host(rows, self._cal_from_parameters(parameters, 1, np))[0])

    def log_likelihood_ratio_batch(self, parameters):
        """NEW (no reference equivalent): evaluate a batch.

        parameters: dict of equal-length numpy arrays (host path: one pinned H2D copy, kernels, D2H copy;
        returns a numpy array) or dict of CUDA torch tensors / an [n, 16] CUDA float64 tensor of packed
        rows (device path: asynchronous on the current stream; returns a CUDA tensor)."""
        net = self.device_network
        torch = net.torch
        if isinstance(parameters, torch.Tensor):
            rows = parameters
            if rows.dtype != torch.float64 or rows.dim() != 2 or rows.shape[1] != _params.NPARAM or not rows.is_cuda:
                raise ValueError("packed rows must be a CUDA float64 tensor of shape [n, 16]")
            rows = rows.contiguous()
            if self.device_network and self._cal_points:
                raise ValueError("packed rows need calibration parameters: use _evaluate_device(rows, cal)")
            return self._evaluate_device(rows)
        n = None
        on_device = False
        for v in parameters.values():
            if isinstance(v, torch.Tensor):
                on_device = on_device or v.is_cuda
                if v.dim() > 0:
                    n = v.shape[0]
            elif isinstance(v, np.ndarray) and v.ndim > 0:
                n = v.shape[0]
        if n is None:
            raise ValueError("log_likelihood_ratio_batch needs at least one array-valued parameter")
        if on_device:
            rows = self._rows_from_parameters(parameters, n, torch, device=net.device)
            return self._evaluate_device(rows, self._cal_from_parameters(parameters, n, torch, device=net.device))
        rows = np.ascontiguousarray(self._rows_from_parameters(parameters, n, np))
        return self.log_likelihood_ratio_rows_host(rows, self._cal_from_parameters(parameters, n, np))

    def log_likelihood_ratio_rows_host(self, rows, cal=None, out=None):
        """Host rows [n,16] (numpy float64, C order) [+ calibration parameters [n,n_det,2,n_points]] -> numpy
        lnL; the C ABI's end-to-end entry point.  Page-locked buffers (`bilby_b200.core.utils.pinned_empty`, also
        for `out`) are copied from / to directly; pageable ones pass through the library's staging buffers
        (one extra host copy: 128 B per sample, which bounds the reduced-order likelihoods)."""
        net = self.device_network
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        if out is None:
            out = np.empty(rows.shape[0])
        elif out.dtype != np.float64 or not out.flags.c_contiguous or out.shape != (rows.shape[0],):
            raise ValueError("out must be a C-contiguous float64 array of length n")
        if self._cal_points:
            if cal is None:
                raise ValueError("this likelihood has a calibration model: calibration parameters are required")
            cal = np.ascontiguousarray(cal, dtype=np.float64)
            _lib.check(net.lib.bb_log_likelihood_ratio_cal_host(net.ptr, rows.ctypes.data, cal.ctypes.data,
                                                                rows.shape[0], out.ctypes.data))
        else:
            _lib.check(net.lib.bb_log_likelihood_ratio_host(net.ptr, rows.ctypes.data, rows.shape[0], out.ctypes.data))
        return out

    def _evaluate_device(self, rows, cal=None):
        """Device rows in, device lnL out, through the TORCH_LIBRARY ops (csrc/bb_torch.cpp -> C ABI) on the current
        stream of the rows' device."""
        net = self.device_network
        ops = _lib.torch_ops()
        if self._cal_points:
            if cal is None:
                raise ValueError("this likelihood has a calibration model: calibration parameters are required")
            return ops.log_likelihood_ratio_cal(net.ptr.value, rows.contiguous(), cal.contiguous())
        return ops.log_likelihood_ratio(net.ptr.value, rows.contiguous())

    def inner_products_batch(self, rows, cal=None):
        """[n,16] CUDA rows -> [n, n_det, 3] (Re<h|d>, Im<h|d>, <h|h>) per detector (base.py:260-354)."""
        net = self.device_network
        torch = net.torch
        out = torch.empty((rows.shape[0], net.n_det, 3), dtype=torch.float64, device=rows.device)
        if self._cal_points:
            if cal is None:
                raise ValueError("this likelihood has a calibration model: calibration parameters are required")
            cal = cal.contiguous()
            _lib.check(net.lib.bb_inner_products_cal_device(net.ptr, rows.data_ptr(), cal.data_ptr(), rows.shape[0],
                                                            out.data_ptr(), net._stream()))
        else:
            _lib.check(net.lib.bb_inner_products_device(net.ptr, rows.data_ptr(), rows.shape[0], out.data_ptr(),
                                                        net._stream()))
        return out

    def likelihood_from_inner_products(self, rows, snrs):
        net = self.device_network
        torch = net.torch
        out = torch.empty(rows.shape[0], dtype=torch.float64, device=rows.device)
        _lib.check(net.lib.bb_likelihood_from_inner_products_device(
            net.ptr, rows.data_ptr(), snrs.data_ptr(), rows.shape[0], out.data_ptr(), net._stream()))
        return out

    def pack(self, parameters, n=None, device=None):
        """dict of arrays / tensors -> packed rows (numpy, or CUDA tensor if device is given)."""
        if device is None:
            if n is None:
                n = max(np.size(v) for v in parameters.values())
            return self._rows_from_parameters(parameters, n, np)
        torch = self.device_network.torch
        if n is None:
            n = max((v.shape[0] for v in parameters.values() if hasattr(v, "shape") and len(v.shape)), default=1)
        return self._rows_from_parameters(parameters, n, torch, device=device)

    def calculate_snrs(self, parameters):
        """Per-detector (d_inner_h, optimal_snr_squared, complex_matched_filter_snr) for one dict
        (base.py:260-354, scalar quantities)."""
        net = self.device_network
        rows = net.torch.from_numpy(np.ascontiguousarray(self._rows_from_parameters(parameters, 1, np))).to(net.device)
        cal = self._cal_from_parameters(parameters, 1, np)
        cal = None if cal is None else net.torch.from_numpy(np.ascontiguousarray(cal)).to(net.device)
        s = self.inner_products_batch(rows, cal)[0].cpu().numpy()
        out = []
        for d in range(net.n_det):
            dih = complex(s[d, 0], s[d, 1])
            out.append(dict(d_inner_h=dih, optimal_snr_squared=s[d, 2],
                            complex_matched_filter_snr=dih / s[d, 2] ** 0.5))
        return out

    def compute_snrs_batch(self, parameters):
        """Per-detector matched-filter and optimal SNRs for a batch of samples: the batched form of
        bilby.gw.conversion.compute_snrs (conversion.py:2215-2288 -> base.py:260-300).  Returns
        ``{"<IFO>_matched_filter_snr": complex[n], "<IFO>_optimal_snr": float[n]}``."""
        net = self.device_network
        torch = net.torch
        n = max(np.size(v) for v in parameters.values())
        rows = torch.from_numpy(np.ascontiguousarray(self._rows_from_parameters(parameters, n, np))).to(net.device)
        cal = self._cal_from_parameters(parameters, n, np)
        cal = None if cal is None else torch.from_numpy(np.ascontiguousarray(cal)).to(net.device)
        s = self.inner_products_batch(rows, cal).cpu().numpy()
        out = {}
        for d, ifo in enumerate(self.interferometers):
            dih = s[:, d, 0] + 1j * s[:, d, 1]
            out[f"{ifo.name}_matched_filter_snr"] = dih / s[:, d, 2] ** 0.5
            out[f"{ifo.name}_optimal_snr"] = s[:, d, 2] ** 0.5
        return out

    # ---- marginalised-parameter reconstruction (base.py:502-773) ----------------------------------------------
    def generate_posterior_samples_from_marginalized_likelihood_batch(self, parameters, uniforms=None, rng=None):
        """Batched generate_posterior_sample_from_marginalized_likelihood (base.py:502-541): dict of arrays in, the
        same dict with new ``geocent_time`` / ``luminosity_distance`` / ``phase`` columns out (only for the
        marginalisations that are on; the steps run in the reference's order and each sees the values drawn
        before it).

        uniforms: [n, 3] unit-interval draws standing for the ``Interped.sample()`` calls of the time, distance
        and phase steps; drawn from ``rng`` (default: a fresh numpy Generator) when omitted.

        With calibration marginalisation (base.py:526-529, 544-578; time marginalisation excluded) column 0 of
        ``uniforms`` is the draw of ``rng.choice`` over the response curves, the result gains ``recalib_index`` and
        the distance / phase steps use the chosen curve (base.py:289-290)."""
        if not self._marginalized_parameters:
            return dict(parameters)
        calmarg = bool(getattr(self, "calibration_marginalization", False))
        if calmarg and self.time_marginalization:
            raise NotImplementedError("time + calibration marginalisation is not built (SURVEY.md section 8f rank 4)")
        net = self.device_network
        torch = net.torch
        n = max(np.size(v) for v in parameters.values())
        if uniforms is None:
            if rng is None:
                from ..core.utils import random
                rng = random.rng
            uniforms = rng.uniform(0, 1, size=(n, 3))
        uniforms = np.array(np.broadcast_to(np.asarray(uniforms, dtype=np.float64), (n, 3)), order="C")
        pars = dict(parameters)
        if self.time_marginalization and "time_jitter" not in pars:
            pars["time_jitter"] = np.zeros(n)
        rows = torch.from_numpy(np.ascontiguousarray(self._rows_from_parameters(pars, n, np))).to(net.device)
        pars.pop("recalib_index", None)              # base.py:562-563
        cal = None if calmarg else self._cal_from_parameters(pars, n, np)
        cal_ptr = None
        if cal is not None:
            cal = torch.from_numpy(np.ascontiguousarray(cal)).to(net.device)
            cal_ptr = cal.data_ptr()
        if self.distance_marginalization and not getattr(self, "_recon_grid_set", False):
            dist = np.ascontiguousarray(self._distance_array, dtype=np.float64)
            prior = np.ascontiguousarray(self.distance_prior_array, dtype=np.float64)
            _lib.check(net.lib.bb_set_reconstruction_grid(net.ptr, dist.ctypes.data, prior.ctypes.data, len(dist)))
            self._recon_grid_set = True
        u_dev = torch.from_numpy(uniforms).to(net.device)
        out = torch.full((n, 3), float("nan"), dtype=torch.float64, device=net.device)
        _lib.check(net.lib.bb_reconstruct_marginalized_device(net.ptr, rows.data_ptr(), cal_ptr, n, u_dev.data_ptr(),
                                                              out.data_ptr(), net._stream()))
        res = out.cpu().numpy()
        new = {k: (np.array(v, copy=True) if np.ndim(v) else v) for k, v in parameters.items()}
        if calmarg:
            new["recalib_index"] = res[:, 0]
        if self.time_marginalization:
            new["geocent_time"] = res[:, 0]
        if self.distance_marginalization:
            new["luminosity_distance"] = res[:, 1]
        if self.phase_marginalization:
            new["phase"] = res[:, 2]
        return new

    def generate_posterior_sample_from_marginalized_likelihood(self, parameters, rng=None):
        """base.py:502-541 for one parameter dict (a batch of one through the same kernels)."""
        if not self._marginalized_parameters:
            return parameters
        one = {k: np.atleast_1d(np.asarray(v, dtype=np.float64)) for k, v in parameters.items()
               if np.isscalar(v) or np.ndim(v) == 0}
        new = self.generate_posterior_samples_from_marginalized_likelihood_batch(one, rng=rng)
        out = dict(parameters)
        for k in ("geocent_time", "luminosity_distance", "phase"):
            if k in new:
                out[k] = float(np.asarray(new[k])[0])
        if "recalib_index" in new and getattr(self, "calibration_marginalization", False):
            out["recalib_index"] = int(np.asarray(new["recalib_index"])[0])
        return out

    def _calculate_noise_log_likelihood(self):
        """base.py:402-411 through the device inner-product kernel."""
        log_l = 0.0
        net = self.device_network
        for i, ifo in enumerate(self.interferometers):
            d = ifo.frequency_domain_strain
            log_l -= abs(net.inner_product_arrays(i, d, None) / 2)
        return log_l

    def noise_log_likelihood(self):
        if self._noise_log_likelihood_value is None:
            self._noise_log_likelihood_value = self._calculate_noise_log_likelihood()
        return self._noise_log_likelihood_value

    def log_likelihood(self, parameters):
        return self.log_likelihood_ratio(parameters=parameters) + self.noise_log_likelihood()

    @property
    def meta_data(self):
        """base.py:1163-1183."""
        return dict(interferometers=self.interferometers.meta_data,
                    time_marginalization=self.time_marginalization,
                    phase_marginalization=self.phase_marginalization,
                    distance_marginalization=self.distance_marginalization,
                    calibration_marginalization=self.calibration_marginalization,
                    waveform_generator_class=self.waveform_generator.__class__,
                    waveform_arguments=self.waveform_generator.waveform_arguments,
                    frequency_domain_source_model=self.waveform_generator.frequency_domain_source_model,
                    time_domain_source_model=None,
                    parameter_conversion=self.waveform_generator.parameter_conversion,
                    sampling_frequency=self.waveform_generator.sampling_frequency,
                    duration=self.waveform_generator.duration,
                    start_time=self.waveform_generator.start_time,
                    time_reference=self.time_reference,
                    reference_frame=self.reference_frame)

    @meta_data.setter
    def meta_data(self, value):
        pass


from .relative import RelativeBinningGravitationalWaveTransient  # noqa: E402,F401
from .roq import ROQGravitationalWaveTransient, BilbyROQParamsRangeError  # noqa: E402,F401
from .multiband import MBGravitationalWaveTransient  # noqa: E402,F401
