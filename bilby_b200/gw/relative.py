"""RelativeBinningGravitationalWaveTransient on B200 (bilby/gw/likelihood/relative.py:26-452).

Set-up (once per fiducial point, host numpy like the reference): fiducial waveforms on the full grid (computed by
the device kernels), bin edges (relative.py:179-240), summary data (relative.py:319-363).  Evaluation: kernel K5
(`bb_relbin_kernel`): waveform at the bin edges, ratio to the fiducial, <d|h>, <h|h>; with time marginalisation
K5t rebuilds the full-grid series from the per-bin ratios and runs the same in-shared-memory FFT as K4.
"""
import copy

import numpy as np

from .. import _lib
from ..core.utils import logger
from .detector.calibration import CubicSpline
from .likelihood import GravitationalWaveTransient


class RelativeBinningGravitationalWaveTransient(GravitationalWaveTransient):
    def __init__(self, interferometers, waveform_generator, fiducial_parameters=None, parameter_bounds=None,
                 maximization_kwargs=None, update_fiducial_parameters=False, distance_marginalization=False,
                 time_marginalization=False, phase_marginalization=False, priors=None,
                 distance_marginalization_lookup_table=None, jitter_time=True, reference_frame="sky",
                 time_reference="geocenter", chi=1, epsilon=0.5, device=None):
        if getattr(waveform_generator.frequency_domain_source_model, "_bb_kind", None) != "relative_binning":
            raise TypeError("RelativeBinningGravitationalWaveTransient needs one of the source models "
                            "lal_binary_black_hole_relative_binning / lal_binary_neutron_star_relative_binning")
        self._rb_host = None
        super().__init__(interferometers=interferometers, waveform_generator=waveform_generator,
                         distance_marginalization=distance_marginalization,
                         phase_marginalization=phase_marginalization, time_marginalization=time_marginalization,
                         priors=priors, distance_marginalization_lookup_table=distance_marginalization_lookup_table,
                         jitter_time=jitter_time, reference_frame=reference_frame, time_reference=time_reference,
                         device=device)
        if fiducial_parameters is None:
            logger.info("Drawing fiducial parameters from prior.")
            fiducial_parameters = priors.sample()
        self.fiducial_parameters = fiducial_parameters.copy()
        self._unset_fiducial()
        if self.time_marginalization:
            self.fiducial_parameters["geocent_time"] = self.interferometers.start_time
        if self.distance_marginalization:
            self.fiducial_parameters["luminosity_distance"] = self._ref_dist
        if self.phase_marginalization:
            self.fiducial_parameters["phase"] = 0.0
        self.chi = chi
        self.epsilon = epsilon
        self.gamma = np.array([-5 / 3, -2 / 3, 1, 5 / 3, 7 / 3])
        self.maximum_frequency = waveform_generator.frequency_array[-1]
        self.fiducial_polarizations = None
        self.per_detector_fiducial_waveforms = dict()
        self.per_detector_fiducial_waveform_points = dict()
        self._setup_all(self.fiducial_parameters)
        if update_fiducial_parameters:
            from ..core.prior import Prior
            self.parameters_to_be_updated = [key for key in priors if isinstance(priors[key], Prior)
                                             and not getattr(priors[key], "is_fixed", False)]
            if parameter_bounds is None:
                self.parameter_bounds = self.get_bounds_from_priors(priors)
            else:
                self.parameter_bounds = self.get_parameter_list_from_dictionary(parameter_bounds)
            self.fiducial_parameters = self.find_maximum_likelihood_parameters(
                self.parameter_bounds, maximization_kwargs=maximization_kwargs)
        logger.info(f"Fiducial likelihood: {self.log_likelihood_ratio(self.fiducial_parameters):.2f}")

    def __repr__(self):
        return (f"{self.__class__.__name__}(interferometers={self.interferometers},\n\twaveform_generator="
                f"{self.waveform_generator},\n\tfiducial_parameters={self.fiducial_parameters})")

    # ---- set-up --------------------------------------------------------------------------------
    def _setup_all(self, parameters):
        self.set_fiducial_waveforms(parameters)
        self.setup_bins()
        self.compute_summary_data()
        self._rb_host = self._pack_host_arrays()
        self._upload_relative_binning()

    def _set_fiducial(self):
        self.waveform_generator.waveform_arguments["fiducial"] = 1
        self.waveform_generator._cache["parameters"] = None

    def _unset_fiducial(self):
        self.waveform_generator.waveform_arguments["fiducial"] = 0
        self.waveform_generator._cache["parameters"] = None

    def set_fiducial_waveforms(self, parameters):
        """relative.py:242-263: full-grid fiducial polarisations and per-detector responses (device kernels)."""
        parameters = parameters.copy()
        parameters.update(self.get_sky_frame_parameters(parameters))
        self._set_fiducial()
        try:
            self.fiducial_polarizations = self.waveform_generator.frequency_domain_strain(parameters)
        finally:
            self._unset_fiducial()
        if self.fiducial_polarizations is None:
            raise ValueError(f"Cannot compute fiducial waveforms for {parameters}")
        frequency_array = self.waveform_generator.frequency_array
        last = np.flatnonzero(self.fiducial_polarizations["plus"] != 0j)[-1]
        self.maximum_frequency = frequency_array[last]
        net = self.device_network
        _lib.check(net.lib.bb_set_relative_binning(net.ptr, 0, None, None, None, None, None))     # full-grid kernels
        torch = net.torch
        rows = torch.from_numpy(np.ascontiguousarray(self._rows_from_parameters(parameters, 1, np))).to(net.device)
        out = torch.empty((1, net.n_det, net.n_freq, 2), dtype=torch.float64, device=net.device)
        _lib.check(net.lib.bb_detector_response_device(net.ptr, rows.data_ptr(), 1, out.data_ptr(), net._stream()))
        resp = out[0].cpu().numpy()
        for d, ifo in enumerate(self.interferometers):
            wf = resp[d, :, 0] + 1j * resp[d, :, 1]
            if isinstance(ifo.calibration_model, CubicSpline):
                wf = wf * ifo.calibration_model.get_calibration_factor(frequency_array,
                                                                       prefix=f"recalib_{ifo.name}_", **parameters)
            wf = wf * (ifo.frequency_array <= self.maximum_frequency)
            self.per_detector_fiducial_waveforms[ifo.name] = wf

    def setup_bins(self):
        """relative.py:179-240 (Zackay et al., arXiv:1806.08792): bins of equal accumulated phase budget."""
        frequency_array = self.waveform_generator.frequency_array
        gamma = self.gamma[:, np.newaxis]
        fmax = max(float(ifo.maximum_frequency) for ifo in self.interferometers)
        fmin = min(float(ifo.minimum_frequency) for ifo in self.interferometers)
        fmax = min(max(fmax, frequency_array[0]), self.maximum_frequency)
        fmin = min(fmin, frequency_array[-1])
        useful = frequency_array[(frequency_array >= fmin) & (frequency_array <= fmax)]
        d_alpha = self.chi * 2 * np.pi / np.abs((fmin ** gamma) * np.heaviside(-gamma, 1)
                                                - (fmax ** gamma) * np.heaviside(gamma, 1))
        d_phi = np.sum(np.sign(gamma) * d_alpha * useful ** gamma, axis=0)
        budget = d_phi - d_phi[0]
        number_of_bins = int(budget[-1] // self.epsilon)
        # first grid point at or beyond each equally spaced phase target, duplicates dropped
        targets = (np.arange(number_of_bins + 1) / number_of_bins) * budget[-1]
        first = np.array([np.flatnonzero(budget >= t)[0] for t in targets])
        keep = np.concatenate(([True], first[1:] != first[:-1]))
        self.bin_freqs = useful[first[keep]]
        self.bin_inds = np.searchsorted(frequency_array, self.bin_freqs, side="left").astype(int)
        self.bin_sizes = np.diff(self.bin_inds)
        self.bin_sizes[-1] += 1
        self.number_of_bins = len(self.bin_inds) - 1
        logger.debug(f"Set up {self.number_of_bins} bins between {fmin} Hz and {fmax} Hz")
        self.waveform_generator.waveform_arguments["frequency_bin_edges"] = self.bin_freqs
        self.bin_widths = self.bin_freqs[1:] - self.bin_freqs[:-1]
        self.bin_centers = (self.bin_freqs[1:] + self.bin_freqs[:-1]) / 2
        for ifo in self.interferometers:
            self.per_detector_fiducial_waveform_points[ifo.name] = \
                self.per_detector_fiducial_waveforms[ifo.name][self.bin_inds]

    def compute_summary_data(self):
        """relative.py:319-363: per bin a0 = <h0|d>, a1 = <h0|d (f - fc)>, b0 = <h0|h0>, b1 = <h0|h0 (f - fc)>."""
        summary_data = dict()
        starts, centres = [], []
        for ifo in self.interferometers:
            mask = ifo.frequency_mask
            mf = ifo.frequency_array[mask]
            idx = np.searchsorted(mf, self.bin_freqs, side="left")
            if not np.array_equal(mf[idx], self.bin_freqs):
                raise ValueError("bin edges must lie on the interferometer's masked frequency grid")
            if idx[-1] < len(mf) - 1:
                idx[-1] += 1                      # the last bin takes the last edge point too
            first = int(np.argmax(mask))              # masked index -> grid index (the mask is one contiguous band)
            starts.append((idx + first).astype(np.int32))
            centres.append((mf[idx[:-1]] + mf[idx[1:]]) / 2)
        if any(not np.array_equal(s_, starts[0]) for s_ in starts[1:]):
            raise NotImplementedError("relative binning needs the same frequency band in every interferometer")
        # the four sums per bin on the device (bb_build_relbin_summary_data) from the data tiles already uploaded
        net = self.device_network
        n_det, n_freq = len(self.interferometers), len(self.interferometers[0].frequency_array)
        fid = np.zeros((n_det, n_freq, 2))
        for d, ifo in enumerate(self.interferometers):
            full = np.asarray(self.per_detector_fiducial_waveforms[ifo.name])
            fid[d, :, 0], fid[d, :, 1] = full.real, full.imag
        out = np.zeros((n_det, 4, self.number_of_bins, 2))
        bs = np.ascontiguousarray(starts[0], dtype=np.int32)
        ce = np.ascontiguousarray(centres[0], dtype=np.float64)
        _lib.check(net.lib.bb_build_relbin_summary_data(net.ptr, self.number_of_bins, bs.ctypes.data, ce.ctypes.data,
                                                        fid.ctypes.data, out.ctypes.data))
        for d, ifo in enumerate(self.interferometers):
            c = out[d, :, :, 0] + 1j * out[d, :, :, 1]
            summary_data[ifo.name] = (c[0], c[1], c[2], c[3])
        self.summary_data = summary_data

    def _pack_host_arrays(self):
        n_det, ne = len(self.interferometers), len(self.bin_freqs)
        fid = np.zeros((n_det, ne, 2))
        summ = np.zeros((n_det, 4, ne - 1, 2))
        grid = np.zeros((n_det, len(self.waveform_generator.frequency_array), 2))
        for d, ifo in enumerate(self.interferometers):
            pts = self.per_detector_fiducial_waveform_points[ifo.name]
            fid[d, :, 0], fid[d, :, 1] = pts.real, pts.imag
            for k, arr in enumerate(self.summary_data[ifo.name]):
                summ[d, k, :, 0], summ[d, k, :, 1] = arr.real, arr.imag
            full = self.per_detector_fiducial_waveforms[ifo.name]
            grid[d, :, 0], grid[d, :, 1] = full.real, full.imag
        return dict(bin_freqs=np.ascontiguousarray(self.bin_freqs, dtype=np.float64), fiducial=fid, summary=summ,
                    grid=grid, bin_inds=np.ascontiguousarray(self.bin_inds, dtype=np.int32))

    def _upload_relative_binning(self):
        net = self.device_network
        hst = self._rb_host
        tm = self.time_marginalization
        _lib.check(net.lib.bb_set_relative_binning(
            net.ptr, len(hst["bin_freqs"]), hst["bin_freqs"].ctypes.data, hst["fiducial"].ctypes.data,
            hst["summary"].ctypes.data, hst["grid"].ctypes.data if tm else None,
            hst["bin_inds"].ctypes.data if tm else None))

    def _configure(self):
        super()._configure()
        if self._rb_host is not None:
            self._upload_relative_binning()

    # ---- fiducial-point optimisation (relative.py:265-317) ----------------------------------------
    def find_maximum_likelihood_parameters(self, parameter_bounds, iterations=5, maximization_kwargs=None):
        from scipy.optimize import differential_evolution
        if maximization_kwargs is None:
            maximization_kwargs = dict()
        parameters = copy.deepcopy(self.fiducial_parameters)
        updated_list = self.get_parameter_list_from_dictionary(self.fiducial_parameters)
        old = self.log_likelihood_ratio(self.fiducial_parameters)
        updated = copy.deepcopy(self.fiducial_parameters)
        for it in range(iterations):
            logger.info(f"Optimizing fiducial parameters. Iteration : {it + 1}")
            output = differential_evolution(self.lnlike_scipy_maximize, bounds=parameter_bounds, args=(parameters,),
                                            x0=updated_list, **maximization_kwargs)
            updated_list = output["x"]
            updated = copy.deepcopy(self.fiducial_parameters)
            updated.update(self.get_parameter_dictionary_from_list(updated_list))
            self._setup_all(updated)
            new = self.log_likelihood_ratio(updated)
            if new - old < 0.1:
                break
            old = new
        return updated

    def lnlike_scipy_maximize(self, parameter_list, parameters):
        parameters.update(self.get_parameter_dictionary_from_list(parameter_list))
        return -self.log_likelihood_ratio(parameters)

    def get_parameter_dictionary_from_list(self, parameter_list):
        out = dict(zip(self.parameters_to_be_updated, parameter_list))
        for key in set(self.fiducial_parameters) - set(self.parameters_to_be_updated):
            out[key] = self.fiducial_parameters[key]
        return out

    def get_parameter_list_from_dictionary(self, parameter_dict):
        return [parameter_dict[k] for k in self.parameters_to_be_updated]

    def get_bounds_from_priors(self, priors):
        return [[priors[key].minimum, priors[key].maximum] for key in self.parameters_to_be_updated]

    # evaluation: the base class entry points (log_likelihood_ratio, log_likelihood_ratio_batch,
    # inner_products_batch, calculate_snrs) run K5 / K5t once the handle is in relative-binning mode.
