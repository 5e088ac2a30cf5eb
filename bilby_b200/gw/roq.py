"""ROQGravitationalWaveTransient on B200 (bilby/gw/likelihood/roq.py:27-1229).

Set-up (once per data set, host numpy like the reference): ROQ time grid (roq.py:747-767), linear weights by one
inverse FFT per basis element and detector (roq.py:849-916), quadratic weights (roq.py:976-1004), or weights
loaded from an .npz file written by ``save_weights``.  Evaluation on the device:
  K6 ``bb_roq_kernel``          waveform at the nodes, <h|h> from the quadratic weights, <d|h> at the five ROQ
                                times around the detector arrival time + cubic interpolation (roq.py:467-602)
  K7 hlinear + DMMA contraction (csrc/bb_gemm.cuh) + epilogue: time marginalisation, the dense all-times contraction
     W conj(h) (roq.py:604-651)
Bases: ndarray, .npy file, precomputed weights (dict / .npz), or the reference's hdf5 layout given as nested dicts
(``{"basis_linear": {"0": {"basis": [n_basis, n_freq], "frequency_nodes": ...}, "1": ...}, "prior_range_linear":
{"chirp_mass": [n_bases, 2], ...}}``; an .hdf5 file is read into that layout when h5py is importable).  With several
bases the one whose prior range contains a sample is chosen per sample (roq.py:368-439 _select_prior_ranges /
_update_basis): a batch is split into groups by (linear, quadratic) basis number, every group runs on its own device
set-up.  Multibanded bases (roq.py:920-1053) are not supported.
"""
import os

import numpy as np
from scipy import fft as _fft

from .. import _lib
from ..core.utils import logger, create_frequency_series
from .likelihood import GravitationalWaveTransient

RADIUS_OF_EARTH = 6378136.6      # bilby/core/utils/constants.py:6
SPEED_OF_LIGHT = 299792458.0


class BilbyROQParamsRangeError(Exception):
    pass


class ROQGravitationalWaveTransient(GravitationalWaveTransient):
    def __init__(self, interferometers, waveform_generator, priors, weights=None, linear_matrix=None,
                 quadratic_matrix=None, roq_params=None, roq_params_check=True, roq_scale_factor=1,
                 distance_marginalization=False, phase_marginalization=False, time_marginalization=False,
                 jitter_time=True, delta_tc=None, distance_marginalization_lookup_table=None,
                 reference_frame="sky", time_reference="geocenter", parameter_conversion=None, device=None):
        if getattr(waveform_generator.frequency_domain_source_model, "_bb_kind", None) != "roq":
            raise TypeError("ROQGravitationalWaveTransient needs one of the source models binary_black_hole_roq / "
                            "binary_neutron_star_roq")
        self._delta_tc = delta_tc
        self._roq_host = None
        super().__init__(interferometers=interferometers, waveform_generator=waveform_generator, priors=priors,
                         distance_marginalization=distance_marginalization,
                         phase_marginalization=phase_marginalization, time_marginalization=time_marginalization,
                         distance_marginalization_lookup_table=distance_marginalization_lookup_table,
                         jitter_time=jitter_time, reference_frame=reference_frame, time_reference=time_reference,
                         device=device)
        self.roq_params_check = roq_params_check
        self.roq_scale_factor = roq_scale_factor
        if isinstance(roq_params, str):
            self.roq_params_file = roq_params
            roq_params = np.genfromtxt(roq_params, names=True)
        elif not (roq_params is None or isinstance(roq_params, np.ndarray)):
            raise TypeError("roq_params should be array or str")
        self.roq_params = roq_params
        if isinstance(weights, dict):
            self.weights = weights
        elif isinstance(weights, str):
            self.weights = self.load_weights(weights)
        else:
            linear_matrix = self._parse_basis(linear_matrix, "linear")
            quadratic_matrix = self._parse_basis(quadratic_matrix, "quadratic")
            if self.roq_params is not None:
                for ifo in self.interferometers:
                    self.perform_roq_params_check(ifo)
            self.weights = dict()
            self._set_weights(linear_matrix, quadratic_matrix)
        self.number_of_bases_linear = len(self.weights[f"{self.interferometers[0].name}_linear"])
        self.number_of_bases_quadratic = len(self.weights[f"{self.interferometers[0].name}_quadratic"])
        self._cache = dict(parameters=None, basis_number_linear=None, basis_number_quadratic=None)
        self.parameter_conversion = parameter_conversion
        for basis_type in ("linear", "quadratic"):          # roq.py:198-204
            if getattr(self, f"number_of_bases_{basis_type}") > 1:
                self._verify_numbers_of_prior_ranges_and_frequency_nodes(basis_type)
            else:
                self._check_frequency_nodes_exist_for_single_basis(basis_type)
            self._verify_prior_ranges(basis_type)
        self._views = {}
        self._roq_pair = (0, 0)
        self._set_waveform_arguments_for_pair(0, 0)
        self._roq_host = self._pack_host_arrays(0, 0)
        if self.number_of_bases_linear == 1 and self.number_of_bases_quadratic == 1:
            self._upload_roq()

    # ---- set-up --------------------------------------------------------------------------------
    @property
    def roq_params(self):
        return self._roq_params

    @roq_params.setter
    def roq_params(self, roq_params):
        if roq_params is not None:
            if roq_params.shape != ():
                raise ValueError(f"roq_params must be a scalar structured array; received shape {roq_params.shape}")
            missing = {"flow", "fhigh", "seglen"}.difference(roq_params.dtype.names)
            if missing:
                raise ValueError("roq_params is missing required fields: " + ", ".join(sorted(missing)))
        self._roq_params = roq_params

    def _setup_time_marginalization(self):
        """roq.py:320-331 (overrides base.py:1027-1035)."""
        if self._delta_tc is None:
            self._delta_tc = self._get_time_resolution()
        tcmin = self.priors["geocent_time"].minimum
        tcmax = self.priors["geocent_time"].maximum
        number_of_time_samples = int(np.ceil((tcmax - tcmin) / self._delta_tc))
        self._delta_tc = (tcmax - tcmin) / number_of_time_samples
        self._times = tcmin + self._delta_tc / 2. + np.arange(number_of_time_samples) * self._delta_tc
        self._beam_pattern_reference_time = (tcmin + tcmax) / 2.
        self.time_prior_array = self.priors["geocent_time"].prob(self._times) * self._delta_tc

    @staticmethod
    def _parse_basis(basis, basis_type):
        """roq.py:333-366 -> the reference's hdf5 layout as nested dicts: ``basis_{type}/<i>/basis`` [n_basis, n_freq]
        (+ ``frequency_nodes``), optional ``prior_range_{type}/<parameter>`` [n_bases, 2]."""
        if isinstance(basis, str):
            fmt = basis.split(".")[-1]
            if fmt == "npy":
                return {f"basis_{basis_type}": {"0": {"basis": np.load(basis)}}}
            if fmt == "hdf5":
                try:
                    import h5py
                except ImportError:
                    raise NotImplementedError("hdf5 ROQ bases need h5py, which is not installed; pass the same layout "
                                              "as nested dicts of arrays instead") from None
                with h5py.File(basis, "r") as f:
                    def rd(g):
                        return {k: (rd(v) if hasattr(v, "keys") else v[()]) for k, v in g.items()}
                    out = rd(f)
                return out
            raise IOError(f"Format {fmt} not recognized.")
        if isinstance(basis, np.ndarray):
            return {f"basis_{basis_type}": {"0": {"basis": basis.T}}}
        if isinstance(basis, dict) and f"basis_{basis_type}" in basis:
            return basis
        raise TypeError("basis needs to be str, np.ndarray or a dict in the hdf5 layout")

    # ---- several bases (roq.py:218-439)
    def _verify_numbers_of_prior_ranges_and_frequency_nodes(self, basis_type):
        """roq.py:218-250."""
        number_of_bases = getattr(self, f"number_of_bases_{basis_type}")
        key = f"prior_range_{basis_type}"
        if key not in self.weights:
            raise AttributeError(f'For the use of multiple {basis_type} ROQ bases, weights should contain "{key}".')
        for param_name, ranges in self.weights[key].items():
            if len(ranges) != number_of_bases:
                raise ValueError(f'The number of prior ranges for "{param_name}" does not match the number of '
                                 f"{basis_type} bases")
        key = f"frequency_nodes_{basis_type}"
        if key not in self.weights:
            raise AttributeError(f'For the use of multiple {basis_type} ROQ bases, weights should contain "{key}".')
        if len(self.weights[key]) != number_of_bases:
            raise ValueError(f"The number of arrays of frequency nodes does not match the number of {basis_type} bases")

    def _verify_prior_ranges(self, basis_type):
        """roq.py:252-277: the union of the bases' ranges must cover the prior."""
        key = f"prior_range_{basis_type}"
        if key not in self.weights:
            return
        for param_name, ranges in self.weights[key].items():
            ranges = np.asarray(ranges)
            if self.priors[param_name].minimum < np.min(ranges[:, 0]):
                raise BilbyROQParamsRangeError(f"Prior minimum of {param_name} {self.priors[param_name].minimum} less "
                                               f"than ROQ basis bound {np.min(ranges[:, 0])}")
            if self.priors[param_name].maximum > np.max(ranges[:, 1]):
                raise BilbyROQParamsRangeError(f"Prior maximum of {param_name} {self.priors[param_name].maximum} "
                                               f"greater than ROQ basis bound {np.max(ranges[:, 1])}")

    def _select_prior_ranges(self, prior_ranges):
        """roq.py:368-400: the bases whose ranges intersect the priors."""
        names = list(prior_ranges.keys())
        n = len(prior_ranges[names[0]])
        keep = np.ones(n, dtype=bool)
        for name in names:
            if self.priors is None or name not in self.priors:
                continue
            r = np.asarray(prior_ranges[name])
            keep &= (r[:, 1] >= self.priors[name].minimum) & (r[:, 0] <= self.priors[name].maximum)
        idx = np.arange(n)[keep]
        return idx, {name: np.asarray(prior_ranges[name])[idx] for name in names}

    def _basis_numbers(self, parameters, n):
        """Vectorised roq.py:402-429 (_update_basis): per sample the FIRST basis whose range contains it, for the
        linear and the quadratic family.  parameters: dict of scalars / length-n arrays, converted first
        (``parameter_conversion``) like the reference."""
        pars = dict(parameters)
        if self.parameter_conversion is not None:
            conv = self.parameter_conversion(pars)
            pars = conv[0] if isinstance(conv, tuple) else conv
        out = []
        for basis_type in ("linear", "quadratic"):
            nb = getattr(self, f"number_of_bases_{basis_type}")
            if nb == 1:
                out.append(np.zeros(n, dtype=int))
                continue
            inside = np.ones((n, nb), dtype=bool)
            for name, ranges in self.weights[f"prior_range_{basis_type}"].items():
                if name not in pars:
                    continue
                v = np.broadcast_to(np.asarray(pars[name], dtype=float), (n,))[:, None]
                r = np.asarray(ranges)
                inside &= (r[None, :, 0] <= v) & (r[None, :, 1] >= v)
            if not np.all(inside.any(axis=1)):
                raise IndexError(f"a sample lies outside every {basis_type} ROQ basis' prior range")
            out.append(np.argmax(inside, axis=1))
        return out[0], out[1]

    def _update_basis(self, parameters):
        """roq.py:402-439 for one parameter dict."""
        bl, bq = self._basis_numbers(parameters, 1)
        self._cache.update(parameters=dict(parameters), basis_number_linear=int(bl[0]), basis_number_quadratic=int(bq[0]))
        self._set_waveform_arguments_for_pair(int(bl[0]), int(bq[0]))

    def _set_waveform_arguments_for_pair(self, bl, bq):
        nodes_l = np.asarray(self.weights["frequency_nodes_linear"][bl], dtype=float)
        nodes_q = np.asarray(self.weights["frequency_nodes_quadratic"][bq], dtype=float)
        unique, inverse = np.unique(np.hstack((nodes_l, nodes_q)), return_inverse=True)
        wa = self.waveform_generator.waveform_arguments
        wa["frequency_nodes"] = unique                      # roq.py:206-213, 433-438
        wa["linear_indices"] = inverse[:len(nodes_l)]
        wa["quadratic_indices"] = inverse[len(nodes_l):]

    def _view(self, bl, bq):
        """This likelihood restricted to ONE (linear, quadratic) basis pair, with its own device set-up."""
        if (bl, bq) not in self._views:
            import copy
            v = copy.copy(self)
            v.waveform_generator = copy.copy(self.waveform_generator)
            v.waveform_generator.waveform_arguments = dict(self.waveform_generator.waveform_arguments)
            v._net, v._net_versions = None, None
            v.number_of_bases_linear = v.number_of_bases_quadratic = 1
            v._roq_pair = (bl, bq)
            v._views = {}
            v._set_waveform_arguments_for_pair(bl, bq)
            v._roq_host = self._pack_host_arrays(bl, bq)
            v._upload_roq()
            self._views[(bl, bq)] = v
        return self._views[(bl, bq)]

    def log_likelihood_ratio_batch(self, parameters):
        if self.number_of_bases_linear == 1 and self.number_of_bases_quadratic == 1:
            return super().log_likelihood_ratio_batch(parameters)
        if not isinstance(parameters, dict):
            raise ValueError("with several ROQ bases the parameters must be a dict (the basis depends on their values)")
        n = max(np.size(v) for v in parameters.values())
        host = {k: (np.asarray(v.cpu()) if hasattr(v, "cpu") else v) for k, v in parameters.items()}
        bl, bq = self._basis_numbers(host, n)
        out = np.empty(n)
        for pair in sorted(set(zip(bl.tolist(), bq.tolist()))):
            idx = np.where((bl == pair[0]) & (bq == pair[1]))[0]
            sub = {k: (np.asarray(v)[idx] if np.ndim(v) else v) for k, v in host.items()}
            out[idx] = self._view(*pair).log_likelihood_ratio_batch(sub)
        return out

    def log_likelihood_ratio(self, parameters):
        """roq.py:463-465."""
        if self.number_of_bases_linear == 1 and self.number_of_bases_quadratic == 1:
            return super().log_likelihood_ratio(parameters)
        self._update_basis(parameters)
        return self._view(self._cache["basis_number_linear"], self._cache["basis_number_quadratic"]).log_likelihood_ratio(
            parameters)

    def _check_frequency_nodes_exist_for_single_basis(self, basis_type):
        """roq.py:278-295."""
        key = f"frequency_nodes_{basis_type}"
        wa = self.waveform_generator.waveform_arguments
        if not (key in self.weights or key in wa):
            raise AttributeError(f"{key} should be contained in weights or waveform arguments.")
        elif key not in wa:
            wa[key] = self.weights[key][0]
        elif key not in self.weights:
            self.weights[key] = [wa[key]]

    def perform_roq_params_check(self, ifo=None):
        """roq.py:653-734 (frequency / duration checks; the CBCPriorDict mass checks need astropy priors)."""
        if self.roq_params_check is False:
            logger.warning("No ROQ params checking performed")
            return
        p = self.roq_params
        if float(ifo.maximum_frequency) > p["fhigh"] * self.roq_scale_factor:
            raise BilbyROQParamsRangeError("Requested maximum frequency {} larger than ROQ basis fhigh {}".format(
                ifo.maximum_frequency, p["fhigh"] * self.roq_scale_factor))
        if float(ifo.minimum_frequency) < p["flow"] * self.roq_scale_factor:
            raise BilbyROQParamsRangeError("Requested minimum frequency {} lower than ROQ basis flow {}".format(
                ifo.minimum_frequency, p["flow"] * self.roq_scale_factor))
        if float(ifo.strain_data.duration) != p["seglen"] / self.roq_scale_factor:
            raise BilbyROQParamsRangeError("Requested duration differs from ROQ basis seglen")

    def _get_time_resolution(self):
        """roq.py:1165-1229: time step from the bandwidth the injected SNR can resolve, rounded so that
        duration / delta_t is a power of two.  (PSD and frequencies of the LAST interferometer, like the
        reference.)"""
        from scipy.integrate import simpson
        inj_snr_sq = 0
        for ifo in self.interferometers:
            inj_snr_sq += max(10, ifo.meta_data.get("optimal_SNR", 30)) ** 2
        psd = ifo.power_spectral_density_array[ifo.frequency_mask]
        freq = ifo.frequency_array[ifo.frequency_mask]
        integral1 = simpson(y=np.power(freq, -7. / 3) / psd, x=freq)
        f_3_bar = simpson(y=np.power(freq, 2. / 3.) / (psd * integral1), x=freq)
        scaling = (np.pi ** 2 * inj_snr_sq / 6) ** (1 / 3)
        delta_t = (scaling * f_3_bar ** (1 / 3)) ** -1 / 5
        duration = self.interferometers.duration
        n = max(duration / delta_t, self.interferometers.frequency_array[-1] * duration + 1)
        n = int(2 ** np.ceil(np.log2(n)))
        return duration / n

    def _roq_time_prior(self):
        """roq.py:747-765: the ROQ time grid spans the prior on ``{time_reference}_time`` (the detector the sampler
        times the signal at), not necessarily the geocentre."""
        key = f"{self.time_reference}_time"
        prior = None if self.priors is None else self.priors.get(key)
        if prior is None or not hasattr(prior, "minimum"):
            raise KeyError(f"ROQGravitationalWaveTransient needs a prior on {key!r} to place the ROQ time samples "
                           "(roq.py:747-765)")
        return prior

    def _set_weights(self, linear_matrix, quadratic_matrix):
        """roq.py:736-767 (time grid), 768-792 (basis selection by prior range), 802-837 (basis / data frequency
        overlap), 849-916 (linear), 976-1004 (quadratic).  linear_matrix / quadratic_matrix: the hdf5 layout as nested
        dicts (_parse_basis)."""
        time_space = self._get_time_resolution()
        duration = self.interferometers.duration
        start_time = self.interferometers.start_time
        prior = self._roq_time_prior()
        number_of_time_samples = int(duration / time_space)
        crossing = 2 * RADIUS_OF_EARTH / SPEED_OF_LIGHT + 5 * time_space
        start_idx = max(0, int(np.floor((prior.minimum - crossing - start_time) / time_space)))
        end_idx = min(number_of_time_samples - 1, int(np.ceil((prior.maximum + crossing - start_time) / time_space)))
        self.weights["time_samples"] = np.arange(start_idx, end_idx + 1) * float(time_space)
        logger.info("Using {} ROQ time samples".format(len(self.weights["time_samples"])))
        # bases inside the prior range, their ranges and nodes (roq.py:768-792)
        selected = {}
        for basis_type, matrix in (("linear", linear_matrix), ("quadratic", quadratic_matrix)):
            key = f"prior_range_{basis_type}"
            if key in matrix:
                ranges = {name: np.asarray(matrix[key][name], dtype=float) for name in matrix[key]}
                idxs, sel = self._select_prior_ranges(ranges)
                if len(idxs) == 0:
                    raise BilbyROQParamsRangeError(f"There are no {basis_type} ROQ bases within the prior range.")
                self.weights[key] = sel
                selected[basis_type] = [int(i) for i in idxs]
            else:
                selected[basis_type] = [0]
            first = matrix[f"basis_{basis_type}"][str(selected[basis_type][0])]
            if "frequency_nodes" in first:
                self.weights[f"frequency_nodes_{basis_type}"] = [
                    np.asarray(matrix[f"basis_{basis_type}"][str(i)]["frequency_nodes"]) * self.roq_scale_factor
                    for i in selected[basis_type]]
        for ifo in self.interferometers:
            self.weights[ifo.name + "_linear"], self.weights[ifo.name + "_quadratic"] = [], []
        # roq.py:792-799, 839-847: a multibanded basis lives on its own banded frequency grid
        if bool(np.asarray(linear_matrix.get("multiband_linear", False))[()]):
            self._set_weights_linear_multiband(linear_matrix, selected["linear"])
        else:
            for i in selected["linear"]:
                w = self._weights_of_one_basis(np.asarray(linear_matrix["basis_linear"][str(i)]["basis"]), None)
                for ifo in self.interferometers:
                    self.weights[ifo.name + "_linear"].append(w[ifo.name + "_linear"])
        if bool(np.asarray(quadratic_matrix.get("multiband_quadratic", False))[()]):
            self._set_weights_quadratic_multiband(quadratic_matrix, selected["quadratic"])
        else:
            for i in selected["quadratic"]:
                w = self._weights_of_one_basis(None, np.asarray(quadratic_matrix["basis_quadratic"][str(i)]["basis"]))
                for ifo in self.interferometers:
                    self.weights[ifo.name + "_quadratic"].append(w[ifo.name + "_quadratic"])

    @staticmethod
    def _bands(matrix, basis_type, scale):
        """Band durations (scaled), [start, end] frequency bins per band, the basis dimension and the highest basis
        frequency of a multibanded basis (roq.py:934-938 / 1022-1026)."""
        tbs = np.atleast_1d(np.asarray(matrix[f"durations_s_{basis_type}"], dtype=float)) / scale
        bins = np.asarray(matrix[f"start_end_frequency_bins_{basis_type}"], dtype=int).reshape(-1, 2)
        dim = int(np.sum(bins[:, 1] - bins[:, 0] + 1))
        return tbs, bins, dim, float(np.max(bins[:, 1] / tbs))

    def _set_weights_linear_multiband(self, linear_matrix, basis_idxs):
        """roq.py:920-974: time-dependent linear weights from a multibanded basis.  Per band b (duration T_b, bins
        [k0, k1] of spacing 1 / T_b) the over-whitened data d / S goes to the time domain, its last 2 f_high T_b samples
        come back as D_b[k], and   w[t, i] = sum_b sum_k conj(B_i[b, k]) (4 / T_b) D_b[k] exp(2 pi i f_k (t - T + T_b)).
        The transforms are host numpy (once per data set); the contraction over the banded points is the library's
        FP64 tensor-core kernel when a CUDA device is there."""
        tbs, bins, dim, fhigh = self._bands(linear_matrix, "linear", self.roq_scale_factor)
        ts = np.asarray(self.weights["time_samples"], dtype=float)
        shifted = {}
        for ifo in self.interferometers:
            mask = ifo.frequency_mask
            spec = np.zeros(int(fhigh * ifo.duration) + 1, dtype=complex)
            spec[np.arange(len(ifo.frequency_domain_strain))[mask]] = (
                ifo.frequency_domain_strain[mask] / ifo.power_spectral_density_array[mask])
            td = np.fft.irfft(spec)
            rows = np.zeros((dim, len(ts)), dtype=complex)
            at = 0
            for (k0, k1), tb in zip(bins, tbs):
                fs = np.arange(k0, k1 + 1) / tb
                db = np.fft.rfft(td[-int(2. * fhigh * tb):])[k0:k1 + 1]
                rows[at:at + k1 - k0 + 1] = 4. / tb * db[:, None] * np.exp(
                    2. * np.pi * 1j * fs[:, None] * (ts[None, :] - ifo.duration + tb))
                at += k1 - k0 + 1
            shifted[ifo.name] = rows
        for i in basis_idxs:
            logger.info(f"Building linear ROQ weights for the {i}-th basis.")
            basis = np.asarray(linear_matrix["basis_linear"][str(i)]["basis"])
            for ifo in self.interferometers:
                self.weights[ifo.name + "_linear"].append(self._contract(shifted[ifo.name].T, basis.conj()))

    def _set_weights_quadratic_multiband(self, quadratic_matrix, basis_idxs):
        """roq.py:1006-1053: quadratic weights from a multibanded basis,  w[i] = sum_b sum_k B_i[b, k] (4 / T_b)
        Re rfft(inverse-PSD time series folded to 2 f_high T_b samples)[k]."""
        tbs, bins, dim, fhigh = self._bands(quadratic_matrix, "quadratic", self.roq_scale_factor)
        folded = {}
        for ifo in self.interferometers:
            mask = ifo.frequency_mask
            inv = np.zeros(int(fhigh * ifo.duration) + 1)
            inv[np.arange(len(ifo.power_spectral_density_array))[mask]] = 1. / ifo.power_spectral_density_array[mask]
            td = np.fft.irfft(inv)
            vec = np.zeros(dim)
            at = 0
            for (k0, k1), tb in zip(bins, tbs):
                half = int(fhigh * tb)
                vec[at:at + k1 - k0 + 1] = 4. / tb * np.fft.rfft(np.concatenate([td[:half], td[-half:]]))[k0:k1 + 1].real
                at += k1 - k0 + 1
            folded[ifo.name] = vec
        for i in basis_idxs:
            logger.info(f"Building quadratic ROQ weights for the {i}-th basis.")
            basis = np.asarray(quadratic_matrix["basis_quadratic"][str(i)]["basis"]).real
            for ifo in self.interferometers:
                self.weights[ifo.name + "_quadratic"].append(basis @ folded[ifo.name])

    def _contract(self, a, b):
        """a [m, k] @ b [n, k].T (complex): bb_contract_device (DMMA) with a CUDA device, numpy without.  Runs on a bare
        handle of its own: the likelihood's device state is configured only after the weights exist."""
        try:
            import torch
            if not torch.cuda.is_available():
                return a @ b.T
        except ImportError:      # pragma: no cover
            return a @ b.T
        import ctypes
        from .. import _lib
        handle = _lib.Handle(self._device_index)
        device = torch.device("cuda", handle.device)
        m, k = a.shape
        n = b.shape[0]

        def dev(x):
            # complex128 arrays are (re, im) pairs in memory
            x = np.ascontiguousarray(x, dtype=np.complex128)
            return torch.from_numpy(x.view(np.float64).reshape(x.shape + (2,))).to(device)
        ad, bd = dev(a), dev(b)
        cd = torch.empty((m, n, 2), dtype=torch.float64, device=device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _lib.check(handle.lib.bb_contract_device(handle.ptr, 1, m, n, k, 1, 0, 0, 1, 0, 0, 0, 1.0, ad.data_ptr(), k,
                                                 bd.data_ptr(), k, 0, cd.data_ptr(), n, stream))
        torch.cuda.synchronize(device)
        c = cd.cpu().numpy()
        return c[..., 0] + 1j * c[..., 1]

    def _weights_of_one_basis(self, linear_basis, quadratic_basis):
        """Linear (roq.py:849-918) or quadratic (roq.py:976-1004) weights of every detector for ONE basis
        [n_basis, n_basis_freq]."""
        duration = self.interferometers.duration
        ts = self.weights["time_samples"]
        space = ts[1] - ts[0]
        n_time = int(duration / space)
        lo, hi = int(ts[0] / space), int(ts[-1] / space)
        out = {}
        n_basis_freq = (linear_basis if linear_basis is not None else quadratic_basis).shape[1]
        # per detector: overlap of the basis frequencies with the data (roq.py:802-837)
        setups = []
        for ifo in self.interferometers:
            mask = ifo.frequency_mask
            if self.roq_params is not None:
                fhigh = self.roq_params["fhigh"] * self.roq_scale_factor
                seglen = self.roq_params["seglen"] / self.roq_scale_factor
                roq_f = create_frequency_series(sampling_frequency=fhigh * 2, duration=seglen)
                roq_f = roq_f[roq_f >= self.roq_params["flow"] * self.roq_scale_factor]
                _, ifo_idxs, roq_idxs = np.intersect1d(ifo.frequency_array[mask], roq_f, return_indices=True)
            else:
                roq_idxs = np.arange(n_basis_freq, dtype=int)
                ifo_idxs = np.arange(int(mask.sum()))
                if len(ifo_idxs) != len(roq_idxs):
                    raise ValueError("Mismatch between ROQ basis and frequency array for {}".format(ifo.name))
            nonzero = ifo_idxs + int(ifo.minimum_frequency * duration)
            d_over_s = ifo.frequency_domain_strain[mask][ifo_idxs] / ifo.power_spectral_density_array[mask][ifo_idxs]
            setups.append((ifo, roq_idxs, ifo_idxs, nonzero, d_over_s))
            if quadratic_basis is not None:
                inv_psd = 1 / ifo.power_spectral_density_array[mask][ifo_idxs]
                out[ifo.name + "_quadratic"] = self._quadratic_weights_device(
                    inv_psd, quadratic_basis.real[:, roq_idxs], duration)
        if linear_basis is None:
            return out
        # linear weights (roq.py:849-918): on the device, all detectors in one call when they share the frequency set
        same = all(np.array_equal(su[1], setups[0][1]) and np.array_equal(su[3], setups[0][3]) for su in setups)
        groups = [setups] if same else [[su] for su in setups]
        for group in groups:
            roq_idxs, nonzero = group[0][1], group[0][3]
            identity = len(roq_idxs) == linear_basis.shape[1] and np.array_equal(roq_idxs, np.arange(linear_basis.shape[1]))
            basis = linear_basis if identity else linear_basis[:, roq_idxs]
            lws = self._linear_weights_device(np.array([su[4] for su in group]), basis, nonzero, n_time, lo, hi, duration)
            for i, (ifo, _, _, _, d_over_s) in enumerate(group):
                if lws is not None:
                    lw = lws[i]
                else:
                    # no CUDA device in this process (set-up only, e.g. building weight files on a login node): one
                    # inverse FFT per basis element with scipy's pocketfft, in slabs of 32 elements to bound memory
                    lw = np.empty((hi - lo + 1, linear_basis.shape[0]), dtype=complex)
                    for b0 in range(0, linear_basis.shape[0], 32):
                        sl = slice(b0, min(b0 + 32, linear_basis.shape[0]))
                        spec = np.zeros((sl.stop - sl.start, n_time), dtype=complex)
                        spec[:, nonzero] = d_over_s[None, :] * basis[sl].conj()
                        lw[:, sl] = _fft.ifft(spec, axis=1, workers=os.cpu_count() or 1)[:, lo:hi + 1].T
                    lw *= 4. * n_time / duration
                out[ifo.name + "_linear"] = lw
        return out

    def _quadratic_weights_device(self, inv_psd, basis_real, duration):
        """roq.py:976-1004 on the device (bb_build_roq_quadratic_weights); numpy when the process has no CUDA device
        (weight files built on a login node)."""
        try:
            import torch
            have = torch.cuda.is_available()
            dev = (torch.cuda.current_device() if self._device_index is None else int(self._device_index)) if have else 0
        except ImportError:      # pragma: no cover
            have = False
        if not have:
            return 4. / duration * basis_real @ inv_psd
        lib = _lib.load()
        p = np.ascontiguousarray(inv_psd, dtype=np.float64)
        b = np.ascontiguousarray(basis_real, dtype=np.float64)
        out = np.empty(b.shape[0])
        _lib.check(lib.bb_build_roq_quadratic_weights(dev, 1, len(p), p.ctypes.data, b.shape[0], b.ctypes.data,
                                                      float(duration), out.ctypes.data))
        return out

    def _linear_weights_device(self, d_over_s, basis, bin_index, n_time, lo, hi, duration):
        """roq.py:849-918 on the device (bb_build_roq_linear_weights: the wanted time samples as one DMMA contraction (csrc/bb_gemm.cuh) against an
        exact phase matrix).  Returns None when the process has no CUDA device."""
        try:
            import torch
            if not torch.cuda.is_available():
                return None
            dev = torch.cuda.current_device() if self._device_index is None else int(self._device_index)
        except ImportError:      # pragma: no cover
            return None
        from .. import _lib
        lib = _lib.load()
        n_win = hi - lo + 1
        # complex128 arrays are (re, im) pairs in memory: hand them over without copies
        dos = np.ascontiguousarray(d_over_s, dtype=np.complex128)
        bas = np.ascontiguousarray(basis, dtype=np.complex128)
        idx = np.ascontiguousarray(bin_index, dtype=np.int32)
        out = np.empty((dos.shape[0], n_win, bas.shape[0]), dtype=np.complex128)
        _lib.check(lib.bb_build_roq_linear_weights(dev, dos.shape[0], len(idx), dos.ctypes.data, bas.shape[0],
                                                   bas.ctypes.data, idx.ctypes.data, int(n_time), int(lo), int(n_win),
                                                   float(duration), out.ctypes.data))
        return out

    def save_weights(self, filename, format="npz"):
        """roq.py:1055-1100.  npz: the reference's single-basis keys; with several bases (which the reference only
        writes as hdf5) the same keys carry a basis index: ``{IFO}_linear/<i>``, ``frequency_nodes_linear/<i>``,
        ``prior_range_linear/<parameter>``.  hdf5 (the reference's layout) when h5py is importable."""
        if format not in ("npz", "hdf5"):
            raise IOError(f"Format {format} not recognized.")
        if format not in filename:
            filename += "." + format
        multi = self.number_of_bases_linear > 1 or self.number_of_bases_quadratic > 1
        flat = dict(time_samples=self.weights["time_samples"])
        for basis_type in ("linear", "quadratic"):
            keys = [f"{ifo.name}_{basis_type}" for ifo in self.interferometers] + [f"frequency_nodes_{basis_type}"]
            for key in keys:
                if key not in self.weights:
                    continue
                if multi or format == "hdf5":
                    for i, w in enumerate(self.weights[key]):
                        flat[f"{key}/{i}"] = np.asarray(w)
                else:
                    flat[key] = np.asarray(self.weights[key][0])
            for name, r in self.weights.get(f"prior_range_{basis_type}", {}).items():
                flat[f"prior_range_{basis_type}/{name}"] = np.asarray(r)
        if format == "npz":
            np.savez(filename, **flat)
            return
        try:
            import h5py
        except ImportError:
            raise IOError("hdf5 weight files need h5py, which is not installed; use format='npz'") from None
        with h5py.File(filename, "w") as f:
            for key, val in flat.items():
                f.create_dataset(key, data=val)

    def load_weights(self, filename, format=None):
        """roq.py:1102-1163 (npz: single basis as the reference writes it, or the indexed keys of save_weights; hdf5
        through h5py when importable).  Bases outside the prior range are dropped (_select_prior_ranges)."""
        if format is None:
            format = filename.split(".")[-1]
        if format not in ("npz", "hdf5"):
            raise IOError(f"Format {format} not recognized.")
        if format == "npz":
            flat = dict(np.load(filename))
        else:
            try:
                import h5py
            except ImportError:
                raise IOError("hdf5 weight files need h5py, which is not installed") from None
            flat = {}
            with h5py.File(filename, "r") as f:
                f.visititems(lambda name, obj: flat.__setitem__(name, obj[()]) if hasattr(obj, "shape") else None)
        weights = dict(time_samples=flat["time_samples"])
        for basis_type in ("linear", "quadratic"):
            pr = {k.split("/", 1)[1]: v for k, v in flat.items() if k.startswith(f"prior_range_{basis_type}/")}
            idxs = None
            if pr:
                idxs, sel = self._select_prior_ranges(pr)
                weights[f"prior_range_{basis_type}"] = sel
            keys = [f"{ifo.name}_{basis_type}" for ifo in self.interferometers] + [f"frequency_nodes_{basis_type}"]
            for key in keys:
                if key in flat:
                    weights[key] = [flat[key]]
                    continue
                n = len([k for k in flat if k.startswith(key + "/")])
                if n == 0:
                    continue
                take = range(n) if idxs is None else idxs
                weights[key] = [flat[f"{key}/{int(i)}"] for i in take]
        return weights

    def _pack_host_arrays(self, bl, bq):
        """Host arrays of bb_set_roq for the (linear basis bl, quadratic basis bq) pair."""
        nodes_l = np.asarray(self.weights["frequency_nodes_linear"][bl], dtype=float)
        nodes_q = np.asarray(self.weights["frequency_nodes_quadratic"][bq], dtype=float)
        ts = np.asarray(self.weights["time_samples"], dtype=float)
        names = [ifo.name for ifo in self.interferometers]
        wl = np.stack([np.asarray(self.weights[n + "_linear"][bl], dtype=complex) for n in names])      # [d, t, i]
        wq = np.stack([np.asarray(self.weights[n + "_quadratic"][bq], dtype=float) for n in names])
        if wl.shape[1] != len(ts) or wl.shape[2] != len(nodes_l) or wq.shape[1] != len(nodes_q):
            raise ValueError("ROQ weights do not match the time samples / frequency nodes")
        # time_samples = arange(start_idx, end_idx + 1) * time_space (roq.py:766): recover both exactly
        if len(ts) > 1 and ts[0] != 0:
            idx0 = int(round(ts[0] / (ts[1] - ts[0])))
            step = ts[0] / idx0 if idx0 != 0 else ts[1] - ts[0]
        else:
            idx0, step = 0, ts[1] - ts[0]
        if not np.allclose((idx0 + np.arange(len(ts))) * step, ts, rtol=0, atol=1e-12):
            raise ValueError("ROQ time samples must be an integer-index multiple of the time step")
        wlv = np.empty(wl.shape + (2,))
        wlv[..., 0], wlv[..., 1] = wl.real, wl.imag
        return dict(nodes_l=np.ascontiguousarray(nodes_l), nodes_q=np.ascontiguousarray(nodes_q),
                    wl=np.ascontiguousarray(wlv), wq=np.ascontiguousarray(wq), n_time=len(ts), idx0=idx0,
                    step=float(step))

    def _upload_roq(self):
        net = self.device_network
        hst = self._roq_host
        n_marg, t0, dtc, tref = 0, 0.0, 0.0, 0.0
        if self.time_marginalization:
            n_marg = len(self._times)
            t0 = float(self.priors["geocent_time"].minimum)
            dtc = float(self._delta_tc)
            tref = float(self._beam_pattern_reference_time)
        _lib.check(net.lib.bb_set_roq(
            net.ptr, len(hst["nodes_l"]), hst["nodes_l"].ctypes.data, len(hst["nodes_q"]), hst["nodes_q"].ctypes.data,
            hst["n_time"], hst["idx0"], hst["step"], hst["wl"].ctypes.data, hst["wq"].ctypes.data, n_marg, t0, dtc,
            tref))

    def _configure(self):
        super()._configure()
        if self._roq_host is not None:
            self._upload_roq()

    @property
    def basis_number_linear(self):
        return self._cache["basis_number_linear"] if self.number_of_bases_linear > 1 else 0

    @property
    def basis_number_quadratic(self):
        return self._cache["basis_number_quadratic"] if self.number_of_bases_quadratic > 1 else 0
