"""TEST INFRASTRUCTURE ONLY (oracle) - CPU restatement of the reference's two reduced-order likelihoods
(SURVEY.md section 8 rows a19, a20):

  * RelativeBinningGravitationalWaveTransient   bilby/gw/likelihood/relative.py:105-430
  * ROQGravitationalWaveTransient               bilby/gw/likelihood/roq.py:112-651, 736-1004

and of the source models that feed them (bilby/gw/source.py:693-898, 1068-1140), backed by the restated
IMRPhenomD / TaylorF2 evaluated on a frequency SEQUENCE (oracle/phenomd.py, oracle/taylorf2.py).

PARITY STATUS: the likelihood arithmetic is PINNED - tests/golden/{relbin,roq}_*.npz were produced by the
UNMODIFIED reference classes (oracle/tools/make_golden_reduced.py) and tests/test_oracle_vs_golden.py checks
this file against them.  The waveform arithmetic itself (lalsimulation) stays unpinned, as in oracle/phenomd.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

from . import cbc_likelihood as ocl
from . import phenomd as _pd
from . import taylorf2 as _tf2

RADIUS_OF_EARTH = 6378136.6   # bilby/core/utils/constants.py
SPEED_OF_LIGHT = 299792458.0


# --------------------------------------------------------------------------------------
# source models on frequency sequences (source.py:1068-1140 _base_waveform_frequency_sequence)
# --------------------------------------------------------------------------------------
def _sequence_polarizations(frequencies, mass_1, mass_2, luminosity_distance, a_1, tilt_1, a_2, tilt_2,
                            theta_jn, phase, lambda_1, lambda_2, waveform_approximant, reference_frequency,
                            catch_waveform_errors=False, **unused):
    for a, tilt in ((a_1, tilt_1), (a_2, tilt_2)):
        if not (a == 0 or tilt in (0, np.pi)):
            raise ValueError("aligned-spin models only")
    s1z, s2z = a_1 * np.cos(tilt_1), a_2 * np.cos(tilt_2)
    dist = luminosity_distance * 1e6 * _pd.PARSEC
    try:
        if waveform_approximant == "IMRPhenomD":
            hp, hc = _pd.choose_fd_waveform_phenomd(frequencies, mass_1, mass_2, s1z, s2z, dist, theta_jn, phase,
                                                    0.0, 0.0, reference_frequency, sequence=True)
        elif waveform_approximant == "TaylorF2":
            hp, hc = _tf2.choose_fd_waveform_taylorf2(frequencies, mass_1, mass_2, s1z, s2z, lambda_1, lambda_2,
                                                      dist, theta_jn, phase, 0.0, 0.0, reference_frequency,
                                                      sequence=True)
        else:
            raise ValueError("oracle restates IMRPhenomD and TaylorF2 only")
    except _pd.WaveformDomainError:
        if catch_waveform_errors:
            return None
        raise
    return dict(plus=hp, cross=hc)


_RB_DROP = ("pn_spin_order", "pn_tidal_order", "pn_phase_order", "pn_amplitude_order")


def lal_binary_black_hole_relative_binning(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1,
                                           phi_12, a_2, tilt_2, phi_jl, theta_jn, phase, **kwargs):
    """source.py:724-761: fiducial=1 -> full grid (lal_binary_black_hole), else the bin-edge sequence."""
    kwargs = dict(kwargs)
    fiducial = kwargs.pop("fiducial", 0)
    wa = dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0,
              maximum_frequency=frequency_array[-1], catch_waveform_errors=False)
    wa.update(kwargs)
    if fiducial == 1:
        wa.pop("frequency_bin_edges", None)
        return ocl.lal_binary_black_hole(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12,
                                         a_2, tilt_2, phi_jl, theta_jn, phase,
                                         **{k: v for k, v in wa.items() if k not in _RB_DROP})
    wa.pop("minimum_frequency", None)
    wa.pop("maximum_frequency", None)
    freqs = wa.pop("frequency_bin_edges")
    return _sequence_polarizations(freqs, mass_1, mass_2, luminosity_distance, a_1, tilt_1, a_2, tilt_2, theta_jn,
                                   phase, 0.0, 0.0, **wa)


def lal_binary_neutron_star_relative_binning(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1,
                                             phi_12, a_2, tilt_2, phi_jl, lambda_1, lambda_2, theta_jn, phase,
                                             **kwargs):
    """source.py:764-799."""
    kwargs = dict(kwargs)
    fiducial = kwargs.pop("fiducial", 0)
    wa = dict(waveform_approximant="TaylorF2", reference_frequency=50.0, minimum_frequency=20.0,
              maximum_frequency=frequency_array[-1], catch_waveform_errors=False)
    wa.update(kwargs)
    if fiducial == 1:
        wa.pop("frequency_bin_edges", None)
        return ocl.lal_binary_neutron_star(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1,
                                           phi_12, a_2, tilt_2, phi_jl, theta_jn, phase, lambda_1, lambda_2,
                                           **{k: v for k, v in wa.items() if k not in _RB_DROP})
    wa.pop("minimum_frequency", None)
    wa.pop("maximum_frequency", None)
    freqs = wa.pop("frequency_bin_edges")
    return _sequence_polarizations(freqs, mass_1, mass_2, luminosity_distance, a_1, tilt_1, a_2, tilt_2, theta_jn,
                                   phase, lambda_1, lambda_2, **wa)


def _base_roq_waveform(mass_1, mass_2, luminosity_distance, a_1, tilt_1, a_2, tilt_2, theta_jn, phase, lambda_1,
                       lambda_2, **wa):
    """source.py:802-898: waveform at the unique nodes, gathered into linear / quadratic node order."""
    wa = dict(wa)
    if "frequency_nodes" not in wa:
        size_linear = len(wa["frequency_nodes_linear"])
        combined = np.hstack((wa.pop("frequency_nodes_linear"), wa.pop("frequency_nodes_quadratic")))
        unique, inverse = np.unique(combined, return_inverse=True)
        linear_indices, quadratic_indices = inverse[:size_linear], inverse[size_linear:]
        freqs = unique
    else:
        linear_indices = wa.pop("linear_indices")
        quadratic_indices = wa.pop("quadratic_indices")
        for key in ("frequency_nodes_linear", "frequency_nodes_quadratic"):
            wa.pop(key, None)
        freqs = wa.pop("frequency_nodes")
    pols = _sequence_polarizations(freqs, mass_1, mass_2, luminosity_distance, a_1, tilt_1, a_2, tilt_2, theta_jn,
                                   phase, lambda_1, lambda_2, **wa)
    if pols is None:
        return None
    return dict(linear=dict(plus=pols["plus"][linear_indices], cross=pols["cross"][linear_indices]),
                quadratic=dict(plus=pols["plus"][quadratic_indices], cross=pols["cross"][quadratic_indices]))


def binary_black_hole_roq(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2, tilt_2,
                          phi_jl, theta_jn, phase, **waveform_arguments):
    """source.py:693-706 (reference_frequency defaults to 20 Hz here)."""
    wa = dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0, catch_waveform_errors=False)
    wa.update(waveform_arguments)
    return _base_roq_waveform(mass_1, mass_2, luminosity_distance, a_1, tilt_1, a_2, tilt_2, theta_jn, phase,
                              0.0, 0.0, **{k: v for k, v in wa.items() if k not in _RB_DROP})


def binary_neutron_star_roq(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2, tilt_2,
                            phi_jl, lambda_1, lambda_2, theta_jn, phase, **waveform_arguments):
    """source.py:709-721."""
    wa = dict(waveform_approximant="TaylorF2", reference_frequency=20.0, catch_waveform_errors=False)
    wa.update(waveform_arguments)
    return _base_roq_waveform(mass_1, mass_2, luminosity_distance, a_1, tilt_1, a_2, tilt_2, theta_jn, phase,
                              lambda_1, lambda_2, **{k: v for k, v in wa.items() if k not in _RB_DROP})


# --------------------------------------------------------------------------------------
# relative binning (relative.py)
# --------------------------------------------------------------------------------------
class OracleRelativeBinning(ocl.OracleLikelihood):
    """relative.py:105-430 for reference_frame='sky', time_reference='geocent', no fiducial update."""

    def __init__(self, interferometers, fiducial_parameters, source_model=lal_binary_black_hole_relative_binning,
                 waveform_arguments=None, chi=1, epsilon=0.5, **kw):
        super().__init__(interferometers, source_model=source_model, waveform_arguments=waveform_arguments, **kw)
        self.fiducial_parameters = dict(fiducial_parameters)
        if self.time_marginalization:
            self.fiducial_parameters["geocent_time"] = self.start_time
        if self.distance_marginalization:
            self.fiducial_parameters["luminosity_distance"] = self._ref_dist
        if self.phase_marginalization:
            self.fiducial_parameters["phase"] = 0.0
        self.chi, self.epsilon = chi, epsilon
        self.gamma = np.array([-5 / 3, -2 / 3, 1, 5 / 3, 7 / 3])
        self.maximum_frequency = self.frequency_array[-1]
        self.set_fiducial_waveforms(self.fiducial_parameters)
        self.setup_bins()
        self.compute_summary_data()

    def polarizations(self, parameters, fiducial=0):
        p = self.parameter_conversion(parameters)
        args = {k: p[k] for k in ocl.SOURCE_ARGS if k in p}
        if "neutron" in getattr(self.source_model, "__name__", ""):
            for k in ("lambda_1", "lambda_2"):
                args[k] = p.get(k, 0.0)
        wa = dict(self.waveform_arguments)
        wa["fiducial"] = fiducial
        return self.source_model(self.frequency_array, **args, **wa)

    def set_fiducial_waveforms(self, parameters):
        """relative.py:242-263."""
        pols = self.polarizations(dict(parameters), fiducial=1)
        last = np.where(pols["plus"] != 0j)[0][-1]
        self.maximum_frequency = self.frequency_array[last]
        conv = self.parameter_conversion(parameters)
        self.per_detector_fiducial_waveforms = {}
        for ifo in self.ifos:
            wf = ifo.get_detector_response(pols, conv)
            wf = wf * (ifo.frequency_array <= self.maximum_frequency)
            self.per_detector_fiducial_waveforms[ifo.name] = wf

    def setup_bins(self):
        """relative.py:179-240."""
        frequency_array = self.frequency_array
        gamma = self.gamma[:, np.newaxis]
        maximum_frequency = frequency_array[0]
        minimum_frequency = frequency_array[-1]
        for ifo in self.ifos:
            maximum_frequency = max(maximum_frequency, ifo.maximum_frequency)
            minimum_frequency = min(minimum_frequency, ifo.minimum_frequency)
        maximum_frequency = min(maximum_frequency, self.maximum_frequency)
        useful = frequency_array[(frequency_array >= minimum_frequency) & (frequency_array <= maximum_frequency)]
        d_alpha = self.chi * 2 * np.pi / np.abs((minimum_frequency ** gamma) * np.heaviside(-gamma, 1)
                                                - (maximum_frequency ** gamma) * np.heaviside(gamma, 1))
        d_phi = np.sum(np.sign(gamma) * d_alpha * useful ** gamma, axis=0)
        d_phi_from_start = d_phi - d_phi[0]
        number_of_bins = int(d_phi_from_start[-1] // self.epsilon)
        bin_inds, bin_freqs = [], []
        last_index = -1
        for i in range(number_of_bins + 1):
            bin_index = np.where(d_phi_from_start >= ((i / number_of_bins) * d_phi_from_start[-1]))[0][0]
            if bin_index == last_index:
                continue
            bin_freq = useful[bin_index]
            last_index = bin_index
            bin_index = np.where(frequency_array >= bin_freq)[0][0]
            bin_inds.append(bin_index)
            bin_freqs.append(bin_freq)
        self.bin_inds = np.array(bin_inds, dtype=int)
        self.bin_sizes = np.diff(bin_inds)
        self.bin_sizes[-1] += 1
        self.bin_freqs = np.array(bin_freqs)
        self.number_of_bins = len(self.bin_inds) - 1
        self.waveform_arguments["frequency_bin_edges"] = self.bin_freqs
        self.bin_widths = self.bin_freqs[1:] - self.bin_freqs[:-1]
        self.bin_centers = (self.bin_freqs[1:] + self.bin_freqs[:-1]) / 2
        self.per_detector_fiducial_waveform_points = {
            ifo.name: self.per_detector_fiducial_waveforms[ifo.name][self.bin_inds] for ifo in self.ifos}

    def compute_summary_data(self):
        """relative.py:319-363."""
        self.summary_data = {}
        for ifo in self.ifos:
            mask = ifo.frequency_mask
            mf = ifo.frequency_array[mask]
            masked_bin_inds = [int(np.where(mf == edge)[0][0]) for edge in self.bin_freqs]
            if masked_bin_inds[-1] < len(mf) - 1:
                masked_bin_inds[-1] += 1
            strain = ifo.frequency_domain_strain[mask]
            h0 = self.per_detector_fiducial_waveforms[ifo.name][mask]
            psd = ifo.power_spectral_density_array[mask]
            a0, b0, a1, b1 = np.zeros((4, self.number_of_bins), dtype=complex)
            for i in range(self.number_of_bins):
                s, e = masked_bin_inds[i], masked_bin_inds[i + 1]
                central = (mf[s] + mf[e]) / 2
                sl = slice(s, e)
                delta = mf[sl] - central
                norm = 4 / self.duration
                a0[i] = norm * np.sum(h0[sl].conj() * strain[sl] / psd[sl])
                b0[i] = norm * np.sum(h0[sl].conj() * h0[sl] / psd[sl])
                a1[i] = norm * np.sum(h0[sl].conj() * (strain[sl] * delta) / psd[sl])
                b1[i] = norm * np.sum(h0[sl].conj() * (h0[sl] * delta) / psd[sl])
            self.summary_data[ifo.name] = (a0, a1, b0, b1)

    def waveform_ratio(self, pols, ifo, parameters):
        """relative.py:365-378."""
        strain = ifo.get_detector_response(pols, parameters, frequencies=self.bin_freqs)
        ratio = strain / self.per_detector_fiducial_waveform_points[ifo.name]
        r0 = (ratio[1:] + ratio[:-1]) / 2
        r1 = (ratio[1:] - ratio[:-1]) / self.bin_widths
        return r0, r1

    def calculate_snrs(self, pols, ifo, parameters):
        """relative.py:398-430."""
        r0, r1 = self.waveform_ratio(pols, ifo, parameters)
        a0, a1, b0, b1 = self.summary_data[ifo.name]
        d_inner_h = (a0 * r0.conj() + a1 * r1.conj()).sum()
        hh = (b0 * abs(r0) ** 2 + 2 * b1 * (r0 * r1.conj()).real).sum().real
        arr = None
        if self.time_marginalization:
            idxs = slice(self.bin_inds[0], self.bin_inds[-1] + 1)
            f = ifo.frequency_array
            ratio = np.zeros(f.shape[0], dtype=complex)
            ratio[idxs] = np.repeat(r0, self.bin_sizes) + np.repeat(r1, self.bin_sizes) * (
                f[idxs] - np.repeat(self.bin_centers, self.bin_sizes))
            full = self.per_detector_fiducial_waveforms[ifo.name] * ratio
            with np.errstate(invalid="ignore", divide="ignore"):
                arr = 4 / self.duration * np.fft.fft(
                    full[0:-1] * ifo.frequency_domain_strain.conj()[0:-1] / ifo.power_spectral_density_array[0:-1])
        return d_inner_h, hh, arr


# --------------------------------------------------------------------------------------
# ROQ (roq.py)
# --------------------------------------------------------------------------------------
def roq_time_resolution(ifos, optimal_snrs=None):
    """roq.py:1165-1229 _get_time_resolution.  optimal_snrs: per-detector injected optimal SNR as the reference
    reads it from ifo.meta_data['optimal_SNR'] (30 when absent).  The PSD / frequency array used are those of
    the LAST interferometer (the reference's loop variable leaks, roq.py:1216-1217)."""
    from scipy.integrate import simpson

    def calc_fhigh(freq, psd, scaling=20.):
        integrand1 = np.power(freq, -7. / 3) / psd
        integral1 = simpson(y=integrand1, x=freq)
        integrand3 = np.power(freq, 2. / 3.) / (psd * integral1)
        f_3_bar = simpson(y=integrand3, x=freq)
        return scaling * f_3_bar ** (1 / 3)

    def c_f_scaling(snr):
        return (np.pi ** 2 * snr ** 2 / 6) ** (1 / 3)

    inj_snr_sq = 0
    for i, ifo in enumerate(ifos):
        snr = 30 if optimal_snrs is None else optimal_snrs[i]
        inj_snr_sq += max(10, snr) ** 2
    ifo = ifos[-1]
    psd = ifo.power_spectral_density_array[ifo.frequency_mask]
    freq = ifo.frequency_array[ifo.frequency_mask]
    fhigh = calc_fhigh(freq, psd, scaling=c_f_scaling(inj_snr_sq ** 0.5))
    delta_t = fhigh ** -1
    delta_t = delta_t / 5
    n = max(ifo.duration / delta_t, ifo.frequency_array[-1] * ifo.duration + 1)
    n = int(2 ** np.ceil(np.log2(n)))
    return ifo.duration / n


class OracleROQ(ocl.OracleLikelihood):
    """roq.py:112-651 + 736-1004 for a single ndarray linear and quadratic basis (no multibanding, no
    roq_params, no basis selection), reference_frame='sky', time_reference='geocent'.

    linear_matrix: [n_masked_freq, n_linear], quadratic_matrix: [n_masked_freq, n_quadratic] (the layout the
    reference accepts for ndarray bases, roq.py:362-363)."""

    def __init__(self, interferometers, linear_matrix, quadratic_matrix, frequency_nodes_linear,
                 frequency_nodes_quadratic, time_prior, source_model=binary_black_hole_roq, waveform_arguments=None,
                 time_space=None, delta_tc=None, optimal_snrs=None, weights=None, **kw):
        self._roq_delta_tc = delta_tc
        self._roq_time_space = time_space
        self._optimal_snrs = optimal_snrs
        self.weights = {}
        self._time_prior_roq = time_prior
        super().__init__(interferometers, source_model=source_model, waveform_arguments=waveform_arguments,
                         time_prior=time_prior, **kw)
        if self.time_marginalization:
            # roq.py:320-331 (overrides base.py:1027-1035)
            if self._roq_delta_tc is None:
                self._roq_delta_tc = self.time_resolution()
            tcmin, tcmax = time_prior.minimum, time_prior.maximum
            n_t = int(np.ceil((tcmax - tcmin) / self._roq_delta_tc))
            self._delta_tc = (tcmax - tcmin) / n_t
            self._times = tcmin + self._delta_tc / 2. + np.arange(n_t) * self._delta_tc
            self._beam_pattern_reference_time = (tcmin + tcmax) / 2.
        self.frequency_nodes_linear = np.asarray(frequency_nodes_linear, dtype=float)
        self.frequency_nodes_quadratic = np.asarray(frequency_nodes_quadratic, dtype=float)
        unique, inverse = np.unique(np.hstack((self.frequency_nodes_linear, self.frequency_nodes_quadratic)),
                                    return_inverse=True)
        self.frequency_nodes = unique
        self.linear_indices = inverse[:len(self.frequency_nodes_linear)]
        self.quadratic_indices = inverse[len(self.frequency_nodes_linear):]
        self.waveform_arguments.update(frequency_nodes=self.frequency_nodes, linear_indices=self.linear_indices,
                                       quadratic_indices=self.quadratic_indices)
        if weights is not None:
            # precomputed weights (the reference accepts them too, roq.py:112-122 `weights=`): dict with time_samples,
            # {IFO}_linear [n_time, n_linear], {IFO}_quadratic [n_quadratic]
            self.weights = dict(weights)
        else:
            self._set_weights(np.asarray(linear_matrix).T, np.asarray(quadratic_matrix).T)

    def time_resolution(self):
        if self._roq_time_space is not None:
            return self._roq_time_space
        return roq_time_resolution(self.ifos, self._optimal_snrs)

    def _set_weights(self, linear_basis, quadratic_basis):
        """roq.py:736-767, 849-916, 976-1004.  linear_basis [n_linear, n_freq_masked]."""
        time_space = self.time_resolution()
        n_time = int(self.duration / time_space)
        light = 2 * RADIUS_OF_EARTH / SPEED_OF_LIGHT + 5 * time_space
        start_idx = max(0, int(np.floor((self._time_prior_roq.minimum - light - self.start_time) / time_space)))
        end_idx = min(n_time - 1, int(np.ceil((self._time_prior_roq.maximum + light - self.start_time) / time_space)))
        self.weights["time_samples"] = np.arange(start_idx, end_idx + 1) * float(time_space)
        ts = self.weights["time_samples"]
        space = ts[1] - ts[0]
        n_time = int(self.duration / space)
        s_idx, e_idx = int(ts[0] / space), int(ts[-1] / space)
        for ifo in self.ifos:
            mask = ifo.frequency_mask
            n_masked = int(mask.sum())
            if linear_basis.shape[1] != n_masked:
                raise ValueError("Mismatch between ROQ basis and frequency array for {}".format(ifo.name))
            nonzero = np.arange(n_masked) + int(ifo.minimum_frequency * self.duration)
            data_over_psd = ifo.frequency_domain_strain[mask] / ifo.power_spectral_density_array[mask]
            lw = np.zeros((linear_basis.shape[0], len(ts)), dtype=complex)
            buf = np.zeros(n_time, dtype=complex)
            for i in range(linear_basis.shape[0]):
                buf[:] = 0
                buf[nonzero] = data_over_psd * linear_basis[i].conj()
                lw[i] = np.fft.ifft(buf)[s_idx:e_idx + 1]
            self.weights[ifo.name + "_linear"] = lw.T * (4. * n_time / self.duration)
            inv_psd = 1 / ifo.power_spectral_density_array[mask]
            self.weights[ifo.name + "_quadratic"] = 4. / self.duration * quadratic_basis.real @ inv_psd

    @staticmethod
    def _interp_five_samples(time_samples, values, time):
        """roq.py:576-602."""
        r1 = (-values[0] + 8. * values[1] - 14. * values[2] + 8. * values[3] - values[4]) / 4.
        r2 = values[2] - 2. * values[3] + values[4]
        a = (time_samples[3] - time) / max(time_samples[1] - time_samples[0], 1e-12)
        b = 1. - a
        c = (a ** 3. - a) / 6.
        d = (b ** 3. - b) / 6.
        return a * values[2] + b * values[3] + c * r1 + d * r2

    def _d_inner_h_array(self, times, h_linear, name):
        """roq.py:604-651."""
        ts = self.weights["time_samples"]
        space = ts[1] - ts[0]
        per = (times - ts[0]) / space
        closest = np.floor(per).astype(int)
        w = self.weights[name + "_linear"]
        hc = h_linear.conj()
        if (times[1] - times[0]) / space > 5:
            m2, m1, z0, p1, p2 = (w[closest + k] @ hc for k in (-2, -1, 0, 1, 2))
        else:
            full = w @ hc
            m2, m1, z0, p1, p2 = (full[closest + k] for k in (-2, -1, 0, 1, 2))
        b = per - closest
        a = 1. - b
        c = (a ** 3. - a) / 6.
        d = (b ** 3. - b) / 6.
        r1 = (-m2 + 8. * m1 - 14. * z0 + 8. * p1 - p2) / 4.
        r2 = z0 - 2. * p1 + p2
        return a * z0 + b * p1 + c * r1 + d * r2

    def calculate_snrs(self, pols, ifo, parameters):
        """roq.py:467-549."""
        time_ref = self._beam_pattern_reference_time if self.time_marginalization else parameters["geocent_time"]
        fp, fc = ifo.antenna_response(parameters["ra"], parameters["dec"], time_ref, parameters["psi"])
        h_linear = pols["linear"]["plus"] * fp + pols["linear"]["cross"] * fc
        h_quadratic = pols["quadratic"]["plus"] * fp + pols["quadratic"]["cross"] * fc
        if ifo.calibration is not None:
            cal = ifo.calibration.get_calibration_factor(self.frequency_nodes, prefix=f"recalib_{ifo.name}_",
                                                         **parameters)
            h_linear = h_linear * cal[self.linear_indices]
            h_quadratic = h_quadratic * cal[self.quadratic_indices]
        hh = np.vdot(np.abs(h_quadratic) ** 2, self.weights[ifo.name + "_quadratic"]).real
        dt = ocl.time_delay_from_geocenter(ifo.vertex, parameters["ra"], parameters["dec"], time_ref)
        ifo_time = (parameters["geocent_time"] - ifo.start_time) + dt
        ts = self.weights["time_samples"]
        closest = int(np.floor((ifo_time - ts[0]) / (ts[1] - ts[0])))
        indices = np.array([closest + ii for ii in (-2, -1, 0, 1, 2)])
        in_bounds = (indices[0] >= 0) & (indices[-1] < len(ts))
        indices = np.clip(indices, 0, len(ts) - 1)
        tc_arr = np.einsum("i,ji->j", np.conj(h_linear), self.weights[ifo.name + "_linear"][indices])
        d_inner_h = self._interp_five_samples(ts[indices], tc_arr, ifo_time)
        with np.errstate(divide="ignore"):
            d_inner_h = d_inner_h + np.log(in_bounds)
        arr = None
        if self.time_marginalization:
            ifo_times = self._times - ifo.start_time + dt
            if self.jitter_time:
                ifo_times = ifo_times + parameters["time_jitter"]
            arr = self._d_inner_h_array(ifo_times, h_linear, ifo.name)
        return d_inner_h, hh, arr

    def time_marginalized_likelihood(self, d_inner_h_tc_array, hh, parameters):
        """base.py:794-820 with the ROQ's own time grid (roq.py:320-331)."""
        return super().time_marginalized_likelihood(d_inner_h_tc_array, hh, parameters)


# --------------------------------------------------------------------------------------
# synthetic ROQ bases (the reference vendors none: test/gw/likelihood_test.py:334-348 downloads them)
# --------------------------------------------------------------------------------------
def _empirical_interpolant(training, n_basis):
    """SVD basis + DEIM nodes.  training [n_train, n_freq] -> (B [n_freq, n_basis] with B[node_i, j] = delta_ij,
    node indices)."""
    u, s, vh = np.linalg.svd(training, full_matrices=False)
    v = vh[:n_basis].T                     # [n_freq, n_basis]
    nodes = [int(np.argmax(np.abs(v[:, 0])))]
    for j in range(1, n_basis):
        c = np.linalg.solve(v[nodes, :j], v[nodes, j])
        r = v[:, j] - v[:, :j] @ c
        nodes.append(int(np.argmax(np.abs(r))))
    nodes = np.array(nodes)
    order = np.argsort(nodes)
    nodes = nodes[order]
    b = v @ np.linalg.inv(v[nodes, :])
    return b, nodes


def build_synthetic_roq_basis(frequencies, waveform_fn, draws, n_linear, n_quadratic):
    """Empirical-interpolation bases for h(f) (linear) and |h(f)|^2 (quadratic) from training waveforms
    ``waveform_fn(draw_i) -> complex array on frequencies``.  Returns dict(linear_matrix, quadratic_matrix,
    frequency_nodes_linear, frequency_nodes_quadratic)."""
    train = np.array([waveform_fn(d) for d in draws])
    train = train / np.linalg.norm(train, axis=1, keepdims=True)
    bl, nl = _empirical_interpolant(train, n_linear)
    bq, nq = _empirical_interpolant((np.abs(train) ** 2).astype(complex), n_quadratic)
    return dict(linear_matrix=bl, quadratic_matrix=bq, frequency_nodes_linear=frequencies[nl],
                frequency_nodes_quadratic=frequencies[nq])
