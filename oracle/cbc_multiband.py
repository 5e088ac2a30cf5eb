"""TEST INFRASTRUCTURE - CPU restatement (numpy) of the reference's multi-banded likelihood
MBGravitationalWaveTransient (bilby/gw/likelihood/multiband.py, S. Morisaki arXiv:2104.07813) for the
linear-interpolation form of (h, h), reference_frame='sky', time_reference='geocent', no time marginalisation.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path never does.
Every function cites the reference lines it follows.  Pinned against golden vectors produced by the UNMODIFIED
reference class (oracle/tools/make_golden_multiband.py -> tests/golden/multiband_*.npz).
"""
import math

import numpy as np

from . import cbc_likelihood as ocl
from . import cbc_reduced as ocr

SOLAR_MASS = 1.988409870698050731911960804878414216e30      # bilby/core/utils/constants.py
GRAVITATIONAL_CONSTANT = 6.6743e-11
SPEED_OF_LIGHT = 299792458.0
RADIUS_OF_EARTH = 6378136.6


# --------------------------------------------------------------------------------------
# source models on the banded frequency points (source.py:901-1065 -> :1068-1140)
# --------------------------------------------------------------------------------------
def binary_black_hole_frequency_sequence(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12,
                                         a_2, tilt_2, phi_jl, theta_jn, phase, **kwargs):
    """source.py:901-979: the waveform at waveform_kwargs['frequencies'] (every point evaluated)."""
    wa = dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, catch_waveform_errors=False)
    wa.update(kwargs)
    freqs = wa.pop("frequencies")
    for key in ocr._RB_DROP + ("minimum_frequency", "maximum_frequency"):
        wa.pop(key, None)
    return ocr._sequence_polarizations(freqs, mass_1, mass_2, luminosity_distance, a_1, tilt_1, a_2, tilt_2,
                                       theta_jn, phase, 0.0, 0.0, **wa)


def binary_neutron_star_frequency_sequence(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12,
                                           a_2, tilt_2, phi_jl, lambda_1, lambda_2, theta_jn, phase, **kwargs):
    """source.py:982-1065."""
    wa = dict(waveform_approximant="TaylorF2", reference_frequency=50.0, catch_waveform_errors=False)
    wa.update(kwargs)
    freqs = wa.pop("frequencies")
    for key in ocr._RB_DROP + ("minimum_frequency", "maximum_frequency"):
        wa.pop(key, None)
    return ocr._sequence_polarizations(freqs, mass_1, mass_2, luminosity_distance, a_1, tilt_1, a_2, tilt_2,
                                       theta_jn, phase, lambda_1, lambda_2, **wa)


def round_up_to_power_of_two(x):
    """bilby/core/utils/calculus.py round_up_to_power_of_two: 2**ceil(log2(x))."""
    return 2 ** math.ceil(np.log2(x))


class OracleMultiband(ocl.OracleLikelihood):
    """multiband.py:93-127 (constructor), :311-320 (setup_multibanding), :728-807 (calculate_snrs)."""

    def __init__(self, interferometers, reference_chirp_mass, source_model=binary_black_hole_frequency_sequence,
                 waveform_arguments=None, highest_mode=2, accuracy_factor=5, time_offset=None, delta_f_end=None,
                 maximum_banding_frequency=None, minimum_banding_duration=0.0, geocent_time_prior=None,
                 linear_interpolation=True, **kw):
        super().__init__(interferometers, source_model=source_model, waveform_arguments=waveform_arguments, **kw)
        self.reference_chirp_mass = reference_chirp_mass
        self.mc_sec = GRAVITATIONAL_CONSTANT * reference_chirp_mass * SOLAR_MASS / SPEED_OF_LIGHT ** 3   # :133-134
        self.highest_mode = highest_mode
        self.accuracy_factor = accuracy_factor
        safety = RADIUS_OF_EARTH / SPEED_OF_LIGHT                                   # :208-212 (geocent time)
        end = self.start_time + self.duration
        if time_offset is None:                                                    # :213-225
            time_offset = (end - geocent_time_prior[0] + safety) if geocent_time_prior is not None else 2.12
        if delta_f_end is None:                                                    # :246-258
            delta_f_end = 100 / (end - geocent_time_prior[1] - safety) if geocent_time_prior is not None else 53.0
        self.time_offset, self.delta_f_end = time_offset, delta_f_end
        fmax_tmp = (15 / 968) ** (3 / 5) * (highest_mode / (2 * np.pi)) ** (8 / 5) / self.mc_sec     # :273-276
        if maximum_banding_frequency is not None and maximum_banding_frequency < fmax_tmp:
            fmax_tmp = maximum_banding_frequency
        self.maximum_banding_frequency = fmax_tmp
        self.minimum_banding_duration = minimum_banding_duration
        self.minimum_frequency = min(ifo.minimum_frequency for ifo in self.ifos)
        self.maximum_frequency = max(ifo.maximum_frequency for ifo in self.ifos)
        self._setup_frequency_bands()
        self._setup_integers()
        self._setup_waveform_frequency_points()
        self._setup_linear_coefficients()
        self.linear_interpolation = linear_interpolation
        if linear_interpolation:
            self._setup_quadratic_coefficients_linear_interp()
        else:
            self._setup_quadratic_coefficients_ifft_fft()
        if self.time_marginalization:
            self._setup_time_marginalization_multiband()

    def _setup_time_marginalization_multiband(self):
        """:714-726."""
        n = int(self.Nbs[-1]) // 2
        self._delta_tc = self.durations[0] / n
        self._times = self.start_time + np.arange(n) * self._delta_tc
        self._full_d_h = np.zeros(n, dtype=complex)
        self._full_to_multiband = [int(f * self.durations[0]) for f in self.banded_frequency_points]
        self._beam_pattern_reference_time = (self.time_prior.minimum + self.time_prior.maximum) / 2
        for ifo in self.ifos:
            ifo.reference_time = self._beam_pattern_reference_time

    # ---- 0PN time to merger (:322-360)
    def _tau(self, f):
        f_22 = 2 * f / self.highest_mode
        return 5 / 256 * self.mc_sec * (np.pi * self.mc_sec * f_22) ** (-8 / 3)

    def _dtaudf(self, f):
        f_22 = 2 * f / self.highest_mode
        return -5 / 96 * self.mc_sec * (np.pi * self.mc_sec * f_22) ** (-8.0 / 3.0) / f

    def _find_starting_frequency(self, duration, fnow):
        """:362-400, bisection on conditions (10) and (51) of the paper."""
        def above(f):
            c1 = duration - self.time_offset - self._tau(f) - self.accuracy_factor * np.sqrt(-self._dtaudf(f)) > 0
            c2 = f - 1.0 / np.sqrt(-self._dtaudf(f)) - fnow > 0
            return c1 and c2
        fmin, fmax = fnow, self.maximum_banding_frequency
        if not above(fmax):
            return None, None
        f = None
        while fmax - fmin > 1e-2 / duration:
            f = (fmin + fmax) / 2.0
            if above(f):
                fmax = f
            else:
                fmin = f
        return f, 1.0 / np.sqrt(-self._dtaudf(f))

    def _setup_frequency_bands(self):
        """:402-426."""
        self.durations = [self.duration]
        fb_dfb = [[self.minimum_frequency, 0.0]]
        dnext = self.duration / 2
        while dnext > max(self.time_offset, self.minimum_banding_duration):
            fnow = fb_dfb[-1][0]
            fnext, dfnext = self._find_starting_frequency(dnext, fnow)
            if fnext is not None and fnext < min(self.maximum_frequency, self.maximum_banding_frequency):
                self.durations.append(dnext)
                fb_dfb.append([fnext, dfnext])
                dnext /= 2
            else:
                break
        fb_dfb.append([self.maximum_frequency + self.delta_f_end, self.delta_f_end])
        self.durations = np.array(self.durations)
        self.fb_dfb = np.array(fb_dfb)
        self.number_of_bands = len(self.durations)

    def _setup_integers(self):
        """:428-447."""
        self.Nbs, self.Mbs, self.Ks_Ke = [], [], []
        for b in range(self.number_of_bands):
            dnow = self.durations[b]
            fnow, dfnow = self.fb_dfb[b]
            fnext = self.fb_dfb[b + 1][0]
            nb = max(round_up_to_power_of_two(2.0 * (fnext * self.duration + 1.0)), 2 ** b)
            self.Nbs.append(nb)
            self.Mbs.append(nb // 2 ** b)
            self.Ks_Ke.append([math.ceil((fnow - dfnow) * dnow), math.floor(fnext * dnow)])
        self.Nbs, self.Mbs, self.Ks_Ke = np.array(self.Nbs), np.array(self.Mbs), np.array(self.Ks_Ke)

    def _setup_waveform_frequency_points(self):
        """:449-478."""
        pts, idxs, start = [], [], 0
        for b in range(self.number_of_bands):
            ks, ke = self.Ks_Ke[b]
            pts.append(np.arange(ks, ke + 1) / self.durations[b])
            idxs.append([start, start + ke - ks])
            start += ke - ks + 1
        self.banded_frequency_points = np.concatenate(pts)
        self.start_end_idxs = np.array(idxs)
        unique, inverse = np.unique(self.banded_frequency_points, return_inverse=True)
        self.waveform_arguments["frequencies"] = unique
        self.unique_to_original_frequencies = inverse

    def _get_window_sequence(self, delta_f, start_idx, length, b):
        """:480-527: Hann-tapered window of band b sampled at (start_idx + i) delta_f."""
        fnow, dfnow = self.fb_dfb[b]
        fnext, dfnext = self.fb_dfb[b + 1]
        w = np.zeros(length)
        inc0 = int(np.clip(math.floor((fnow - dfnow) / delta_f) - start_idx + 1, 0, length))
        one0 = int(np.clip(math.ceil(fnow / delta_f) - start_idx, 0, length))
        dec0 = int(np.clip(math.floor((fnext - dfnext) / delta_f) - start_idx + 1, 0, length))
        dec1 = int(np.clip(math.ceil(fnext / delta_f) - start_idx, 0, length))
        w[one0:dec0] = 1.0
        if inc0 < one0:
            fr = (np.arange(inc0, one0) + start_idx) * delta_f
            w[inc0:one0] = (1.0 + np.cos(np.pi * (fr - fnow) / dfnow)) / 2.0
        if dec0 < dec1:
            fr = (np.arange(dec0, dec1) + start_idx) * delta_f
            w[dec0:dec1] = (1.0 - np.cos(np.pi * (fr - fnext) / dfnext)) / 2.0
        return w

    def _setup_linear_coefficients(self):
        """:529-549: data/PSD down-sampled and shortened per band (irfft, keep the last M^(b) samples, rfft)."""
        self.linear_coeffs = {}
        n_full = self.Nbs[-1]
        for ifo in self.ifos:
            fddata = np.zeros(n_full // 2 + 1, dtype=complex)
            mask = ifo.frequency_mask
            fddata[:len(ifo.frequency_domain_strain)][mask[:len(fddata)]] += \
                ifo.frequency_domain_strain[mask] / ifo.power_spectral_density_array[mask]
            out = []
            for b in range(self.number_of_bands):
                ks, ke = self.Ks_Ke[b]
                windows = self._get_window_sequence(1.0 / self.durations[b], ks, ke - ks + 1, b)
                band = np.copy(fddata[:int(self.Nbs[b] / 2 + 1)])
                band[-1] = 0.0
                td = np.fft.irfft(band)[-self.Mbs[b]:]
                fd = np.fft.rfft(td)[ks:ke + 1]
                out.append((4.0 / self.durations[b]) * windows * np.conj(fd))
            self.linear_coeffs[ifo.name] = np.concatenate(out)

    def _setup_quadratic_coefficients_linear_interp(self):
        """:551-611: |h|^2 linearly interpolated between banded points, summed against window/PSD on the full grid."""
        self.quadratic_coeffs = {ifo.name: [] for ifo in self.ifos}
        t_full = float(self.duration)
        for b in range(self.number_of_bands):
            s, e = self.start_end_idxs[b]
            fpts = self.banded_frequency_points[s:e + 1]
            prefactor = 4 * self.durations[b] / t_full
            fnow, dfnow = self.fb_dfb[b]
            fnext = self.fb_dfb[b + 1][0]
            i0 = math.ceil((fnow - dfnow) * t_full)
            win = self._get_window_sequence(1 / t_full, i0, math.floor(fnext * t_full) - i0 + 1, b)
            for ifo in self.ifos:
                psd = ifo.power_spectral_density_array
                i1 = min(i0 + len(win) - 1, len(psd) - 1)
                msk = np.asarray(ifo.frequency_mask[i0:i1 + 1])
                wop = np.zeros(i1 + 1 - i0)
                wop[msk] = 1.0 / psd[i0:i1 + 1][msk]
                wop *= win[:len(wop)]
                coeffs = np.zeros(len(fpts))
                for k in range(len(coeffs) - 1):
                    lo = i0 if k == 0 else max(i0, math.ceil(t_full * fpts[k]))
                    hi = i1 if k == len(coeffs) - 2 else min(i1, math.ceil(t_full * fpts[k + 1]) - 1)
                    fs = np.arange(lo, hi + 1) / t_full
                    seg = wop[lo - i0:hi - i0 + 1]
                    coeffs[k] += prefactor * np.sum((fpts[k + 1] - fs) * seg)
                    coeffs[k + 1] += prefactor * np.sum((fs - fpts[k]) * seg)
                self.quadratic_coeffs[ifo.name].append(coeffs)
        for name in self.quadratic_coeffs:
            self.quadratic_coeffs[name] = np.concatenate(self.quadratic_coeffs[name])

    def _setup_quadratic_coefficients_ifft_fft(self):
        """:613-646."""
        n_full = int(self.Nbs[-1])
        nhat = [min(2 * int(mb), int(nb)) for mb, nb in zip(self.Mbs, self.Nbs)]
        self.Tbhats = [self.duration * nh / nb for nb, nh in zip(self.Nbs, nhat)]
        self.Ibcs = {ifo.name: [] for ifo in self.ifos}
        self.hbcs = {ifo.name: [] for ifo in self.ifos}
        self.wths = {ifo.name: [] for ifo in self.ifos}
        for ifo in self.ifos:
            inv = np.zeros(n_full // 2 + 1)
            psd, mask = ifo.power_spectral_density_array, ifo.frequency_mask
            inv[:len(psd)][mask[:len(inv)]] = 1 / psd[mask]
            for b in range(self.number_of_bands):
                imb = np.fft.irfft(inv[:int(self.Nbs[b]) // 2 + 1])
                half = nhat[b] // 2
                imbc = np.append(imb[:half + 1], imb[-(nhat[b] - half - 1):])
                self.Ibcs[ifo.name].append(np.fft.rfft(imbc))
                self.hbcs[ifo.name].append(np.zeros(nhat[b]))
                self.wths[ifo.name].append(np.zeros(int(self.Mbs[b]) // 2 + 1, dtype=complex))
        self.windows, self.square_root_windows = np.array([]), np.array([])
        for b in range(self.number_of_bands):
            ks, ke = self.Ks_Ke[b]
            ws = self._get_window_sequence(1.0 / self.durations[b], ks, ke - ks + 1, b)
            self.windows = np.append(self.windows, ws)
            self.square_root_windows = np.append(self.square_root_windows, np.sqrt(ws))

    def _optimal_snr_squared_ifft_fft(self, strain, ifo):
        """:766-787."""
        out = 0.0
        for b in range(self.number_of_bands):
            ks, ke = self.Ks_Ke[b]
            s0, e0 = self.start_end_idxs[b]
            mb = int(self.Mbs[b])
            if b == 0:
                out += (4.0 / self.duration) * np.vdot(
                    np.abs(strain[s0:e0 + 1]) ** 2,
                    ifo.frequency_mask[ks:ke + 1] * self.windows[s0:e0 + 1] / ifo.power_spectral_density_array[ks:ke + 1])
            else:
                self.wths[ifo.name][b][ks:ke + 1] = self.square_root_windows[s0:e0 + 1] * strain[s0:e0 + 1]
                self.hbcs[ifo.name][b][-mb:] = np.fft.irfft(self.wths[ifo.name][b])
                thbc = np.fft.rfft(self.hbcs[ifo.name][b])
                out += (4.0 / self.Tbhats[b]) * np.vdot(np.abs(thbc) ** 2, self.Ibcs[ifo.name][b].real)
        return out

    # ---- evaluation (:728-765)
    def calculate_snrs(self, pols, ifo, parameters):
        modes = {m: v[self.unique_to_original_frequencies] for m, v in pols.items()}
        strain = ifo.get_detector_response(modes, parameters, frequencies=self.banded_frequency_points)
        d_inner_h = np.conj(np.dot(strain, self.linear_coeffs[ifo.name]))
        if self.linear_interpolation:
            hh = np.vdot(np.abs(strain) ** 2, self.quadratic_coeffs[ifo.name])
        else:
            hh = self._optimal_snr_squared_ifft_fft(strain, ifo)
        arr = None
        if self.time_marginalization:                                  # :789-797
            idx = np.asarray(self._full_to_multiband)
            self._full_d_h[idx] *= 0
            for b in range(self.number_of_bands):
                s0, e0 = self.start_end_idxs[b]
                self._full_d_h[idx[s0:e0 + 1]] += strain[s0:e0 + 1] * self.linear_coeffs[ifo.name][s0:e0 + 1]
            arr = np.fft.fft(self._full_d_h)
        return d_inner_h, float(np.real(hh)), arr
