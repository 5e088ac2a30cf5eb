"""MBGravitationalWaveTransient on B200 (bilby/gw/likelihood/multiband.py:19-844; S. Morisaki, arXiv:2104.07813).

Set-up (once per data set, host numpy like the reference): frequency bands of geometrically decreasing duration
(multiband.py:402-447), the banded frequency points (:449-478), the linear coefficients from the down-sampled and
shortened data / PSD of every band (:529-549) and the quadratic coefficients of the linear-interpolation form of
(h, h) (:551-611).  Evaluation: the waveform at the banded points on the device and the two sums
<d|h> = conj(sum h L), <h|h> = sum |h|^2 Q in kernel K5's edge form (`bb_set_multiband`; one warp per sample, lanes
over the points); the marginalisations run in the usual epilogue.

Time marginalisation (multiband.py:714-726, 789-797): the reference's FFT of the scattered strain * linear_coeffs array is
evaluated for the times inside the geocent_time prior as a dense contraction on the FP64 tensor cores
(`bb_set_multiband_time_marginalization`, csrc/bb_reduced.cuh bb_mb_*).

IFFT-FFT form of (h, h) (`linear_interpolation=False`, multiband.py:613-646, 766-787): band 0 and the even bins of every
band's zero-padded spectrum are per-point weights (they ride in the same kernel as the linear-interpolation form); the
odd bins take two transforms per sample, detector and band on the device (csrc/bb_fft.cuh, bb_mb_band_*).

Not provided: weights files (HDF5; h5py is absent) - the `weights` dict round-trips.
"""
import math
import numbers

import numpy as np

from .. import _lib
from ..core.utils import logger, gravitational_constant, solar_mass, speed_of_light, radius_of_earth
from .likelihood import GravitationalWaveTransient


def _next_power_of_two(x):
    return 2 ** math.ceil(np.log2(x))


class MBGravitationalWaveTransient(GravitationalWaveTransient):
    def __init__(self, interferometers, waveform_generator, reference_chirp_mass=None, highest_mode=2,
                 linear_interpolation=True, accuracy_factor=5, time_offset=None, delta_f_end=None,
                 maximum_banding_frequency=None, minimum_banding_duration=0., weights=None,
                 distance_marginalization=False, phase_marginalization=False, priors=None,
                 time_marginalization=False, jitter_time=True, distance_marginalization_lookup_table=None,
                 reference_frame="sky", time_reference="geocenter", device=None):
        if getattr(waveform_generator.frequency_domain_source_model, "_bb_kind", None) != "frequency_sequence":
            raise TypeError("MBGravitationalWaveTransient needs one of the source models "
                            "binary_black_hole_frequency_sequence / binary_neutron_star_frequency_sequence")
        if isinstance(weights, str):
            raise NotImplementedError("multiband weights files are HDF5 (h5py is absent): pass the weights dict")
        self._mb_host = None
        super().__init__(interferometers=interferometers, waveform_generator=waveform_generator, priors=priors,
                         distance_marginalization=distance_marginalization,
                         phase_marginalization=phase_marginalization, time_marginalization=time_marginalization,
                         distance_marginalization_lookup_table=distance_marginalization_lookup_table,
                         jitter_time=jitter_time, reference_frame=reference_frame, time_reference=time_reference,
                         device=device)
        if weights is None:
            self.reference_chirp_mass = reference_chirp_mass
            self.highest_mode = highest_mode
            self.linear_interpolation = linear_interpolation
            self.accuracy_factor = accuracy_factor
            self.time_offset = time_offset
            self.delta_f_end = delta_f_end
            self.maximum_banding_frequency = maximum_banding_frequency
            self.minimum_banding_duration = minimum_banding_duration
            self.setup_multibanding()
        else:
            self.setup_multibanding_from_weights(weights)
        if self.time_marginalization:
            self._setup_time_marginalization_multiband()
        self._mb_host = self._pack_host_arrays()
        self._upload_multiband()

    # ---- validated settings (multiband.py:128-298) -------------------------------------------------
    @staticmethod
    def _number(value, name):
        if not isinstance(value, numbers.Number):
            raise TypeError(f"{name} must be a number")
        return value

    @property
    def reference_chirp_mass(self):
        return self._reference_chirp_mass

    @reference_chirp_mass.setter
    def reference_chirp_mass(self, value):
        if isinstance(value, numbers.Number):
            self._reference_chirp_mass = value
            return
        # multiband.py:140-158: fall back to the prior's minimum chirp mass
        minimum = getattr(self.priors, "minimum_chirp_mass", None)
        if minimum is None:
            raise TypeError(f"priors: {self.priors} cannot provide the minimum chirp mass; pass reference_chirp_mass")
        self._reference_chirp_mass = minimum
        logger.info(f"reference_chirp_mass is automatically set to prior minimum of chirp mass: {minimum}.")

    @property
    def reference_chirp_mass_in_second(self):
        return gravitational_constant * self._reference_chirp_mass * solar_mass / speed_of_light ** 3.

    @property
    def highest_mode(self):
        return self._highest_mode

    @highest_mode.setter
    def highest_mode(self, value):
        self._highest_mode = self._number(value, "highest_mode")

    @property
    def linear_interpolation(self):
        return self._linear_interpolation

    @linear_interpolation.setter
    def linear_interpolation(self, value):
        if not isinstance(value, (bool, np.bool_)):
            raise TypeError("linear_interpolation must be a bool")
        self._linear_interpolation = value

    @property
    def accuracy_factor(self):
        return self._accuracy_factor

    @accuracy_factor.setter
    def accuracy_factor(self, value):
        self._accuracy_factor = self._number(value, "accuracy_factor")

    def _time_parameter_and_safety(self):
        key = self.time_reference + "_time"
        light_time = radius_of_earth / speed_of_light
        return key, (light_time if key == "geocent_time" else 2 * light_time)

    @property
    def time_offset(self):
        return self._time_offset

    @time_offset.setter
    def time_offset(self, value):
        """multiband.py:197-225: (end of data) - (earliest arrival) + light travel time, 2.12 s without a prior."""
        key, safety = self._time_parameter_and_safety()
        if value is not None:
            self._time_offset = self._number(value, "time_offset")
        elif self.priors is not None and key in self.priors:
            ifos = self.interferometers
            self._time_offset = ifos.start_time + ifos.duration - self.priors[key].minimum + safety
        else:
            self._time_offset = 2.12
            logger.warning(f"time offset can not be inferred. Use the standard time offset of {self._time_offset} seconds.")

    @property
    def delta_f_end(self):
        return self._delta_f_end

    @delta_f_end.setter
    def delta_f_end(self, value):
        """multiband.py:230-258: 100 / (minimum time offset), 53 Hz without a prior."""
        key, safety = self._time_parameter_and_safety()
        if value is not None:
            self._delta_f_end = self._number(value, "delta_f_end")
        elif self.priors is not None and key in self.priors:
            ifos = self.interferometers
            self._delta_f_end = 100 / (ifos.start_time + ifos.duration - self.priors[key].maximum - safety)
        else:
            self._delta_f_end = 53.
            logger.warning(f"delta_f_end can not be inferred. Use the standard delta_f_end of {self._delta_f_end} Hz.")

    @property
    def maximum_banding_frequency(self):
        return self._maximum_banding_frequency

    @maximum_banding_frequency.setter
    def maximum_banding_frequency(self, value):
        """multiband.py:264-286: capped where f - 1/sqrt(-dtau/df) stops increasing (0PN)."""
        cap = ((15 / 968) ** (3 / 5) * (self.highest_mode / (2 * np.pi)) ** (8 / 5)
               / self.reference_chirp_mass_in_second)
        if value is not None:
            if self._number(value, "maximum_banding_frequency") < cap:
                cap = value
            else:
                logger.warning(f"The input maximum_banding_frequency is too large. It is set to be {cap} Hz.")
        self._maximum_banding_frequency = cap

    @property
    def minimum_banding_duration(self):
        return self._minimum_banding_duration

    @minimum_banding_duration.setter
    def minimum_banding_duration(self, value):
        self._minimum_banding_duration = self._number(value, "minimum_banding_duration")

    @property
    def minimum_frequency(self):
        return np.min([i.minimum_frequency for i in self.interferometers])

    @property
    def maximum_frequency(self):
        return np.max([i.maximum_frequency for i in self.interferometers])

    @property
    def number_of_bands(self):
        return len(self.durations)

    # ---- set-up ----------------------------------------------------------------------------------
    def setup_multibanding(self):
        """multiband.py:311-320."""
        self._setup_frequency_bands()
        self._setup_integers()
        self._setup_waveform_frequency_points()
        self._setup_linear_coefficients()
        if self.linear_interpolation:
            self._setup_quadratic_coefficients_linear_interp()
        else:
            self._setup_quadratic_coefficients_ifft_fft()

    def _setup_quadratic_coefficients_ifft_fft(self):
        """multiband.py:613-646.  Also folds everything that is a per-point weight into ``quadratic_coeffs`` (see
        the module docstring): band 0 (multiband.py:771-775) and the even bins of the bands b >= 1."""
        logger.info("IFFT-FFT algorithm is used for (h, h).")
        N = int(self.Nbs[-1])
        Nhatbs = [min(2 * int(Mb), int(Nb)) for Mb, Nb in zip(self.Mbs, self.Nbs)]
        self.Tbhats = [self.interferometers.duration * Nbhat / Nb for Nb, Nbhat in zip(self.Nbs, Nhatbs)]
        self.Ibcs = {ifo.name: [] for ifo in self.interferometers}
        for ifo in self.interferometers:
            full_inv_psds = np.zeros(N // 2 + 1)
            psd, mask = ifo.power_spectral_density_array, ifo.frequency_mask
            full_inv_psds[:len(psd)][mask[:len(full_inv_psds)]] = 1 / psd[mask]
            for b in range(self.number_of_bands):
                Imb = np.fft.irfft(full_inv_psds[:int(self.Nbs[b]) // 2 + 1])
                half_length = Nhatbs[b] // 2
                Imbc = np.append(Imb[:half_length + 1], Imb[-(Nhatbs[b] - half_length - 1):])
                self.Ibcs[ifo.name].append(np.fft.rfft(Imbc))
        self.windows = np.array([])
        self.square_root_windows = np.array([])
        for b in range(self.number_of_bands):
            Ks, Ke = self.Ks_Ke[b]
            ws = self._get_window_sequence(1. / self.durations[b], Ks, Ke - Ks + 1, b)
            self.windows = np.append(self.windows, ws)
            self.square_root_windows = np.append(self.square_root_windows, np.sqrt(ws))
        self._fold_ifft_fft_point_weights(Nhatbs)

    def _fold_ifft_fft_point_weights(self, Nhatbs=None):
        if Nhatbs is None:
            Nhatbs = [min(2 * int(Mb), int(Nb)) for Mb, Nb in zip(self.Mbs, self.Nbs)]
        self.quadratic_coeffs = {}
        for ifo in self.interferometers:
            q = np.zeros(len(self.banded_frequency_points))
            for b in range(self.number_of_bands):
                Ks, Ke = (int(x) for x in self.Ks_Ke[b])
                s0, e0 = (int(x) for x in self.start_end_idxs[b])
                w = self.windows[s0:e0 + 1]
                if b == 0:          # multiband.py:771-775: plain inner product on the full grid
                    q[s0:e0 + 1] = (4. / self.interferometers.duration) * ifo.frequency_mask[Ks:Ke + 1] * w \
                        / ifo.power_spectral_density_array[Ks:Ke + 1]
                elif Nhatbs[b] == 2 * int(self.Mbs[b]):
                    # even bins of the 2 M-point spectrum are the band's points: |sqrt(w) h_k|^2 I[2 k]
                    q[s0:e0 + 1] = (4. / self.Tbhats[b]) * w * self.Ibcs[ifo.name][b].real[2 * np.arange(Ks, Ke + 1)]
                else:               # pragma: no cover - Nb >= 2 Mb for every band b >= 1 (multiband.py:428-447)
                    raise NotImplementedError("IFFT-FFT form with Nhat^(b) != 2 M^(b)")
            self.quadratic_coeffs[ifo.name] = q

    def _tau(self, f):
        """0PN time to merger from frequency f (multiband.py:322-340)."""
        mc = self.reference_chirp_mass_in_second
        return 5 / 256 * mc * (np.pi * mc * (2 * f / self.highest_mode)) ** (-8 / 3)

    def _dtaudf(self, f):
        """multiband.py:342-360."""
        mc = self.reference_chirp_mass_in_second
        return -5 / 96 * mc * (np.pi * mc * (2 * f / self.highest_mode)) ** (-8. / 3.) / f

    def _find_starting_frequency(self, duration, fnow):
        """Lowest start frequency of the next band satisfying (10) and (51) of the paper, by bisection
        (multiband.py:362-400)."""
        def admissible(f):
            smoothing = np.sqrt(-self._dtaudf(f))
            fits = duration - self.time_offset - self._tau(f) - self.accuracy_factor * smoothing > 0
            return fits and f - 1. / smoothing - fnow > 0
        lo, hi = fnow, self.maximum_banding_frequency
        if not admissible(hi):
            return None, None
        f = None
        while hi - lo > 1e-2 / duration:
            f = (lo + hi) / 2.
            if admissible(f):
                hi = f
            else:
                lo = f
        return f, 1. / np.sqrt(-self._dtaudf(f))

    def _setup_frequency_bands(self):
        """multiband.py:402-426: durations T, T/2, T/4, ... and (f^(b), Delta f^(b))."""
        total = self.interferometers.duration
        durations, bands = [total], [[self.minimum_frequency, 0.]]
        nxt = total / 2
        while nxt > max(self.time_offset, self.minimum_banding_duration):
            f, df = self._find_starting_frequency(nxt, bands[-1][0])
            if f is None or not f < min(self.maximum_frequency, self.maximum_banding_frequency):
                break
            durations.append(nxt)
            bands.append([f, df])
            nxt /= 2
        bands.append([self.maximum_frequency + self.delta_f_end, self.delta_f_end])
        self.durations = np.array(durations)
        self.fb_dfb = np.array(bands)
        logger.info("The total frequency range is divided into {} bands with frequency intervals of {}.".format(
            self.number_of_bands, ", ".join(f"1/{d} Hz" for d in self.durations)))

    def _setup_integers(self):
        """multiband.py:428-447: N^(b), M^(b), K^(b)_s, K^(b)_e."""
        total = self.interferometers.duration
        nbs, mbs, ks_ke = [], [], []
        for b, d in enumerate(self.durations):
            (f0, df0), f1 = self.fb_dfb[b], self.fb_dfb[b + 1][0]
            nb = max(_next_power_of_two(2. * (f1 * total + 1.)), 2 ** b)
            nbs.append(nb)
            mbs.append(nb // 2 ** b)
            ks_ke.append([math.ceil((f0 - df0) * d), math.floor(f1 * d)])
        self.Nbs, self.Mbs, self.Ks_Ke = np.array(nbs, dtype=int), np.array(mbs, dtype=int), np.array(ks_ke)

    def _setup_waveform_frequency_points(self):
        """multiband.py:449-478."""
        counts = self.Ks_Ke[:, 1] - self.Ks_Ke[:, 0] + 1
        ends = np.cumsum(counts) - 1
        self.start_end_idxs = np.stack([ends - counts + 1, ends], axis=1)
        self.banded_frequency_points = np.concatenate(
            [np.arange(ks, ke + 1) / d for (ks, ke), d in zip(self.Ks_Ke, self.durations)])
        unique, inverse = np.unique(self.banded_frequency_points, return_inverse=True)
        self.waveform_generator.waveform_arguments["frequencies"] = unique
        self.unique_to_original_frequencies = inverse
        logger.info(f"The number of frequency points where waveforms are evaluated is {len(unique)}.")
        logger.info("The speed-up gain of multi-banding is {}.".format(
            (self.maximum_frequency - self.minimum_frequency) * self.interferometers.duration / len(unique)))

    def _get_window_sequence(self, delta_f, start_idx, length, b):
        """Window of band b at the frequencies (start_idx + i) delta_f (multiband.py:480-527): Hann rise over
        (f^(b) - Delta f^(b), f^(b)), one up to f^(b+1) - Delta f^(b+1), Hann fall to f^(b+1)."""
        (f0, df0), (f1, df1) = self.fb_dfb[b], self.fb_dfb[b + 1]
        idx = np.arange(length) + start_idx
        rise_from = math.floor((f0 - df0) / delta_f) + 1
        one_from = math.ceil(f0 / delta_f)
        fall_from = math.floor((f1 - df1) / delta_f) + 1
        zero_from = math.ceil(f1 / delta_f)
        freqs = idx * delta_f
        window = np.zeros(length)
        window[(idx >= one_from) & (idx < fall_from)] = 1.
        rise = (idx >= rise_from) & (idx < one_from)
        if np.any(rise):                      # df0 = 0 in the first band: no rise
            window[rise] = (1. + np.cos(np.pi * (freqs[rise] - f0) / df0)) / 2.
        fall = (idx >= fall_from) & (idx < zero_from)
        window[fall] = (1. - np.cos(np.pi * (freqs[fall] - f1) / df1)) / 2.
        return window

    def _setup_linear_coefficients(self):
        """multiband.py:529-549: per band, the whitened data is down-sampled (spectrum truncated at N^(b)/2), only its
        last M^(b) time samples are kept, and the transform of that segment at the band's points is windowed."""
        self.linear_coeffs = {}
        n_half = self.Nbs[-1] // 2 + 1
        for ifo in self.interferometers:
            logger.info(f"Pre-computing linear coefficients for {ifo.name}")
            mask = np.asarray(ifo.frequency_mask)
            whitened = np.zeros(len(mask), dtype=complex)
            whitened[mask] = ifo.frequency_domain_strain[mask] / ifo.power_spectral_density_array[mask]
            spectrum = np.zeros(n_half, dtype=complex)
            m = min(n_half, len(whitened))
            spectrum[:m] = whitened[:m]
            pieces = []
            for b, d in enumerate(self.durations):
                ks, ke = self.Ks_Ke[b]
                low = spectrum[:self.Nbs[b] // 2 + 1].copy()
                low[-1] = 0.                                  # Nyquist bin of the down-sampled series
                tail = np.fft.irfft(low)[-self.Mbs[b]:]
                window = self._get_window_sequence(1. / d, ks, ke - ks + 1, b)
                pieces.append((4. / d) * window * np.conj(np.fft.rfft(tail)[ks:ke + 1]))
            self.linear_coeffs[ifo.name] = np.concatenate(pieces)

    def _setup_quadratic_coefficients_linear_interp(self):
        """multiband.py:551-611: |h|^2 is interpolated linearly between neighbouring banded points and summed against
        window / PSD on the original grid; every grid bin is assigned to its interval by one searchsorted and the
        two interpolation weights are scattered with bincount."""
        logger.info("Linear-interpolation algorithm is used for (h, h).")
        total = float(self.interferometers.duration)
        pieces = {ifo.name: [] for ifo in self.interferometers}
        for b, d in enumerate(self.durations):
            s, e = self.start_end_idxs[b]
            points = self.banded_frequency_points[s:e + 1]
            n_pts = len(points)
            (f0, df0), f1 = self.fb_dfb[b], self.fb_dfb[b + 1][0]
            first = math.ceil((f0 - df0) * total)
            window = self._get_window_sequence(1 / total, first, math.floor(f1 * total) - first + 1, b)
            lower = np.array([math.ceil(total * f) for f in points])       # first grid bin of each interval
            for ifo in self.interferometers:
                psd = np.asarray(ifo.power_spectral_density_array)
                last = min(first + len(window) - 1, len(psd) - 1)
                bins = np.arange(first, last + 1)
                weight = np.zeros(len(bins))
                ok = np.asarray(ifo.frequency_mask[first:last + 1])
                weight[ok] = 1. / psd[first:last + 1][ok]
                weight *= window[:len(bins)]
                k = np.clip(np.searchsorted(lower, bins, side="right") - 1, 0, n_pts - 2)
                freqs = bins / total
                scale = 4 * d / total
                coeffs = np.bincount(k, weights=(points[k + 1] - freqs) * weight, minlength=n_pts)
                coeffs += np.bincount(k + 1, weights=(freqs - points[k]) * weight, minlength=n_pts)
                pieces[ifo.name].append(scale * coeffs)
        self.quadratic_coeffs = {name: np.concatenate(v) for name, v in pieces.items()}

    _WEIGHT_KEYS = ("reference_chirp_mass", "highest_mode", "linear_interpolation", "accuracy_factor", "time_offset",
                    "delta_f_end", "maximum_banding_frequency", "minimum_banding_duration", "durations", "fb_dfb",
                    "Nbs", "Mbs", "Ks_Ke", "banded_frequency_points", "start_end_idxs",
                    "unique_to_original_frequencies", "linear_coeffs")

    @property
    def weights(self):
        """multiband.py:647-671."""
        out = {key: getattr(self, key) for key in self._WEIGHT_KEYS}
        out["waveform_frequencies"] = self.waveform_generator.waveform_arguments["frequencies"]
        out["quadratic_coeffs"] = self.quadratic_coeffs
        if not self.linear_interpolation:
            for key in ("Tbhats", "Ibcs", "windows", "square_root_windows"):
                out[key] = getattr(self, key)
        return out

    def save_weights(self, filename):
        raise NotImplementedError("multiband weights files are HDF5 (h5py is absent): keep the `weights` dict")

    def setup_multibanding_from_weights(self, weights):
        """multiband.py:689-712 (dict form)."""
        self.reference_chirp_mass = weights["reference_chirp_mass"]
        for key, value in weights.items():
            if key == "reference_chirp_mass":
                continue
            if key == "waveform_frequencies":
                self.waveform_generator.waveform_arguments["frequencies"] = value
            else:
                setattr(self, key, value)

    def _setup_time_marginalization_multiband(self):
        """multiband.py:714-726."""
        N = int(self.Nbs[-1]) // 2
        self._delta_tc = self.durations[0] / N
        self._times = self.interferometers.start_time + np.arange(N) * self._delta_tc
        self.time_prior_array = self.priors["geocent_time"].prob(self._times) * self._delta_tc
        self._full_to_multiband = [int(f * self.durations[0]) for f in self.banded_frequency_points]
        self._beam_pattern_reference_time = (self.priors["geocent_time"].minimum
                                             + self.priors["geocent_time"].maximum) / 2
        for ifo in self.interferometers:
            ifo.reference_time = self._beam_pattern_reference_time

    # ---- device ----------------------------------------------------------------------------------
    def _pack_host_arrays(self):
        n_det, n = len(self.interferometers), len(self.banded_frequency_points)
        lin = np.zeros((n_det, n, 2))
        quad = np.zeros((n_det, n))
        for d, ifo in enumerate(self.interferometers):
            lin[d, :, 0], lin[d, :, 1] = self.linear_coeffs[ifo.name].real, self.linear_coeffs[ifo.name].imag
            quad[d] = self.quadratic_coeffs[ifo.name]
        return dict(freqs=np.ascontiguousarray(self.banded_frequency_points, dtype=np.float64), linear=lin,
                    quadratic=quad)

    def _upload_multiband(self):
        net = self.device_network
        hst = self._mb_host
        _lib.check(net.lib.bb_set_multiband(net.ptr, len(hst["freqs"]), hst["freqs"].ctypes.data,
                                            hst["linear"].ctypes.data, hst["quadratic"].ctypes.data))
        if not self.linear_interpolation and self.number_of_bands > 1:
            bands = range(1, self.number_of_bands)
            ints = lambda vals: np.ascontiguousarray(list(vals), dtype=np.int32)      # noqa: E731
            m = ints(int(self.Mbs[b]) for b in bands)
            ks = ints(int(self.Ks_Ke[b][0]) for b in bands)
            ke = ints(int(self.Ks_Ke[b][1]) for b in bands)
            st = ints(int(self.start_end_idxs[b][0]) for b in bands)
            norm = np.ascontiguousarray([4. / self.Tbhats[b] for b in bands], dtype=np.float64)
            sw = np.ascontiguousarray(self.square_root_windows, dtype=np.float64)
            i_odd = np.ascontiguousarray(np.concatenate([
                np.concatenate([self.Ibcs[ifo.name][b].real[1::2][:int(self.Mbs[b]) // 2] for ifo in self.interferometers])
                for b in bands]), dtype=np.float64)
            _lib.check(net.lib.bb_set_multiband_ifft_fft(net.ptr, len(m), m.ctypes.data, ks.ctypes.data, ke.ctypes.data,
                                                         st.ctypes.data, norm.ctypes.data, sw.ctypes.data,
                                                         i_odd.ctypes.data))
        if self.time_marginalization:
            idx = np.ascontiguousarray(self._full_to_multiband, dtype=np.int32)
            _lib.check(net.lib.bb_set_multiband_time_marginalization(
                net.ptr, int(self.Nbs[-1]) // 2, idx.ctypes.data, float(self._delta_tc),
                float(self._beam_pattern_reference_time)))

    def _configure(self):
        super()._configure()
        if self._mb_host is not None:
            self._upload_multiband()

    # evaluation: the base class entry points (log_likelihood_ratio, log_likelihood_ratio_batch,
    # inner_products_batch, calculate_snrs) run K5 once the handle carries the multi-band tables.
