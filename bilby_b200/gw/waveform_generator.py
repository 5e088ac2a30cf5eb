"""WaveformGenerator with the reference's constructor and call semantics
(bilby/gw/waveform_generator.py:24-110 constructor, :113-141 frequency_domain_strain,
:178-209 _calculate_strain + 1-entry cache, :260-269 _format_parameters,
:271-286 _parameters_from_source_model).

Difference that matters: the source model is a device source model (bilby_b200.gw.source); on the
likelihood path polarisations are never materialised - the generator only hands the converted
parameter rows and the approximant to the fused kernel.  ``frequency_domain_strain`` still returns
{"plus", "cross"} numpy arrays for injections and tests (computed by bb_frequency_domain_strain_device).
"""
import ctypes
import inspect

import numpy as np

from ..core.utils import create_frequency_series, create_time_series, logger
from . import _params
from .conversion import convert_to_lal_binary_black_hole_parameters


def infer_parameters_from_function(func):
    """bilby/core/utils/introspection.py:5-38: named arguments except the first, no *args/**kwargs."""
    sig = inspect.signature(func)
    names = [name for name, p in sig.parameters.items()
             if p.kind not in (p.VAR_POSITIONAL, p.VAR_KEYWORD)]
    return names[1:]


class WaveformGenerator:
    def __init__(self, duration=None, sampling_frequency=None, start_time=0, frequency_domain_source_model=None,
                 time_domain_source_model=None, parameters=None, parameter_conversion=None,
                 waveform_arguments=None, use_cache=True):
        if time_domain_source_model is not None:
            raise NotImplementedError("time-domain source models are outside the hot path (SURVEY.md section 8)")
        self.duration = duration
        self.sampling_frequency = sampling_frequency
        self.start_time = start_time
        self.frequency_domain_source_model = frequency_domain_source_model
        self.time_domain_source_model = None
        self.source_parameter_keys = self._parameters_from_source_model()
        self.parameter_conversion = (convert_to_lal_binary_black_hole_parameters
                                     if parameter_conversion is None else parameter_conversion)
        self.waveform_arguments = dict(waveform_arguments) if waveform_arguments is not None else dict()
        self._cache = dict(parameters=None, waveform=None, model=None)
        self.use_cache = use_cache
        self._handle = None
        self._handle_key = None

    def __repr__(self):
        return (f"{self.__class__.__name__}(duration={self.duration}, sampling_frequency={self.sampling_frequency}, "
                f"start_time={self.start_time}, frequency_domain_source_model="
                f"{getattr(self.frequency_domain_source_model, '__name__', None)}, "
                f"waveform_arguments={self.waveform_arguments})")

    # ---- grids (bilby/core/series.py CoupledTimeAndFrequencySeries)
    @property
    def frequency_array(self):
        return create_frequency_series(self.sampling_frequency, self.duration)

    @property
    def time_array(self):
        return create_time_series(self.sampling_frequency, self.duration, self.start_time)

    def _parameters_from_source_model(self):
        if self.frequency_domain_source_model is None:
            raise AttributeError("Either time or frequency domain source model must be provided.")
        return set(infer_parameters_from_function(self.frequency_domain_source_model))

    # ---- device-facing description of the source model
    @property
    def full_waveform_arguments(self):
        defaults = getattr(self.frequency_domain_source_model, "_bb_defaults", None)
        if defaults is None:
            raise TypeError(
                "bilby_b200 evaluates waveforms on the device: frequency_domain_source_model must be one of "
                "bilby_b200.gw.source.lal_binary_black_hole / lal_binary_neutron_star")
        wa = dict(defaults)
        wa.update(self.waveform_arguments)
        return wa

    def approximant_config(self):
        """(approximant id, f_ref, f_min, f_max) for bb_set_waveform; f_max <= 0 means 'last bin'."""
        wa = self.full_waveform_arguments
        name = wa["waveform_approximant"]
        if name not in _params.APPROXIMANTS:
            raise ValueError(f"waveform_approximant '{name}' has no device kernel "
                             f"(available: {sorted(_params.APPROXIMANTS)})")
        known = {"waveform_approximant", "reference_frequency", "minimum_frequency", "maximum_frequency",
                 "catch_waveform_errors", "pn_spin_order", "pn_tidal_order", "pn_phase_order",
                 "pn_amplitude_order", "mode_array"}
        kind = getattr(self.frequency_domain_source_model, "_bb_kind", "grid")
        if kind == "relative_binning":       # source.py:724-799
            known |= {"fiducial", "frequency_bin_edges"}
        elif kind == "roq":                  # source.py:802-898
            known |= {"frequency_nodes", "linear_indices", "quadratic_indices", "frequency_nodes_linear",
                      "frequency_nodes_quadratic"}
        elif kind == "frequency_sequence":   # source.py:901-1140
            known |= {"frequencies"}
        unused = set(wa) - known
        if unused:
            raise ValueError(f"There are unused waveform kwargs: {sorted(unused)}")   # source.py:687-688
        for key in ("pn_spin_order", "pn_tidal_order", "pn_phase_order"):
            if wa.get(key, -1) != -1:
                raise NotImplementedError(f"{key} != -1 is not supported by the device kernels")
        if wa.get("pn_amplitude_order", 0) != 0:
            raise NotImplementedError("pn_amplitude_order != 0 is not supported by the device kernels")
        return (_params.APPROXIMANTS[name], float(wa["reference_frequency"]), float(wa.get("minimum_frequency", 20.0)),
                float(wa.get("maximum_frequency", 0.0) or 0.0))

    @property
    def catch_waveform_errors(self):
        return bool(self.full_waveform_arguments.get("catch_waveform_errors", False))

    def _format_parameters(self, parameters):
        """waveform_generator.py:260-269 (without the update by waveform_arguments: those go to the device
        through bb_set_waveform)."""
        if not isinstance(parameters, dict):
            raise TypeError('"parameters" must be a dictionary.')
        new_parameters = parameters.copy()
        new_parameters, _ = self.parameter_conversion(new_parameters)
        for key in self.source_parameter_keys.symmetric_difference(new_parameters):
            new_parameters.pop(key)
        return new_parameters

    def convert(self, parameters):
        """Converted parameters keeping the extrinsic keys the detector projection needs."""
        converted, _ = self.parameter_conversion(dict(parameters))
        missing = [k for k in self.source_parameter_keys if k not in converted]
        if missing:
            raise KeyError(missing[0])
        return converted

    # ---- polarisations for injections / tests
    def _get_handle(self):
        import torch
        from .. import _lib
        key = (self.duration, self.sampling_frequency, torch.cuda.current_device() if torch.cuda.is_available() else -1)
        if self._handle is None or self._handle_key != key:
            h = _lib.Handle()
            n_freq = len(self.frequency_array)
            tens = np.zeros(9)
            vert = np.zeros(3)
            strain = np.zeros((1, n_freq, 2))
            psd = np.ones((1, n_freq))
            mask = np.ones((1, n_freq), dtype=np.uint8)
            _lib.check(h.lib.bb_set_network(h.ptr, 1, n_freq, float(self.duration), float(self.sampling_frequency),
                                            0.0, tens.ctypes.data, vert.ctypes.data, strain.ctypes.data,
                                            psd.ctypes.data, mask.ctypes.data))
            self._handle, self._handle_key = h, key
        return self._handle

    def frequency_domain_strain(self, parameters=None):
        """waveform_generator.py:113-141.  Returns {"plus","cross"} complex128 numpy arrays, or None when
        the waveform is outside its domain and catch_waveform_errors is set (source.py:644-662)."""
        import torch
        from .. import _lib
        if parameters is None:
            parameters = self._cache.get("parameters", None)
        if parameters is None:
            raise ValueError("No parameters given to generate waveform.")
        if self.use_cache and parameters == self._cache.get("parameters", None) \
                and self._cache["model"] == self.frequency_domain_source_model:
            return self._cache["waveform"]
        self._cache["parameters"] = parameters.copy()
        self._cache["model"] = self.frequency_domain_source_model
        src = self._format_parameters(parameters)
        if getattr(self.frequency_domain_source_model, "_bb_kind", "grid") == "frequency_sequence":
            # source.py:901-1140: the polarisations at waveform_arguments['frequencies'], not on the grid
            result = self.frequency_sequence_strain(parameters, self.waveform_arguments["frequencies"])
            self._cache["waveform"] = result
            return result
        h = self._get_handle()
        approx, f_ref, f_min, f_max = self.approximant_config()
        _lib.check(h.lib.bb_set_waveform(h.ptr, approx, f_ref, f_min, f_max))
        rows = _params.pack_rows(src, 1, np)
        n_freq = len(self.frequency_array)
        dev = torch.device("cuda", h.device)
        rows_d = torch.from_numpy(rows).to(dev)
        out = torch.empty((1, 2, n_freq, 2), dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(h.lib.bb_frequency_domain_strain_device(h.ptr, rows_d.data_ptr(), 1, out.data_ptr(),
                                                           ctypes.c_void_p(stream)))
        arr = out.cpu().numpy()
        bad = self._domain_error(src, f_min)
        if bad:
            if self.catch_waveform_errors:
                self._cache["waveform"] = None
                return None
            raise RuntimeError("Internal function call failed: Input domain error")
        result = dict(plus=arr[0, 0, :, 0] + 1j * arr[0, 0, :, 1], cross=arr[0, 1, :, 0] + 1j * arr[0, 1, :, 1])
        self._cache["waveform"] = result
        return result

    def frequency_sequence_strain(self, parameters, frequencies):
        """{"plus","cross"} at arbitrary frequencies (source.py:1068-1140 _base_waveform_frequency_sequence)."""
        import torch
        from .. import _lib
        src = self._format_parameters(parameters)
        h = self._get_handle()
        approx, f_ref, f_min, f_max = self.approximant_config()
        _lib.check(h.lib.bb_set_waveform(h.ptr, approx, f_ref, f_min, f_max))
        rows = _params.pack_rows(src, 1, np)
        dev = torch.device("cuda", h.device)
        rows_d = torch.from_numpy(rows).to(dev)
        fr = torch.from_numpy(np.ascontiguousarray(frequencies, dtype=np.float64)).to(dev)
        nn = fr.shape[0]
        out = torch.empty((1, 2, nn, 2), dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(h.lib.bb_frequency_sequence_strain_device(h.ptr, rows_d.data_ptr(), 1, fr.data_ptr(), nn,
                                                             float(frequencies[0]), out.data_ptr(),
                                                             ctypes.c_void_p(stream)))
        arr = out.cpu().numpy()
        if self._domain_error(src, f_min):
            if self.catch_waveform_errors:
                return None
            raise RuntimeError("Internal function call failed: Input domain error")
        return dict(plus=arr[0, 0, :, 0] + 1j * arr[0, 0, :, 1], cross=arr[0, 1, :, 0] + 1j * arr[0, 1, :, 1])

    @staticmethod
    def _domain_error(src, f_min):
        """Host mirror of the prologue's status flag for the scalar API (error convention only)."""
        m1, m2 = float(src["mass_1"]), float(src["mass_2"])
        chi1 = float(src.get("a_1", 0.0)) * np.cos(float(src.get("tilt_1", 0.0)))
        chi2 = float(src.get("a_2", 0.0)) * np.cos(float(src.get("tilt_2", 0.0)))
        if not (m1 > 0 and m2 > 0 and float(src["luminosity_distance"]) > 0):
            return True
        return abs(chi1) > 1 or abs(chi2) > 1

    @property
    def meta_data(self):
        return dict(frequency_domain_source_model=getattr(self.frequency_domain_source_model, "__name__", None),
                    waveform_arguments=self.waveform_arguments, start_time=self.start_time,
                    sampling_frequency=self.sampling_frequency, duration=self.duration)
