"""Source models with the reference's named-parameter signatures (bilby/gw/source.py:269-348
``lal_binary_black_hole``, :351-432 ``lal_binary_neutron_star``).

In the reference these call lalsimulation on the CPU and return {"plus", "cross"} arrays.  Here
they are *device source models*: WaveformGenerator introspects the signature exactly like the
reference (waveform_generator.py:271-286) and the approximant named in ``waveform_arguments``
selects the CUDA kernel; polarisations are only materialised when a caller asks for them
(injection, tests) through ``bb_frequency_domain_strain_device``.
"""
import numpy as np

_DEFAULTS_BBH = dict(waveform_approximant="IMRPhenomPv2", reference_frequency=50.0, minimum_frequency=20.0,
                     catch_waveform_errors=False, pn_spin_order=-1, pn_tidal_order=-1, pn_phase_order=-1,
                     pn_amplitude_order=0)
_DEFAULTS_BNS = dict(_DEFAULTS_BBH, waveform_approximant="IMRPhenomPv2_NRTidal")

SUPPORTED_APPROXIMANTS = {"lal_binary_black_hole": ("IMRPhenomD",),
                          "lal_binary_neutron_star": ("TaylorF2", "IMRPhenomD")}


def _evaluate(model, frequency_array, params, kwargs, defaults):
    from .waveform_generator import WaveformGenerator
    wa = dict(defaults)
    wa.update(kwargs)
    n = len(frequency_array)
    fs = 2 * float(frequency_array[-1])
    duration = (n - 1) / float(frequency_array[-1])
    wfg = WaveformGenerator(duration=duration, sampling_frequency=fs, frequency_domain_source_model=model,
                            waveform_arguments=wa, parameter_conversion=lambda p: (p, []))
    return wfg.frequency_domain_strain(params)


def lal_binary_black_hole(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2,
                          tilt_2, phi_jl, theta_jn, phase, **kwargs):
    """source.py:269-348.  Supported approximant on the device: IMRPhenomD (aligned spins)."""
    params = dict(mass_1=mass_1, mass_2=mass_2, luminosity_distance=luminosity_distance, a_1=a_1, tilt_1=tilt_1,
                  phi_12=phi_12, a_2=a_2, tilt_2=tilt_2, phi_jl=phi_jl, theta_jn=theta_jn, phase=phase)
    return _evaluate(lal_binary_black_hole, np.asarray(frequency_array), params, kwargs, _DEFAULTS_BBH)


def lal_binary_neutron_star(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2,
                            tilt_2, phi_jl, theta_jn, phase, lambda_1, lambda_2, **kwargs):
    """source.py:351-432.  Supported approximant on the device: TaylorF2 (+tides)."""
    params = dict(mass_1=mass_1, mass_2=mass_2, luminosity_distance=luminosity_distance, a_1=a_1, tilt_1=tilt_1,
                  phi_12=phi_12, a_2=a_2, tilt_2=tilt_2, phi_jl=phi_jl, theta_jn=theta_jn, phase=phase,
                  lambda_1=lambda_1, lambda_2=lambda_2)
    return _evaluate(lal_binary_neutron_star, np.asarray(frequency_array), params, kwargs, _DEFAULTS_BNS)


def _sequence(model, frequencies, params, kwargs, defaults):
    """Polarisations at arbitrary frequencies (source.py:1068-1140) through bb_frequency_sequence_strain_device."""
    from .waveform_generator import WaveformGenerator
    wa = dict(defaults)
    wa.update(kwargs)
    for key in ("minimum_frequency", "maximum_frequency"):
        wa.pop(key, None)
    wfg = WaveformGenerator(duration=4.0, sampling_frequency=2048.0, frequency_domain_source_model=model,
                            waveform_arguments=wa, parameter_conversion=lambda p: (p, []))
    return wfg.frequency_sequence_strain(params, np.asarray(frequencies, dtype=float))


def _bbh_params(mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2, tilt_2, phi_jl, theta_jn, phase):
    return dict(mass_1=mass_1, mass_2=mass_2, luminosity_distance=luminosity_distance, a_1=a_1, tilt_1=tilt_1,
                phi_12=phi_12, a_2=a_2, tilt_2=tilt_2, phi_jl=phi_jl, theta_jn=theta_jn, phase=phase)


def _relative_binning(model, grid_model, frequency_array, params, kwargs, defaults):
    """source.py:724-799: fiducial = 1 -> waveform on the full grid, fiducial = 0 -> at frequency_bin_edges."""
    kwargs = dict(kwargs)
    fiducial = kwargs.pop("fiducial", 0)
    if fiducial == 1:
        kwargs.pop("frequency_bin_edges", None)
        return _evaluate(grid_model, np.asarray(frequency_array), params, kwargs, defaults)
    edges = kwargs.pop("frequency_bin_edges")
    return _sequence(model, edges, params, kwargs, defaults)


def lal_binary_black_hole_relative_binning(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12,
                                           a_2, tilt_2, phi_jl, theta_jn, phase, **kwargs):
    """source.py:724-761."""
    params = _bbh_params(mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2, tilt_2, phi_jl, theta_jn, phase)
    return _relative_binning(lal_binary_black_hole_relative_binning, lal_binary_black_hole, frequency_array, params,
                             kwargs, _DEFAULTS_BBH)


def lal_binary_neutron_star_relative_binning(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1,
                                             phi_12, a_2, tilt_2, phi_jl, lambda_1, lambda_2, theta_jn, phase,
                                             **kwargs):
    """source.py:764-799."""
    params = _bbh_params(mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2, tilt_2, phi_jl, theta_jn, phase)
    params.update(lambda_1=lambda_1, lambda_2=lambda_2)
    return _relative_binning(lal_binary_neutron_star_relative_binning, lal_binary_neutron_star, frequency_array,
                             params, kwargs, _DEFAULTS_BNS)


def _roq(model, params, kwargs, defaults):
    """source.py:802-898 _base_roq_waveform: waveform at the unique nodes, gathered into linear / quadratic order."""
    wa = dict(kwargs)
    if "frequency_nodes" not in wa:
        size_linear = len(wa["frequency_nodes_linear"])
        combined = np.hstack((wa.pop("frequency_nodes_linear"), wa.pop("frequency_nodes_quadratic")))
        nodes, inverse = np.unique(combined, return_inverse=True)
        linear_indices, quadratic_indices = inverse[:size_linear], inverse[size_linear:]
    else:
        linear_indices = wa.pop("linear_indices")
        quadratic_indices = wa.pop("quadratic_indices")
        for key in ("frequency_nodes_linear", "frequency_nodes_quadratic"):
            wa.pop(key, None)
        nodes = wa.pop("frequency_nodes")
    pols = _sequence(model, nodes, params, wa, defaults)
    if pols is None:
        return None
    return dict(linear=dict(plus=pols["plus"][linear_indices], cross=pols["cross"][linear_indices]),
                quadratic=dict(plus=pols["plus"][quadratic_indices], cross=pols["cross"][quadratic_indices]))


def binary_black_hole_roq(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2, tilt_2,
                          phi_jl, theta_jn, phase, **waveform_arguments):
    """source.py:693-706 (reference_frequency defaults to 20 Hz)."""
    params = _bbh_params(mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2, tilt_2, phi_jl, theta_jn, phase)
    return _roq(binary_black_hole_roq, params, waveform_arguments, _DEFAULTS_BBH_ROQ)


def binary_neutron_star_roq(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2, tilt_2,
                            phi_jl, lambda_1, lambda_2, theta_jn, phase, **waveform_arguments):
    """source.py:709-721."""
    params = _bbh_params(mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2, tilt_2, phi_jl, theta_jn, phase)
    params.update(lambda_1=lambda_1, lambda_2=lambda_2)
    return _roq(binary_neutron_star_roq, params, waveform_arguments, _DEFAULTS_BNS_ROQ)


def binary_black_hole_frequency_sequence(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12,
                                         a_2, tilt_2, phi_jl, theta_jn, phase, **waveform_kwargs):
    """source.py:901-979: the waveform at waveform_kwargs['frequencies'] (the multi-banded likelihood's points)."""
    params = _bbh_params(mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2, tilt_2, phi_jl, theta_jn, phase)
    wa = dict(waveform_kwargs)
    return _sequence(binary_black_hole_frequency_sequence, wa.pop("frequencies"), params, wa, _DEFAULTS_BBH_SEQ)


def binary_neutron_star_frequency_sequence(frequency_array, mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12,
                                           a_2, tilt_2, phi_jl, lambda_1, lambda_2, theta_jn, phase,
                                           **waveform_kwargs):
    """source.py:982-1065."""
    params = _bbh_params(mass_1, mass_2, luminosity_distance, a_1, tilt_1, phi_12, a_2, tilt_2, phi_jl, theta_jn, phase)
    params.update(lambda_1=lambda_1, lambda_2=lambda_2)
    wa = dict(waveform_kwargs)
    return _sequence(binary_neutron_star_frequency_sequence, wa.pop("frequencies"), params, wa, _DEFAULTS_BNS_SEQ)


_DEFAULTS_BBH_SEQ = dict(waveform_approximant="IMRPhenomPv2", reference_frequency=50.0, catch_waveform_errors=False,
                         pn_spin_order=-1, pn_tidal_order=-1, pn_phase_order=-1, pn_amplitude_order=0)
_DEFAULTS_BNS_SEQ = dict(_DEFAULTS_BBH_SEQ, waveform_approximant="IMRPhenomPv2_NRTidal")
_DEFAULTS_BBH_ROQ = dict(waveform_approximant="IMRPhenomPv2", reference_frequency=20.0, catch_waveform_errors=False,
                         pn_spin_order=-1, pn_tidal_order=-1, pn_phase_order=-1, pn_amplitude_order=0)
_DEFAULTS_BNS_ROQ = dict(_DEFAULTS_BBH_ROQ, waveform_approximant="IMRPhenomD_NRTidal")

lal_binary_black_hole._bb_defaults = _DEFAULTS_BBH
lal_binary_neutron_star._bb_defaults = _DEFAULTS_BNS
lal_binary_black_hole_relative_binning._bb_defaults = _DEFAULTS_BBH
lal_binary_neutron_star_relative_binning._bb_defaults = _DEFAULTS_BNS
binary_black_hole_roq._bb_defaults = _DEFAULTS_BBH_ROQ
binary_neutron_star_roq._bb_defaults = _DEFAULTS_BNS_ROQ
binary_black_hole_frequency_sequence._bb_defaults = _DEFAULTS_BBH_SEQ
binary_neutron_star_frequency_sequence._bb_defaults = _DEFAULTS_BNS_SEQ
for _f, _k in ((lal_binary_black_hole, "grid"), (lal_binary_neutron_star, "grid"),
               (lal_binary_black_hole_relative_binning, "relative_binning"),
               (lal_binary_neutron_star_relative_binning, "relative_binning"),
               (binary_black_hole_roq, "roq"), (binary_neutron_star_roq, "roq"),
               (binary_black_hole_frequency_sequence, "frequency_sequence"),
               (binary_neutron_star_frequency_sequence, "frequency_sequence")):
    _f._bb_kind = _k
