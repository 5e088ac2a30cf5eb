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


lal_binary_black_hole._bb_defaults = _DEFAULTS_BBH
lal_binary_neutron_star._bb_defaults = _DEFAULTS_BNS
