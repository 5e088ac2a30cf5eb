"""BUILD TOOL - golden vectors for the multi-banded likelihood from the UNMODIFIED
reference class MBGravitationalWaveTransient (bilby imported from /root/reference with the stand-ins), fed by the
restated frequency-sequence source models of oracle/cbc_multiband.py (lalsimulation is absent).

    PYTHONPATH=oracle/standins:/root/reference python oracle/tools/make_golden_multiband.py

Writes tests/golden/multiband_bbh_8s_H1L1V1.npz and multiband_bns_32s_H1L1V1.npz.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import bilby  # noqa: E402  (the reference)
from bilby.core.prior import Uniform, PowerLaw, PriorDict  # noqa: E402
from oracle import cbc_likelihood as ocl  # noqa: E402
from oracle import cbc_multiband as ocm  # noqa: E402
from make_golden_reduced import make_ifos, near, evaluate, BNS_INJ, T_INJ, NOISE_SEED  # noqa: E402

bilby.core.utils.logger.setLevel("ERROR")


def multiband(tag, duration, fs, inj, model, grid_model, conversion, approximant, bns, ref_mc):
    names = ["H1", "L1", "V1"]
    start_time = T_INJ - duration + 2
    wa = dict(waveform_approximant=approximant, reference_frequency=50.0)
    ifos = make_ifos(duration, fs, names, start_time)
    wfg_full = bilby.gw.WaveformGenerator(
        duration=duration, sampling_frequency=fs, start_time=start_time, frequency_domain_source_model=grid_model,
        parameter_conversion=conversion, waveform_arguments=dict(wa, minimum_frequency=20.0))
    pols = wfg_full.frequency_domain_strain(dict(inj))
    for ifo in ifos:
        ifo.inject_signal_from_waveform_polarizations(parameters=dict(inj), injection_polarizations=pols)

    def wfg_new():
        return bilby.gw.WaveformGenerator(duration=duration, sampling_frequency=fs, start_time=start_time,
                                          frequency_domain_source_model=model, parameter_conversion=conversion,
                                          waveform_arguments=dict(wa))
    n = 24
    draws = near(inj, n, np.random.default_rng(20261017), bns=bns)
    tprior = Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time")
    res = dict(start_time=start_time, duration=duration, sampling_frequency=fs, detectors=np.array(names),
               noise_seed=NOISE_SEED, approximant=approximant, reference_chirp_mass=ref_mc,
               geocent_time_prior=np.array([tprior.minimum, tprior.maximum]))
    for k in draws:
        res["param_" + k] = draws[k]
    for k, v in inj.items():
        res["inj_" + k] = v
    like = bilby.gw.likelihood.MBGravitationalWaveTransient(
        ifos, wfg_new(), reference_chirp_mass=ref_mc, priors=PriorDict(dict(geocent_time=tprior)))
    for key in ("durations", "fb_dfb", "Nbs", "Mbs", "Ks_Ke", "banded_frequency_points", "start_end_idxs",
                "unique_to_original_frequencies"):
        res[key] = np.asarray(getattr(like, key))
    res["time_offset"], res["delta_f_end"] = like.time_offset, like.delta_f_end
    res["maximum_banding_frequency"] = like.maximum_banding_frequency
    for ifo in ifos:
        res[f"linear_coeffs_{ifo.name}"] = like.linear_coeffs[ifo.name]
        res[f"quadratic_coeffs_{ifo.name}"] = like.quadratic_coeffs[ifo.name]
    res["lnl_none"] = evaluate(like, draws, n)
    dh = np.zeros((n, 3), dtype=complex)
    hh = np.zeros((n, 3))
    for i in range(n):
        p = {k: float(v[i]) for k, v in draws.items()}
        p.update(like.get_sky_frame_parameters(p))
        pols_i = like.waveform_generator.frequency_domain_strain(p)
        for j, ifo in enumerate(ifos):
            snr = like.calculate_snrs(pols_i, ifo, parameters=p)
            dh[i, j], hh[i, j] = snr.d_inner_h, snr.optimal_snr_squared
    res["d_inner_h"], res["optimal_snr_squared"] = dh, hh
    # the full-grid likelihood of the same draws (multi-banding is an approximation of it)
    full = bilby.gw.likelihood.GravitationalWaveTransient(ifos, wfg_full)
    res["lnl_full_grid"] = evaluate(full, draws, n)
    pri = PriorDict(dict(geocent_time=tprior, phase=Uniform(0, 2 * np.pi, "phase")))
    like = bilby.gw.likelihood.MBGravitationalWaveTransient(
        ifos, wfg_new(), reference_chirp_mass=ref_mc, priors=pri, phase_marginalization=True)
    res["lnl_phase"] = evaluate(like, draws, n)
    dmin, dmax = (10.0, 500.0) if bns else (100.0, 5000.0)
    pri = PriorDict(dict(geocent_time=tprior, phase=Uniform(0, 2 * np.pi, "phase"),
                         luminosity_distance=PowerLaw(2, dmin, dmax, "luminosity_distance")))
    like = bilby.gw.likelihood.MBGravitationalWaveTransient(
        ifos, wfg_new(), reference_chirp_mass=ref_mc, priors=pri, phase_marginalization=True,
        distance_marginalization=True, distance_marginalization_lookup_table=f"/tmp/golden_mb_{tag}_lookup.npz")
    res["lnl_distance_phase"] = evaluate(like, draws, n)
    res["distance_prior"] = np.array([dmin, dmax])
    # the IFFT-FFT form of (h, h) (multiband.py:613-646, 766-787)
    like = bilby.gw.likelihood.MBGravitationalWaveTransient(
        ifos, wfg_new(), reference_chirp_mass=ref_mc, priors=PriorDict(dict(geocent_time=tprior)),
        linear_interpolation=False)
    res["lnl_ifft_fft"] = evaluate(like, draws, n)
    hh2 = np.zeros((n, 3))
    for i in range(n):
        p = {k: float(v[i]) for k, v in draws.items()}
        p.update(like.get_sky_frame_parameters(p))
        pols_i = like.waveform_generator.frequency_domain_strain(p)
        for j, ifo in enumerate(ifos):
            hh2[i, j] = like.calculate_snrs(pols_i, ifo, parameters=p).optimal_snr_squared
    res["optimal_snr_squared_ifft_fft"] = hh2
    # time marginalisation (multiband.py:714-726, 789-797): geocent_time := start_time, jitter drawn inside its prior
    tdraws = dict(draws)
    tdraws["geocent_time"] = np.full(n, float(start_time))
    pri = PriorDict(dict(geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time")))
    like = bilby.gw.likelihood.MBGravitationalWaveTransient(
        ifos, wfg_new(), reference_chirp_mass=ref_mc, priors=pri, time_marginalization=True, jitter_time=True)
    jmax = float(pri["time_jitter"].maximum)
    tdraws["time_jitter"] = np.random.default_rng(5).uniform(-jmax, jmax, n)
    res["param_time_jitter"] = tdraws["time_jitter"]
    res["time_marg_delta_tc"] = like._delta_tc
    res["lnl_time"] = evaluate(like, tdraws, n)
    pri = PriorDict(dict(geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time"), phase=Uniform(0, 2 * np.pi, "phase")))
    like = bilby.gw.likelihood.MBGravitationalWaveTransient(
        ifos, wfg_new(), reference_chirp_mass=ref_mc, priors=pri, time_marginalization=True, jitter_time=True,
        phase_marginalization=True)
    res["lnl_time_phase"] = evaluate(like, tdraws, n)
    for ifo in ifos:
        ifo.reference_time = None
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"multiband_{tag}.npz"), **res)
    print(tag, "bands", len(res["durations"]), "points", len(res["banded_frequency_points"]), "lnl",
          res["lnl_none"][:3], res["lnl_full_grid"][:3], res["lnl_distance_phase"][:3])


if __name__ == "__main__":
    what = sys.argv[1:] or ["bbh", "bns"]
    if "bbh" in what:
        inj = dict(ocl.INJECTION)
        mc = (inj["mass_1"] * inj["mass_2"]) ** 0.6 / (inj["mass_1"] + inj["mass_2"]) ** 0.2
        multiband("bbh_8s_H1L1V1", 8.0, 2048.0, inj, ocm.binary_black_hole_frequency_sequence,
                  ocl.lal_binary_black_hole, bilby.gw.conversion.convert_to_lal_binary_black_hole_parameters,
                  "IMRPhenomD", False, 0.9 * mc)
    if "bns" in what:
        multiband("bns_32s_H1L1V1", 32.0, 4096.0, dict(BNS_INJ), ocm.binary_neutron_star_frequency_sequence,
                  ocl.lal_binary_neutron_star, bilby.gw.conversion.convert_to_lal_binary_neutron_star_parameters,
                  "TaylorF2", True, 1.2)
