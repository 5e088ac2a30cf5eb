"""BUILD TOOL - golden vectors for the reduced-order likelihoods from the
UNMODIFIED reference classes RelativeBinningGravitationalWaveTransient and ROQGravitationalWaveTransient
(bilby imported from /root/reference with the stand-ins), fed by the restated source models of
oracle/cbc_reduced.py (lalsimulation is absent).

    PYTHONPATH=oracle/standins:/root/reference python oracle/tools/make_golden_reduced.py

Writes tests/golden/relbin_bbh_4s_H1L1V1.npz, relbin_bns_32s_H1L1V1.npz and roq_bbh_4s_H1L1V1.npz.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

import bilby  # noqa: E402  (the reference)
from bilby.core.prior import Uniform, PowerLaw, PriorDict  # noqa: E402
from oracle import cbc_likelihood as ocl  # noqa: E402
from oracle import cbc_reduced as ocr  # noqa: E402

bilby.core.utils.logger.setLevel("ERROR")
NOISE_SEED = 88170235
T_INJ = ocl.INJECTION["geocent_time"]
BNS_INJ = dict(mass_1=1.5, mass_2=1.3, chi_1=0.02, chi_2=0.01, luminosity_distance=100.0, theta_jn=0.4, psi=2.659,
               phase=1.3, geocent_time=T_INJ, ra=1.375, dec=-1.2108, lambda_1=400.0, lambda_2=600.0)


def make_ifos(duration, fs, names, start_time, maximum_frequency=None):
    ifos = bilby.gw.detector.InterferometerList(names)
    rng = np.random.default_rng(NOISE_SEED)
    for ifo in ifos:
        ifo.minimum_frequency = 20.0
        ifo.maximum_frequency = fs / 2 if maximum_frequency is None else maximum_frequency
        o = ocl.OracleInterferometer(ifo.name, fs, duration, start_time, maximum_frequency=maximum_frequency)
        o.set_gaussian_noise(rng)
        ifo.set_strain_data_from_frequency_domain_strain(o.frequency_domain_strain.copy(), sampling_frequency=fs,
                                                         duration=duration, start_time=start_time)
    return ifos


def near(inj, n, rng, bns=False):
    """Draws in a small box around the injection (where relative binning / the synthetic ROQ basis are valid)."""
    m1, m2 = inj["mass_1"], inj["mass_2"]
    mc = (m1 * m2) ** 0.6 / (m1 + m2) ** 0.2
    d = dict(chirp_mass=mc * (1 + rng.uniform(-2e-3, 2e-3, n) * (0.05 if bns else 1)),
             mass_ratio=np.clip(m2 / m1 + rng.uniform(-0.05, 0.05, n), 0.2, 1.0),
             chi_1=inj["chi_1"] + rng.uniform(-0.02, 0.02, n), chi_2=inj["chi_2"] + rng.uniform(-0.02, 0.02, n),
             luminosity_distance=inj["luminosity_distance"] * rng.uniform(0.7, 1.5, n),
             theta_jn=inj["theta_jn"] + rng.uniform(-0.2, 0.2, n), psi=inj["psi"] + rng.uniform(-0.2, 0.2, n),
             phase=rng.uniform(0, 2 * np.pi, n), ra=inj["ra"] + rng.uniform(-0.05, 0.05, n),
             dec=inj["dec"] + rng.uniform(-0.05, 0.05, n),
             geocent_time=inj["geocent_time"] + rng.uniform(-2e-3, 2e-3, n))
    if bns:
        d["lambda_1"] = inj["lambda_1"] + rng.uniform(-200, 200, n)
        d["lambda_2"] = inj["lambda_2"] + rng.uniform(-200, 200, n)
    return d


def evaluate(like, draws, n, extra=None):
    out = np.zeros(n)
    for i in range(n):
        p = {k: float(v[i]) for k, v in draws.items()}
        if extra:
            p.update(extra(i))
        out[i] = like.log_likelihood_ratio(p)
    return out


def relbin(tag, duration, fs, inj, model, conversion, approximant, bns):
    names = ["H1", "L1", "V1"]
    start_time = T_INJ - duration + 2
    wa = dict(waveform_approximant=approximant, reference_frequency=50.0, minimum_frequency=20.0)
    ifos = make_ifos(duration, fs, names, start_time)

    def wfg_new():
        return bilby.gw.WaveformGenerator(duration=duration, sampling_frequency=fs, start_time=start_time,
                                          frequency_domain_source_model=model, parameter_conversion=conversion,
                                          waveform_arguments=dict(wa))
    wfg = wfg_new()
    wfg.waveform_arguments["fiducial"] = 1
    pols = wfg.frequency_domain_strain(dict(inj))
    for ifo in ifos:
        ifo.inject_signal_from_waveform_polarizations(parameters=dict(inj), injection_polarizations=pols)
    n = 24
    rng = np.random.default_rng(20261017)
    draws = near(inj, n, rng, bns=bns)
    # the injection itself first
    for k in draws:
        base = inj.get(k)
        if k == "chirp_mass":
            base = (inj["mass_1"] * inj["mass_2"]) ** 0.6 / (inj["mass_1"] + inj["mass_2"]) ** 0.2
        if k == "mass_ratio":
            base = inj["mass_2"] / inj["mass_1"]
        draws[k][0] = base
    fid = {k: float(v[0]) for k, v in draws.items()}
    res = dict(start_time=start_time, duration=duration, sampling_frequency=fs, detectors=np.array(names),
               noise_seed=NOISE_SEED, approximant=approximant)
    for k in draws:
        res["param_" + k] = draws[k]
    for k, v in inj.items():
        res["inj_" + k] = v
    like = bilby.gw.likelihood.RelativeBinningGravitationalWaveTransient(
        ifos, wfg_new(), fiducial_parameters=dict(fid), epsilon=0.5, chi=1)
    res["bin_freqs"] = like.bin_freqs
    res["bin_inds"] = like.bin_inds
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
    for j, ifo in enumerate(ifos):
        a0, a1, b0, b1 = like.summary_data[ifo.name]
        res[f"summary_{ifo.name}"] = np.array([a0, a1, b0, b1])
    pri = PriorDict(dict(phase=Uniform(0, 2 * np.pi, "phase")))
    like = bilby.gw.likelihood.RelativeBinningGravitationalWaveTransient(
        ifos, wfg_new(), fiducial_parameters=dict(fid), priors=pri, phase_marginalization=True)
    res["lnl_phase"] = evaluate(like, draws, n)
    dmin, dmax = (10.0, 500.0) if bns else (100.0, 5000.0)
    pri = PriorDict(dict(phase=Uniform(0, 2 * np.pi, "phase"),
                         luminosity_distance=PowerLaw(2, dmin, dmax, "luminosity_distance")))
    like = bilby.gw.likelihood.RelativeBinningGravitationalWaveTransient(
        ifos, wfg_new(), fiducial_parameters=dict(fid), priors=pri, phase_marginalization=True,
        distance_marginalization=True, distance_marginalization_lookup_table=f"/tmp/golden_rb_{tag}_lookup.npz")
    res["lnl_distance_phase"] = evaluate(like, draws, n)
    res["distance_prior"] = np.array([dmin, dmax])
    # time (+phase) marginalisation: full waveform reconstruction + FFT (relative.py:380-421)
    jit = np.random.default_rng(7).uniform(-1 / fs, 1 / fs, n)
    res["param_time_jitter"] = jit
    pri = PriorDict(dict(phase=Uniform(0, 2 * np.pi, "phase"),
                         geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time")))
    like = bilby.gw.likelihood.RelativeBinningGravitationalWaveTransient(
        ifos, wfg_new(), fiducial_parameters=dict(fid, time_jitter=0.0), priors=pri, phase_marginalization=True,
        time_marginalization=True, jitter_time=True)
    res["lnl_time_phase"] = evaluate(like, draws, n, extra=lambda i: dict(geocent_time=float(start_time),
                                                                           time_jitter=float(jit[i])))
    res["bin_freqs_time"] = like.bin_freqs
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"relbin_{tag}.npz"), **res)
    print(tag, "bins", len(res["bin_freqs"]) - 1, "lnl", res["lnl_none"][:3], res["lnl_distance_phase"][:3],
          res["lnl_time_phase"][:3])


def roq():
    names = ["H1", "L1", "V1"]
    duration, fs, fmax = 4.0, 2048.0, 512.0
    inj = dict(ocl.INJECTION)
    start_time = T_INJ - duration + 2
    ifos = make_ifos(duration, fs, names, start_time, maximum_frequency=fmax)
    wa = dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0)
    conv = bilby.gw.conversion.convert_to_lal_binary_black_hole_parameters
    # injection on the full grid
    wfg_full = bilby.gw.WaveformGenerator(
        duration=duration, sampling_frequency=fs, start_time=start_time,
        frequency_domain_source_model=ocl.lal_binary_black_hole, parameter_conversion=conv,
        waveform_arguments=dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0, minimum_frequency=20.0))
    pols = wfg_full.frequency_domain_strain(dict(inj))
    for ifo in ifos:
        ifo.inject_signal_from_waveform_polarizations(parameters=dict(inj), injection_polarizations=pols)
    # synthetic basis on the masked grid, trained near the injection; rounded to complex64 so that the
    # committed fixture reproduces it exactly
    mask = ifos[0].frequency_mask
    freqs = ifos[0].frequency_array[mask]
    rng = np.random.default_rng(5)
    train = near(inj, 160, rng)

    def h22(i):
        p, _ = conv({k: float(v[i]) for k, v in train.items()})
        out = ocr._sequence_polarizations(freqs, p["mass_1"], p["mass_2"], 1.0, p["a_1"], p["tilt_1"], p["a_2"],
                                          p["tilt_2"], 0.0, 0.0, 0.0, 0.0, "IMRPhenomD", 20.0)
        return out["plus"]
    basis = ocr.build_synthetic_roq_basis(freqs, h22, range(160), 20, 10)
    lin = basis["linear_matrix"].astype(np.complex64)
    quad = basis["quadratic_matrix"].astype(np.complex64)
    fnl, fnq = basis["frequency_nodes_linear"], basis["frequency_nodes_quadratic"]

    def wfg_new():
        return bilby.gw.WaveformGenerator(
            duration=duration, sampling_frequency=fs, start_time=start_time,
            frequency_domain_source_model=ocr.binary_black_hole_roq, parameter_conversion=conv,
            waveform_arguments=dict(wa, frequency_nodes_linear=fnl, frequency_nodes_quadratic=fnq))
    n = 24
    draws = near(inj, n, np.random.default_rng(20261017))
    # two draws outside the ROQ time window: the reference returns -inf there (roq.py:532-533)
    draws["geocent_time"][-1] = T_INJ + 0.5
    draws["geocent_time"][-2] = T_INJ - 0.5
    res = dict(start_time=start_time, duration=duration, sampling_frequency=fs, detectors=np.array(names),
               noise_seed=NOISE_SEED, maximum_frequency=fmax, linear_matrix=lin, quadratic_matrix=quad,
               frequency_nodes_linear=fnl, frequency_nodes_quadratic=fnq)
    for k in draws:
        res["param_" + k] = draws[k]
    tprior = Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time")
    pri = PriorDict(dict(geocent_time=tprior))
    like = bilby.gw.likelihood.ROQGravitationalWaveTransient(
        ifos, wfg_new(), pri, linear_matrix=lin.astype(complex), quadratic_matrix=quad.astype(complex))
    res["time_samples"] = like.weights["time_samples"]
    res["optimal_snrs"] = np.array([ifo.meta_data["optimal_SNR"] for ifo in ifos])
    res["weights_H1_linear_row0"] = like.weights["H1_linear"][0][0]
    res["weights_H1_quadratic"] = like.weights["H1_quadratic"][0]
    with np.errstate(divide="ignore", invalid="ignore"):
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
    pri = PriorDict(dict(geocent_time=tprior, phase=Uniform(0, 2 * np.pi, "phase"),
                         luminosity_distance=PowerLaw(2, 100.0, 5000.0, "luminosity_distance")))
    like = bilby.gw.likelihood.ROQGravitationalWaveTransient(
        ifos, wfg_new(), pri, linear_matrix=lin.astype(complex), quadratic_matrix=quad.astype(complex),
        phase_marginalization=True, distance_marginalization=True,
        distance_marginalization_lookup_table="/tmp/golden_dp_lookup.npz")
    with np.errstate(divide="ignore", invalid="ignore"):
        res["lnl_distance_phase"] = evaluate(like, draws, n)
    # time + phase marginalisation: the all-times contraction W @ conj(h_linear) (roq.py:604-651)
    pri = PriorDict(dict(geocent_time=tprior, phase=Uniform(0, 2 * np.pi, "phase")))
    like = bilby.gw.likelihood.ROQGravitationalWaveTransient(
        ifos, wfg_new(), pri, linear_matrix=lin.astype(complex), quadratic_matrix=quad.astype(complex),
        phase_marginalization=True, time_marginalization=True, jitter_time=True)
    jit = np.random.default_rng(7).uniform(-like._delta_tc / 2, like._delta_tc / 2, n)
    res["param_time_jitter"] = jit
    res["delta_tc"] = like._delta_tc
    res["lnl_time_phase"] = evaluate(like, draws, n, extra=lambda i: dict(
        geocent_time=float(pri["geocent_time"]), time_jitter=float(jit[i])))
    res["time_marg_geocent_time"] = float(pri["geocent_time"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "roq_bbh_4s_H1L1V1.npz"), **res)
    print("roq", "time samples", len(res["time_samples"]), "lnl", res["lnl_none"][:3], res["lnl_none"][-2:],
          res["lnl_distance_phase"][:3], res["lnl_time_phase"][:3])


if __name__ == "__main__":
    what = sys.argv[1:] or ["relbin_bbh", "relbin_bns", "roq"]
    if "relbin_bbh" in what:
        relbin("bbh_4s_H1L1V1", 4.0, 2048.0, dict(ocl.INJECTION), ocr.lal_binary_black_hole_relative_binning,
               bilby.gw.conversion.convert_to_lal_binary_black_hole_parameters, "IMRPhenomD", False)
    if "relbin_bns" in what:
        relbin("bns_32s_H1L1V1", 32.0, 4096.0, dict(BNS_INJ), ocr.lal_binary_neutron_star_relative_binning,
               bilby.gw.conversion.convert_to_lal_binary_neutron_star_parameters, "TaylorF2", True)
    if "roq" in what:
        roq()
