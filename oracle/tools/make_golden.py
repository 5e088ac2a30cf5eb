"""BUILD TOOL - golden vectors from the UNMODIFIED reference (bilby imported from /root/reference
with the numpy-only stand-ins of oracle/standins) + the restated IMRPhenomD source model.

Run (build container only):
    PYTHONPATH=oracle/standins:/root/reference python oracle/tools/make_golden.py

Writes tests/golden/bbh_4s_*.npz: parameter draws and the reference's own
``GravitationalWaveTransient.log_likelihood_ratio`` for each marginalisation mode, plus
per-detector (d_inner_h, optimal_snr_squared).  tests/test_oracle_vs_golden.py pins
oracle/cbc_likelihood.py against these; GPU tests pin the CUDA path against them too.
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

bilby.core.utils.logger.setLevel("ERROR")


def build(duration, fs, names, noise_seed=None, wf_args=None):
    inj = dict(ocl.INJECTION)
    start_time = inj["geocent_time"] - duration + 2
    wf_args = wf_args or dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0,
                              minimum_frequency=20.0)
    wfg = bilby.gw.WaveformGenerator(
        duration=duration, sampling_frequency=fs, start_time=start_time,
        frequency_domain_source_model=ocl.lal_binary_black_hole,
        parameter_conversion=bilby.gw.conversion.convert_to_lal_binary_black_hole_parameters,
        waveform_arguments=wf_args)
    ifos = bilby.gw.detector.InterferometerList(names)
    oifos = [ocl.OracleInterferometer(n, fs, duration, start_time) for n in names]
    rng = np.random.default_rng(noise_seed) if noise_seed is not None else None
    for ifo, oifo in zip(ifos, oifos):
        ifo.minimum_frequency = 20.0
        ifo.maximum_frequency = fs / 2
        if rng is None:
            ifo.set_strain_data_from_zero_noise(sampling_frequency=fs, duration=duration,
                                                start_time=start_time)
        else:
            oifo.set_gaussian_noise(rng)
            ifo.set_strain_data_from_frequency_domain_strain(
                oifo.frequency_domain_strain.copy(), sampling_frequency=fs, duration=duration,
                start_time=start_time)
    pols = wfg.frequency_domain_strain(dict(inj))
    for ifo in ifos:
        ifo.inject_signal_from_waveform_polarizations(parameters=dict(inj),
                                                      injection_polarizations=pols)
    return inj, start_time, wfg, ifos


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    n = 64
    draws = ocl.draw_bbh_prior(n, np.random.default_rng(20261017))
    # a few hand-picked edge cases appended: the injection itself, equal mass, extreme spins
    extra = [dict(chirp_mass=28.0956, mass_ratio=29.0 / 36.0, chi_1=0.4, chi_2=0.3, luminosity_distance=2000.0,
                  theta_jn=0.4, psi=2.659, phase=1.3, ra=1.375, dec=-1.2108, geocent_time=1126259642.413),
             dict(chirp_mass=30.0, mass_ratio=1.0, chi_1=0.0, chi_2=0.0, luminosity_distance=1000.0,
                  theta_jn=1.2, psi=0.3, phase=0.2, ra=3.0, dec=0.4, geocent_time=1126259642.35),
             dict(chirp_mass=26.0, mass_ratio=0.2, chi_1=0.99, chi_2=-0.99, luminosity_distance=400.0,
                  theta_jn=2.8, psi=1.3, phase=5.2, ra=5.0, dec=-0.4, geocent_time=1126259642.49)]
    for k in draws:
        draws[k] = np.concatenate([draws[k], [e[k] for e in extra]])
    n = len(draws["chirp_mass"])

    for tag, names, seed in (("zero_H1L1", ["H1", "L1"], None), ("noise_H1L1V1", ["H1", "L1", "V1"], 88170235)):
        inj, start_time, wfg, ifos = build(4.0, 2048.0, names, noise_seed=seed)
        res = dict(start_time=start_time, duration=4.0, sampling_frequency=2048.0,
                   detectors=np.array(names), noise_seed=-1 if seed is None else seed)
        for k in draws:
            res["param_" + k] = draws[k]
        for i, ifo in enumerate(ifos):
            res[f"strain_{ifo.name}"] = ifo.frequency_domain_strain
            res[f"psd_{ifo.name}"] = ifo.power_spectral_density_array
        # plain
        like = bilby.gw.likelihood.GravitationalWaveTransient(ifos, wfg)
        lnl = np.zeros(n)
        dh = np.zeros((n, len(names)), dtype=complex)
        hh = np.zeros((n, len(names)))
        for i in range(n):
            p = {k: float(draws[k][i]) for k in draws}
            lnl[i] = like.log_likelihood_ratio(p)
            p.update(like.get_sky_frame_parameters(p))
            pols = wfg.frequency_domain_strain(p)
            for j, ifo in enumerate(ifos):
                snr = like.calculate_snrs(pols, ifo, parameters=p)
                dh[i, j] = snr.d_inner_h
                hh[i, j] = snr.optimal_snr_squared
        res["lnl_none"] = lnl
        res["d_inner_h"] = dh
        res["optimal_snr_squared"] = hh
        res["noise_log_likelihood"] = like.noise_log_likelihood()
        # phase
        priors = PriorDict(dict(phase=Uniform(0, 2 * np.pi, "phase")))
        like = bilby.gw.likelihood.GravitationalWaveTransient(ifos, wfg, phase_marginalization=True,
                                                              priors=priors)
        res["lnl_phase"] = np.array([like.log_likelihood_ratio({k: float(draws[k][i]) for k in draws})
                                     for i in range(n)])
        # distance + phase (table cached in /tmp to keep reruns quick)
        priors = PriorDict(dict(phase=Uniform(0, 2 * np.pi, "phase"),
                                luminosity_distance=PowerLaw(2, 100.0, 5000.0, "luminosity_distance")))
        like = bilby.gw.likelihood.GravitationalWaveTransient(
            ifos, wfg, phase_marginalization=True, distance_marginalization=True, priors=priors,
            distance_marginalization_lookup_table="/tmp/golden_dp_lookup.npz")
        res["lnl_distance_phase"] = np.array(
            [like.log_likelihood_ratio({k: float(draws[k][i]) for k in draws}) for i in range(n)])
        res["ref_dist"] = like._ref_dist
        # a thin slice of the reference's lookup table (rows 0, 133, 266, 399) to pin table builders
        res["lookup_rows"] = np.array([0, 133, 266, 399])
        res["lookup_table_rows_dp"] = like._dist_margd_loglikelihood_array[[0, 133, 266, 399]]
        # distance only
        priors = PriorDict(dict(luminosity_distance=PowerLaw(2, 100.0, 5000.0, "luminosity_distance")))
        like = bilby.gw.likelihood.GravitationalWaveTransient(
            ifos, wfg, distance_marginalization=True, priors=priors,
            distance_marginalization_lookup_table="/tmp/golden_d_lookup.npz")
        res["lnl_distance"] = np.array(
            [like.log_likelihood_ratio({k: float(draws[k][i]) for k in draws}) for i in range(n)])
        res["lookup_table_rows_d"] = like._dist_margd_loglikelihood_array[[0, 133, 266, 399]]
        # time (+ phase), jitter on
        t_inj = inj["geocent_time"]
        jit = np.random.default_rng(7).uniform(-1 / 2048.0, 1 / 2048.0, n)
        res["param_time_jitter"] = jit
        for mode, kw in (("time", {}), ("time_phase", dict(phase_marginalization=True)),
                         ("time_distance_phase", dict(phase_marginalization=True, distance_marginalization=True,
                                                      distance_marginalization_lookup_table="/tmp/golden_dp_lookup.npz"))):
            pri = dict(geocent_time=Uniform(t_inj - 0.1, t_inj + 0.1, "geocent_time"))
            if "phase_marginalization" in kw:
                pri["phase"] = Uniform(0, 2 * np.pi, "phase")
            if "distance_marginalization" in kw:
                pri["luminosity_distance"] = PowerLaw(2, 100.0, 5000.0, "luminosity_distance")
            like = bilby.gw.likelihood.GravitationalWaveTransient(
                ifos, wfg, time_marginalization=True, jitter_time=True, priors=PriorDict(pri), **kw)
            vals = np.zeros(n)
            for i in range(n):
                p = {k: float(draws[k][i]) for k in draws}
                p["geocent_time"] = float(start_time)
                p["time_jitter"] = float(jit[i])
                vals[i] = like.log_likelihood_ratio(p)
            res["lnl_" + mode] = vals
        np.savez_compressed(os.path.join(out_dir, f"bbh_4s_{tag}.npz"), **res)
        print(tag, "lnl_none[:4]", lnl[:4], "dp", res["lnl_distance_phase"][:4], "t", res["lnl_time"][:4])


if __name__ == "__main__":
    main()
