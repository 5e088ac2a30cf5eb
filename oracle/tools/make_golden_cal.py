"""BUILD TOOL - golden vectors for the calibration path (BASELINE.json configs[2]): BBH 8 s H1L1V1, CubicSpline
calibration (10 nodes per detector), plain and time(+phase)-marginalised likelihood, from the UNMODIFIED
reference.   PYTHONPATH=oracle/standins:/root/reference python oracle/tools/make_golden_cal.py"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

import bilby  # noqa: E402
from bilby.core.prior import Uniform, PriorDict  # noqa: E402
from oracle import cbc_likelihood as ocl  # noqa: E402

bilby.core.utils.logger.setLevel("ERROR")


def main():
    duration, fs, names = 8.0, 2048.0, ["H1", "L1", "V1"]
    inj = dict(ocl.INJECTION)
    start_time = inj["geocent_time"] - duration + 2
    wfg = bilby.gw.WaveformGenerator(
        duration=duration, sampling_frequency=fs, start_time=start_time,
        frequency_domain_source_model=ocl.lal_binary_black_hole,
        parameter_conversion=bilby.gw.conversion.convert_to_lal_binary_black_hole_parameters,
        waveform_arguments=dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0))
    ifos = bilby.gw.detector.InterferometerList(names)
    rng = np.random.default_rng(88170235)
    for ifo in ifos:
        ifo.minimum_frequency = 20.0
        ifo.maximum_frequency = fs / 2
        o = ocl.OracleInterferometer(ifo.name, fs, duration, start_time)
        o.set_gaussian_noise(rng)
        ifo.set_strain_data_from_frequency_domain_strain(o.frequency_domain_strain.copy(), sampling_frequency=fs,
                                                         duration=duration, start_time=start_time)
    pols = wfg.frequency_domain_strain(dict(inj))
    for ifo in ifos:
        ifo.inject_signal_from_waveform_polarizations(parameters=dict(inj), injection_polarizations=pols)
    for ifo in ifos:
        ifo.calibration_model = bilby.gw.calibration.CubicSpline(
            prefix=f"recalib_{ifo.name}_", minimum_frequency=ifo.minimum_frequency,
            maximum_frequency=ifo.maximum_frequency, n_points=10)
    n = 24
    draws = ocl.draw_bbh_prior(n, np.random.default_rng(20261017))
    crng = np.random.default_rng(99)
    for name in names:
        for i in range(10):
            draws[f"recalib_{name}_amplitude_{i}"] = crng.normal(0, 0.05, n)
            draws[f"recalib_{name}_phase_{i}"] = crng.normal(0, 0.05, n)
    res = dict(start_time=start_time, duration=duration, sampling_frequency=fs, detectors=np.array(names))
    for k in draws:
        res["param_" + k] = draws[k]
    for ifo in ifos:
        res[f"strain_{ifo.name}"] = ifo.frequency_domain_strain
    like = bilby.gw.likelihood.GravitationalWaveTransient(ifos, wfg)
    res["lnl_none"] = np.array([like.log_likelihood_ratio({k: float(draws[k][i]) for k in draws}) for i in range(n)])
    hh = np.zeros((n, 3))
    for i in range(n):
        p = {k: float(draws[k][i]) for k in draws}
        pl = wfg.frequency_domain_strain(p)
        for j, ifo in enumerate(ifos):
            hh[i, j] = like.calculate_snrs(pl, ifo, parameters=p).optimal_snr_squared
    res["optimal_snr_squared"] = hh
    t_inj = inj["geocent_time"]
    jit = np.random.default_rng(7).uniform(-1 / 2048.0, 1 / 2048.0, n)
    res["param_time_jitter"] = jit
    for mode, kw in (("time", {}), ("time_phase", dict(phase_marginalization=True))):
        pri = dict(geocent_time=Uniform(t_inj - 0.1, t_inj + 0.1, "geocent_time"))
        if kw:
            pri["phase"] = Uniform(0, 2 * np.pi, "phase")
        like = bilby.gw.likelihood.GravitationalWaveTransient(ifos, wfg, time_marginalization=True, jitter_time=True,
                                                              priors=PriorDict(pri), **kw)
        vals = np.zeros(n)
        for i in range(n):
            p = {k: float(draws[k][i]) for k in draws}
            p["geocent_time"] = float(start_time)
            p["time_jitter"] = float(jit[i])
            vals[i] = like.log_likelihood_ratio(p)
        res["lnl_" + mode] = vals
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bbh_8s_cal_H1L1V1.npz"), **res)
    print(res["lnl_none"][:4], res["lnl_time"][:4])


if __name__ == "__main__":
    main()
