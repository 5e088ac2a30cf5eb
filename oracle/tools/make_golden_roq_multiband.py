"""BUILD TOOL - golden vectors for ROQ weights built from a MULTIBANDED basis
(ROQGravitationalWaveTransient._set_weights_linear_multiband / _set_weights_quadratic_multiband,
bilby/gw/likelihood/roq.py:920-974, 1006-1053) from the UNMODIFIED reference.

    PYTHONPATH=oracle/standins:/root/reference python oracle/tools/make_golden_roq_multiband.py

The reference reads such bases from hdf5 files (h5py is absent here), but the two weight builders only index their
`linear_matrix` / `quadratic_matrix` argument like a mapping, so they are called directly with nested dicts of arrays on
a likelihood object that the reference constructed from an ordinary basis; the likelihood is then evaluated with the
multibanded weights.  Writes tests/golden/roq_multiband_bbh_4s_H1L1V1.npz."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

import bilby  # noqa: E402
from bilby.core.prior import PriorDict, Uniform  # noqa: E402
from oracle import cbc_likelihood as ocl, cbc_reduced as ocr  # noqa: E402
from make_golden_reduced import make_ifos, near, evaluate, T_INJ, NOISE_SEED  # noqa: E402

bilby.core.utils.logger.setLevel("ERROR")

# three bands: 4 s up to 64 Hz, 2 s up to 128 Hz, 1 s up to 512 Hz (bins are in units of 1 / T_b)
DURATIONS = np.array([4.0, 2.0, 1.0])
BINS = np.array([[80, 255], [128, 255], [128, 512]])


def main():
    names = ["H1", "L1", "V1"]
    duration, fs, fmax = 4.0, 2048.0, 512.0
    inj = dict(ocl.INJECTION)
    start_time = T_INJ - duration + 0.4        # merger inside the last second: every band's cropped data holds it
    ifos = make_ifos(duration, fs, names, start_time, maximum_frequency=fmax)
    conv = bilby.gw.conversion.convert_to_lal_binary_black_hole_parameters
    wfg_full = bilby.gw.WaveformGenerator(
        duration=duration, sampling_frequency=fs, start_time=start_time,
        frequency_domain_source_model=ocl.lal_binary_black_hole, parameter_conversion=conv,
        waveform_arguments=dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0, minimum_frequency=20.0))
    pols = wfg_full.frequency_domain_strain(dict(inj))
    for ifo in ifos:
        ifo.inject_signal_from_waveform_polarizations(parameters=dict(inj), injection_polarizations=pols)
    banded = np.concatenate([np.arange(k0, k1 + 1) / tb for (k0, k1), tb in zip(BINS, DURATIONS)])
    rng = np.random.default_rng(5)
    train = near(inj, 160, rng)

    def h22_on(freqs):
        def fn(i):
            p, _ = conv({k: float(v[i]) for k, v in train.items()})
            return ocr._sequence_polarizations(freqs, p["mass_1"], p["mass_2"], 1.0, p["a_1"], p["tilt_1"], p["a_2"],
                                               p["tilt_2"], 0.0, 0.0, 0.0, 0.0, "IMRPhenomD", 20.0)["plus"]
        return fn
    mb = ocr.build_synthetic_roq_basis(banded, h22_on(banded), range(160), 20, 10)
    lin = mb["linear_matrix"].astype(np.complex64).T          # [n_basis, basis_dimension]
    quad = mb["quadratic_matrix"].astype(np.complex64).T
    fnl, fnq = mb["frequency_nodes_linear"], mb["frequency_nodes_quadratic"]
    # an ordinary basis of the same sizes on the masked grid, only to let the reference construct the object
    mask = ifos[0].frequency_mask
    freqs = ifos[0].frequency_array[mask]
    plain = ocr.build_synthetic_roq_basis(freqs, h22_on(freqs), range(160), 20, 10)
    wa = dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0)
    wfg = bilby.gw.WaveformGenerator(
        duration=duration, sampling_frequency=fs, start_time=start_time,
        frequency_domain_source_model=ocr.binary_black_hole_roq, parameter_conversion=conv,
        waveform_arguments=dict(wa, frequency_nodes_linear=fnl, frequency_nodes_quadratic=fnq))
    pri = PriorDict(dict(geocent_time=Uniform(T_INJ - 0.1, T_INJ + 0.1, "geocent_time")))
    like = bilby.gw.likelihood.ROQGravitationalWaveTransient(
        ifos, wfg, pri, linear_matrix=plain["linear_matrix"].astype(complex),
        quadratic_matrix=plain["quadratic_matrix"].astype(complex))
    lin_dict = dict(multiband_linear=np.array(True), durations_s_linear=DURATIONS, start_end_frequency_bins_linear=BINS,
                    basis_linear={"0": dict(basis=lin.astype(complex))})
    quad_dict = dict(multiband_quadratic=np.array(True), durations_s_quadratic=DURATIONS,
                     start_end_frequency_bins_quadratic=BINS, basis_quadratic={"0": dict(basis=quad.astype(complex))})
    like._set_weights_linear_multiband(lin_dict, [0])
    like._set_weights_quadratic_multiband(quad_dict, [0])
    n = 24
    draws = near(inj, n, np.random.default_rng(20261017))
    res = dict(start_time=start_time, duration=duration, sampling_frequency=fs, detectors=np.array(names),
               noise_seed=NOISE_SEED, maximum_frequency=fmax, durations_s=DURATIONS, start_end_frequency_bins=BINS,
               basis_linear=lin, basis_quadratic=quad, frequency_nodes_linear=fnl, frequency_nodes_quadratic=fnq,
               time_samples=like.weights["time_samples"],
               optimal_snrs=np.array([ifo.meta_data["optimal_SNR"] for ifo in ifos]))
    for k in draws:
        res["param_" + k] = draws[k]
    for ifo in ifos:
        res[f"weights_{ifo.name}_linear"] = like.weights[f"{ifo.name}_linear"][0][::37]     # every 37th ROQ time
        res[f"weights_{ifo.name}_quadratic"] = like.weights[f"{ifo.name}_quadratic"][0]
    with np.errstate(divide="ignore", invalid="ignore"):
        res["lnl_none"] = evaluate(like, draws, n)
    # the full-grid likelihood of the same draws (the multibanded ROQ approximates it)
    full = bilby.gw.likelihood.GravitationalWaveTransient(ifos, wfg_full)
    res["lnl_full_grid"] = evaluate(full, draws, n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "roq_multiband_bbh_4s_H1L1V1.npz"), **res)
    print("time samples", len(res["time_samples"]), "lnl", res["lnl_none"][:4], "full grid", res["lnl_full_grid"][:4])


if __name__ == "__main__":
    main()
