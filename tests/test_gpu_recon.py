"""GPU parity of the batched marginalised-parameter reconstruction (bb_reconstruct_marginalized_device) and of the
batched per-detector SNRs against the UNMODIFIED reference's outputs (tests/golden/recon_4s_H1L1V1.npz, made by
oracle/tools/make_golden_recon.py: base.py:502-773 and conversion.py:2215-2288 with the same unit-interval draws)."""
import os

import numpy as np
import pytest

from test_gpu_parity import _build, _priors

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
MODES = {
    "phase": dict(phase=True),
    "distance": dict(luminosity_distance=True),
    "distance_phase": dict(luminosity_distance=True, phase=True),
    "time": dict(geocent_time=True),
    "time_phase": dict(geocent_time=True, phase=True),
    "time_distance_phase": dict(geocent_time=True, luminosity_distance=True, phase=True),
}


def _likelihood(mode):
    on = MODES[mode]
    kw = dict(phase_marginalization=on.get("phase", False), distance_marginalization=on.get("luminosity_distance", False),
              time_marginalization=on.get("geocent_time", False), priors=_priors(**on))
    if kw["time_marginalization"]:
        kw["jitter_time"] = True
    _, like, _ = _build("noise_H1L1V1", **kw)
    g = np.load(os.path.join(GOLDEN, "recon_4s_H1L1V1.npz"))
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_")}
    return g, like, draws


@pytest.mark.parametrize("mode", list(MODES))
def test_reconstruction_vs_reference(mode):
    g, like, draws = _likelihood(mode)
    d = {k: v for k, v in draws.items() if "time" in mode or k != "time_jitter"}
    uni = np.nan_to_num(g["uniforms_" + mode], nan=0.5)
    new = like.generate_posterior_samples_from_marginalized_likelihood_batch(d, uniforms=uni)
    ref = g["recon_" + mode]
    on = MODES[mode]
    if on.get("geocent_time"):
        # the time grid itself is only resolved to ~2.4e-7 s in float64 at GPS 1.1e9 s
        assert np.abs(new["geocent_time"] - ref[:, 0]).max() < 1e-9
    else:
        assert np.array_equal(new["geocent_time"], draws["geocent_time"])
    if on.get("luminosity_distance"):
        assert np.abs(new["luminosity_distance"] / ref[:, 1] - 1).max() < 1e-9
    else:
        assert np.array_equal(new["luminosity_distance"], draws["luminosity_distance"])
    if on.get("phase"):
        assert np.abs(new["phase"] - ref[:, 2]).max() < 1e-9
    else:
        assert np.array_equal(new["phase"], draws["phase"])


def test_scalar_reconstruction_is_a_batch_of_one():
    g, like, draws = _likelihood("distance_phase")
    p = {k: float(v[13]) for k, v in draws.items() if k != "time_jitter"}
    rng = np.random.default_rng(1013)
    new = like.generate_posterior_sample_from_marginalized_likelihood(p, rng=rng)
    u = np.random.default_rng(1013).uniform(0, 1, size=(1, 3))
    batch = like.generate_posterior_samples_from_marginalized_likelihood_batch(
        {k: np.array([v]) for k, v in p.items()}, uniforms=u)
    assert new["luminosity_distance"] == batch["luminosity_distance"][0]
    assert new["phase"] == batch["phase"][0]
    assert 100.0 <= new["luminosity_distance"] <= 5000.0 and 0.0 <= new["phase"] <= 2 * np.pi


def test_reconstruction_large_batch_statistics():
    """Property at batch scale: with uniform draws the reconstructed phases of one row follow the analytic phase
    posterior exp(Re(<d|h> e^{-2 i phi})) (base.py:766-771) - checked through its first circular moment."""
    g, like, draws = _likelihood("phase")
    n = 20000
    row = {k: np.full(n, float(v[12])) for k, v in draws.items() if k != "time_jitter"}
    u = np.random.default_rng(3).uniform(0, 1, size=(n, 3))
    new = like.generate_posterior_samples_from_marginalized_likelihood_batch(row, uniforms=u)
    snr = like.compute_snrs_batch({k: v[:1] for k, v in row.items()})
    dih = sum(snr[f"{ifo.name}_matched_filter_snr"][0] * snr[f"{ifo.name}_optimal_snr"][0] for ifo in like.interferometers)
    phi = np.linspace(0, 2 * np.pi, 20001)
    w = np.exp(np.real(dih * np.exp(-2j * phi)) - np.abs(dih))
    expect = np.trapezoid(w * np.exp(2j * phi), phi) / np.trapezoid(w, phi)
    got = np.mean(np.exp(2j * new["phase"]))
    assert abs(got - expect) < 5 / np.sqrt(n)


def test_per_detector_snrs_vs_reference():
    g, like, draws = _likelihood("phase")
    d = {k: v for k, v in draws.items() if k != "time_jitter"}
    snr = like.compute_snrs_batch(d)
    for j, ifo in enumerate(like.interferometers):
        mf = snr[f"{ifo.name}_matched_filter_snr"]
        assert np.abs(mf - g["matched_filter_snr"][:, j]).max() < 1e-8 * np.abs(g["matched_filter_snr"][:, j]).max()
        assert np.abs(snr[f"{ifo.name}_optimal_snr"] / g["optimal_snr"][:, j] - 1).max() < 1e-8


def test_reconstruction_8s_vs_reference():
    """duration != 4 s: the reference's 16384 Hz transform carries no 4/T factor (base.py:626); zero-noise H1+L1
    injection regenerated here (device waveform), time + phase marginalisation."""
    import bilby_b200 as bb
    from bilby_b200.gw.detector import InterferometerList
    from bilby_b200.gw.source import lal_binary_black_hole
    from bilby_b200.workloads import INJECTION
    g = np.load(os.path.join(GOLDEN, "recon_8s_zero_H1L1.npz"))
    start = float(g["start_time"])
    wfg = bb.gw.WaveformGenerator(duration=8.0, sampling_frequency=2048.0, start_time=start,
                                  frequency_domain_source_model=lal_binary_black_hole,
                                  waveform_arguments=dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0,
                                                          minimum_frequency=20.0))
    ifos = InterferometerList([str(n) for n in g["detectors"]])
    ifos.set_strain_data_from_zero_noise(2048.0, 8.0, start)
    ifos.inject_signal(parameters=dict(INJECTION), waveform_generator=wfg)
    like = bb.gw.GravitationalWaveTransient(ifos, wfg, time_marginalization=True, phase_marginalization=True,
                                            jitter_time=True, priors=_priors(geocent_time=True, phase=True))
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_")}
    uni = np.nan_to_num(g["uniforms_time_phase"], nan=0.5)
    new = like.generate_posterior_samples_from_marginalized_likelihood_batch(draws, uniforms=uni)
    ref = g["recon_time_phase"]
    assert np.abs(new["geocent_time"] - ref[:, 0]).max() < 1e-9
    assert np.abs(new["phase"] - ref[:, 2]).max() < 1e-8


def test_conversion_module_mirrors_on_a_dataframe():
    """bilby.gw.conversion.compute_snrs / generate_posterior_samples_from_marginalized_likelihood on a DataFrame
    (conversion.py:2215-2271, 2400-2492)."""
    import pandas as pd
    from bilby_b200.gw import conversion
    g, like, draws = _likelihood("distance_phase")
    frame = pd.DataFrame({k: v for k, v in draws.items() if k != "time_jitter"})
    before = frame["luminosity_distance"].to_numpy().copy()
    out = conversion.generate_posterior_samples_from_marginalized_likelihood(frame, like, rng=np.random.default_rng(0))
    assert out is frame and not np.array_equal(frame["luminosity_distance"].to_numpy(), before)
    assert ((frame["luminosity_distance"] >= 100.0) & (frame["luminosity_distance"] <= 5000.0)).all()
    conversion.compute_snrs(frame, like)
    assert np.iscomplexobj(frame["H1_matched_filter_snr"].to_numpy()) and (frame["L1_optimal_snr"] > 0).all()
    d = dict(luminosity_distance=1.0)
    assert conversion.generate_posterior_samples_from_marginalized_likelihood(d, like) is d
