"""SURVEY.md section 8f rank 1 end to end: the batched ensemble sampler (bilby plugin "b200_ensemble") drives
``log_likelihood_ratio_batch`` on a 4 s H1L1V1 zero-noise injection; the points it visited are replayed through the
ORACLE (one parameter dict at a time, the reference's calling pattern) and the posterior brackets the injection."""
import numpy as np
import pytest

import baseline_common as bc
from oracle import cbc_likelihood as ocl

pytestmark = pytest.mark.gpu


def test_ensemble_sampler_on_gw_likelihood_visited_points_vs_oracle():
    import bilby_b200 as bb
    from bilby_b200.core.prior import PriorDict, Uniform, PowerLaw
    from bilby_b200.core.sampler import run_sampler
    from bilby_b200.gw.detector import InterferometerList
    from bilby_b200.gw.source import lal_binary_black_hole
    from bilby_b200.workloads import INJECTION
    inj = dict(INJECTION)
    start = inj["geocent_time"] - 2.0
    wa = dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0)
    wfg = bb.gw.WaveformGenerator(duration=4.0, sampling_frequency=2048.0, start_time=start,
                                  frequency_domain_source_model=lal_binary_black_hole, waveform_arguments=dict(wa))
    ifos = InterferometerList(["H1", "L1", "V1"])
    ifos.set_strain_data_from_zero_noise(2048.0, 4.0, start)
    ifos.inject_signal(parameters=inj, waveform_generator=wfg)
    mc = (36.0 * 29.0) ** 0.6 / 65.0 ** 0.2
    priors = PriorDict({k: v for k, v in inj.items() if k not in ("mass_1", "mass_2", "phase")})
    priors["chirp_mass"] = Uniform(mc - 2.0, mc + 2.0, "chirp_mass")
    priors["mass_ratio"] = Uniform(0.4, 1.0, "mass_ratio")
    priors["luminosity_distance"] = PowerLaw(2, 500.0, 5000.0, "luminosity_distance")
    priors["phase"] = Uniform(0, 2 * np.pi, "phase")
    like = bb.gw.GravitationalWaveTransient(ifos, wfg, phase_marginalization=True, priors=priors)
    assert priors["phase"] == 0.0                      # base.py:205-207: the sampler never sees the marginalised phase
    res = run_sampler(like, priors, nwalkers=512, nsteps=300, seed=11, record_visited=True)
    assert res["search_parameter_keys"] == ["luminosity_distance", "chirp_mass", "mass_ratio"]
    assert res["num_likelihood_evaluations"] > 1e5
    # visited points vs the oracle (phase-marginalised likelihood, one dict per call)
    theta, lnl = res["visited_theta"], res["visited_log_likelihood"]
    pick = np.random.default_rng(0).choice(len(theta), 2000, replace=False)
    draws = {k: np.full(len(pick), float(v)) for k, v in priors.items() if not hasattr(v, "rescale")}
    for j, key in enumerate(res["search_parameter_keys"]):
        draws[key] = theta[pick, j]
    olike = ocl.OracleLikelihood(bc.oracle_ifos_like(ifos), waveform_arguments=dict(wa), phase_marginalization=True)
    ref = bc.oracle_map(olike, draws, len(pick))
    hh = bc.total_optimal_snr_squared(like, draws)
    err = np.abs(lnl[pick] - ref) / bc.scale_of(ref, hh)
    assert err.max() < 1e-8, err.max()
    # the posterior (after burn-in) brackets the injected values
    s = res["samples"]
    truth = dict(luminosity_distance=inj["luminosity_distance"], chirp_mass=mc, mass_ratio=29.0 / 36.0)
    for j, key in enumerate(res["search_parameter_keys"]):
        lo, hi = np.percentile(s[:, j], [0.5, 99.5])
        assert lo < truth[key] < hi, (key, lo, truth[key], hi)
    snr2 = sum(ifo.meta_data["optimal_SNR"] ** 2 for ifo in ifos)
    assert res["log_likelihood_evaluations"].max() > 0.95 * 0.5 * snr2       # the chain found the peak (lnL ~ SNR^2 / 2)
    print(f"sampler: {res['num_likelihood_evaluations']} evaluations in {res['sampling_time']:.2f} s, acceptance "
          f"{res['acceptance_fraction']:.2f}, max scaled |dlnL| vs oracle {err.max():.2e}")
