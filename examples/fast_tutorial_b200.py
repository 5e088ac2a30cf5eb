#!/usr/bin/env python
"""The reference's examples/gw_examples/injection_examples/fast_tutorial.py on a B200: a 4 s binary-black-hole injection
in H1 + L1, distance + phase marginalised likelihood, sampled with the batched ensemble sampler (every half-ensemble
move is ONE launch of the fused kernels).  Needs a B200 (sm_100a); there is no CPU path.

    python examples/fast_tutorial_b200.py [--nwalkers 1024] [--nsteps 400]

What changes for a bilby user: `import bilby_b200 as bb` instead of `import bilby`, the sampler name, and - optionally -
`BatchedLikelihood.log_likelihood_from_unit_cube` for samplers that live in the unit hypercube on the device."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

import bilby_b200 as bb  # noqa: E402
from bilby_b200.core.prior import PriorDict, Uniform, PowerLaw, Sine, Cosine  # noqa: E402
from bilby_b200.core.sampler import BatchedLikelihood, run_sampler  # noqa: E402
from bilby_b200.gw.detector import InterferometerList  # noqa: E402
from bilby_b200.gw.source import lal_binary_black_hole  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nwalkers", type=int, default=1024)
    ap.add_argument("--nsteps", type=int, default=400)
    args = ap.parse_args()
    duration, sampling_frequency = 4.0, 2048.0
    injection = dict(mass_1=36.0, mass_2=29.0, a_1=0.4, a_2=0.3, tilt_1=0.0, tilt_2=0.0, phi_12=0.0, phi_jl=0.0,
                     luminosity_distance=2000.0, theta_jn=0.4, psi=2.659, phase=1.3, geocent_time=1126259642.413,
                     ra=1.375, dec=-1.2108)
    start = injection["geocent_time"] - 2.0
    waveform_arguments = dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0)
    waveform_generator = bb.gw.WaveformGenerator(
        duration=duration, sampling_frequency=sampling_frequency, start_time=start,
        frequency_domain_source_model=lal_binary_black_hole, waveform_arguments=waveform_arguments)
    ifos = InterferometerList(["H1", "L1"])
    ifos.set_strain_data_from_power_spectral_densities(sampling_frequency, duration, start,
                                                       rng=np.random.default_rng(88170235))
    ifos.inject_signal(parameters=injection, waveform_generator=waveform_generator)

    # the tutorial's priors: everything fixed to the injection except chirp mass, mass ratio, distance, phase, theta_jn
    mc = (36.0 * 29.0) ** 0.6 / 65.0 ** 0.2
    priors = PriorDict({k: v for k, v in injection.items() if k not in ("mass_1", "mass_2")})
    priors["chirp_mass"] = Uniform(mc - 3.0, mc + 3.0, "chirp_mass")
    priors["mass_ratio"] = Uniform(0.3, 1.0, "mass_ratio")
    priors["luminosity_distance"] = PowerLaw(2, 500.0, 5000.0, "luminosity_distance")
    priors["phase"] = Uniform(0, 2 * np.pi, "phase")
    priors["theta_jn"] = Sine(name="theta_jn")
    likelihood = bb.gw.GravitationalWaveTransient(ifos, waveform_generator, priors=priors, phase_marginalization=True,
                                                  distance_marginalization=True)
    print("one point, the reference's call:", likelihood.log_likelihood_ratio(
        dict(injection, chirp_mass=mc, mass_ratio=29.0 / 36.0)))

    t0 = time.time()
    result = run_sampler(likelihood, priors, sampler="b200_ensemble", nwalkers=args.nwalkers, nsteps=args.nsteps, seed=1)
    dt = time.time() - t0
    print(f"{result['num_likelihood_evaluations']} likelihood evaluations in {dt:.1f} s "
          f"({result['num_likelihood_evaluations'] / dt:.3g} per second), acceptance {result['acceptance_fraction']:.2f}")
    for j, key in enumerate(result["search_parameter_keys"]):
        lo, med, hi = np.percentile(result["samples"][:, j], [5, 50, 95])
        print(f"  {key:22s} {med:10.4f}  (+{hi - med:.4f} / -{med - lo:.4f})")

    # the same likelihood fed from unit-cube points that never leave the device (nested-sampler style)
    import torch
    batched = BatchedLikelihood(likelihood, priors)
    u = torch.rand((1_000_000, batched.ndim), dtype=torch.float64, device="cuda")
    batched.log_likelihood_from_unit_cube(u)          # first call at this size allocates the scratch buffers
    torch.cuda.synchronize()
    t0 = time.time()
    lnl = batched.log_likelihood_from_unit_cube(u)
    torch.cuda.synchronize()
    print(f"1e6 prior draws from the unit cube on the device: {1e6 / (time.time() - t0):.3g} evaluations per second, "
          f"max lnL {float(lnl.max()):.2f}")


if __name__ == "__main__":
    main()
