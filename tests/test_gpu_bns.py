"""GPU parity for the long-signal path (BASELINE.json configs[3]): TaylorF2 + tides, 128 s @ 4096 Hz, H1L1V1,
262145 bins per detector, and the frequency-sharded evaluation (partial inner products summed, SURVEY.md
section 8e) - on one GPU by emulating the shards, and over NCCL when two GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
T_INJ = 1126259642.413
DURATION, FS = 128.0, 4096.0
START = T_INJ - DURATION + 2
WA = dict(waveform_approximant="TaylorF2", reference_frequency=50.0, minimum_frequency=20.0)
INJ = dict(mass_1=1.5, mass_2=1.3, chi_1=0.02, chi_2=0.01, luminosity_distance=100.0, theta_jn=0.4, psi=2.659,
           phase=1.3, geocent_time=T_INJ, ra=1.375, dec=-1.2108, lambda_1=400.0, lambda_2=600.0)


def bns_draws(n, rng):
    mc = rng.uniform(1.15, 1.25, n)
    q = rng.uniform(0.5, 1.0, n)
    return dict(chirp_mass=mc, mass_ratio=q, chi_1=rng.uniform(-0.05, 0.05, n), chi_2=rng.uniform(-0.05, 0.05, n),
                luminosity_distance=rng.uniform(10, 500, n), theta_jn=np.arccos(rng.uniform(-1, 1, n)),
                psi=rng.uniform(0, np.pi, n), phase=rng.uniform(0, 2 * np.pi, n), ra=rng.uniform(0, 2 * np.pi, n),
                dec=np.arcsin(rng.uniform(-1, 1, n)), geocent_time=rng.uniform(T_INJ - 0.1, T_INJ + 0.1, n),
                lambda_1=rng.uniform(0, 5000, n), lambda_2=rng.uniform(0, 5000, n))


def build_pair(noise_seed=3, **kw):
    """(oracle likelihood, product likelihood) on identical data."""
    import bilby_b200 as bb
    from bilby_b200.gw.conversion import convert_to_lal_binary_neutron_star_parameters
    from bilby_b200.gw.detector import InterferometerList
    from bilby_b200.gw.source import lal_binary_neutron_star
    from oracle import cbc_likelihood as ocl
    oifos = [ocl.OracleInterferometer(n, FS, DURATION, START) for n in ("H1", "L1", "V1")]
    rng = np.random.default_rng(noise_seed)
    conv = ocl.convert_to_lal_binary_black_hole_parameters(INJ)
    pols = ocl.lal_binary_neutron_star(oifos[0].frequency_array, *[conv[k] for k in ocl.SOURCE_ARGS],
                                       INJ["lambda_1"], INJ["lambda_2"], **WA)
    for o in oifos:
        o.set_gaussian_noise(rng)
        o.frequency_domain_strain = o.frequency_domain_strain + o.get_detector_response(pols, conv)
    olike = ocl.OracleLikelihood(oifos, source_model=ocl.lal_binary_neutron_star, waveform_arguments=WA, **kw)
    ifos = InterferometerList(["H1", "L1", "V1"])
    for ifo, o in zip(ifos, oifos):
        ifo.minimum_frequency = 20.0
        ifo.maximum_frequency = FS / 2
        ifo.set_strain_data_from_frequency_domain_strain(o.frequency_domain_strain, sampling_frequency=FS,
                                                         duration=DURATION, start_time=START)
    wfg = bb.gw.WaveformGenerator(duration=DURATION, sampling_frequency=FS,
                                  frequency_domain_source_model=lal_binary_neutron_star,
                                  parameter_conversion=convert_to_lal_binary_neutron_star_parameters,
                                  waveform_arguments=WA)
    return olike, ifos, wfg


def test_taylorf2_tides_128s_vs_oracle():
    import bilby_b200 as bb
    olike, ifos, wfg = build_pair()
    like = bb.gw.GravitationalWaveTransient(ifos, wfg)
    n = 6
    draws = bns_draws(n, np.random.default_rng(20261017))
    draws = {k: np.concatenate([v, [INJ[k] if k in INJ else (1.5 * 1.3) ** 0.6 / 2.8 ** 0.2 if k == "chirp_mass"
                                    else 1.3 / 1.5]]) for k, v in draws.items()}
    got = like.log_likelihood_ratio_batch(draws)
    snr = like.inner_products_batch(__import__("torch").from_numpy(like.pack(draws)).cuda()).cpu().numpy()
    for i in range(n + 1):
        p = {k: float(v[i]) for k, v in draws.items()}
        ref = olike.log_likelihood_ratio(p)
        per_det = olike.log_likelihood_ratio(p, return_snrs=True)
        scale = max(abs(ref), 0.5 * sum(h for _, h in per_det))
        assert abs(got[i] - ref) < 1e-8 * scale, (i, got[i], ref)
        for d, (dh, hh) in enumerate(per_det):
            assert abs(complex(snr[i, d, 0], snr[i, d, 1]) - dh) < 1e-8 * hh
            assert abs(snr[i, d, 2] - hh) < 1e-9 * hh


def test_frequency_shards_sum_to_full_on_one_gpu():
    import torch
    import bilby_b200 as bb
    from bilby_b200 import _lib
    from bilby_b200.parallel import frequency_shards
    _, ifos, wfg = build_pair()
    like = bb.gw.GravitationalWaveTransient(ifos, wfg)
    draws = bns_draws(64, np.random.default_rng(1))
    rows = torch.from_numpy(like.pack(draws)).cuda()
    full = like.inner_products_batch(rows).clone()
    lnl_full = like.log_likelihood_ratio_batch(rows).cpu().numpy()
    net = like.device_network
    total = torch.zeros_like(full)
    for b, e in frequency_shards(2560, 262144, 8, net.n_freq):
        _lib.check(net.lib.bb_set_frequency_shard(net.ptr, b, e))
        total += like.inner_products_batch(rows)
    _lib.check(net.lib.bb_set_frequency_shard(net.ptr, 0, net.n_freq))
    scale = full[..., 2].abs().max().item()
    assert (total - full).abs().max().item() < 1e-11 * scale
    lnl = like.likelihood_from_inner_products(rows, total).cpu().numpy()
    assert np.max(np.abs(lnl - lnl_full)) < 1e-10 * np.max(np.abs(lnl_full))


def test_frequency_sharding_over_nccl_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "dist_freq_shard_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "FREQ_SHARD_OK" in res.stdout
