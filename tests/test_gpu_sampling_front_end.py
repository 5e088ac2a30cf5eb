"""Device-resident sampling front end (csrc/bb_sampling.cuh): one kernel does
PriorDict.rescale + the waveform generator's parameter conversion + the packing into parameter rows.

Golden vectors come from the UNMODIFIED reference (oracle/tools/make_golden_prior.py: bilby.core.prior.PriorDict.rescale,
bilby.gw.conversion.convert_to_lal_binary_{black_hole,neutron_star}_parameters); eight parameterisations."""
import ctypes
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(GOLDEN_DIR, "prior_transform.npz"))
CASES = [str(c) for c in G["case_names"]]


class Spec(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("key", ctypes.c_int), ("a", ctypes.c_double), ("b", ctypes.c_double),
                ("c", ctypes.c_double)]


def _setup(handle, case):
    from bilby_b200 import _lib
    from bilby_b200.core.sampler import BatchedLikelihood
    keys = [str(k) for k in G[f"{case}_keys"]]
    spec = G[f"{case}_spec"]
    fkeys = [str(k) for k in G[f"{case}_fixed_keys"]]
    fvals = G[f"{case}_fixed_values"]
    index = BatchedLikelihood.SOURCE_KEYS.index
    specs = (Spec * len(keys))(*[Spec(int(spec[j, 0]), index(k), spec[j, 1], spec[j, 2], spec[j, 3])
                                 for j, k in enumerate(keys)])
    fk = (ctypes.c_int * max(1, len(fkeys)))(*[index(k) for k in fkeys])
    fv = (ctypes.c_double * max(1, len(fkeys)))(*[float(v) for v in fvals])
    _lib.check(handle.lib.bb_set_sampling_priors(handle.ptr, len(keys), ctypes.cast(specs, ctypes.c_void_p), len(fkeys),
                                                 ctypes.cast(fk, ctypes.c_void_p), ctypes.cast(fv, ctypes.c_void_p),
                                                 int(bool(G[f"{case}_bns"]))))
    return keys


def _close(got, ref, rtol, what):
    scale = np.maximum(np.abs(ref), 1e-300)
    err = np.abs(got - ref) / scale
    err[(ref == 0) & (np.abs(got) < 1e-15)] = 0.0
    assert err.max() < rtol, (what, err.max(), np.unravel_index(err.argmax(), err.shape))


@pytest.mark.parametrize("case", CASES)
def test_unit_cube_to_rows_vs_reference(case):
    import torch
    from bilby_b200 import _lib
    h = _lib.Handle()
    keys = _setup(h, case)
    u = torch.from_numpy(G[f"{case}_unit"]).cuda()
    n = u.shape[0]
    theta = torch.empty_like(u)
    rows = torch.empty((n, 16), dtype=torch.float64, device="cuda")
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(h.lib.bb_rows_from_unit_cube_device(h.ptr, u.data_ptr(), n, theta.data_ptr(), rows.data_ptr(), stream))
    torch.cuda.synchronize()
    # sampled parameters: transcendental functions of the two libraries agree to a few ulp; the phase wrap and the
    # arccos near +-1 amplify that (acos'(x) ~ 1/sqrt(1 - x^2))
    ref_theta = G[f"{case}_theta"]
    got_theta = theta.cpu().numpy()
    for j, k in enumerate(keys):
        gaussian = int(G[f"{case}_spec"][j, 0]) == 5
        _close(got_theta[2:, j], ref_theta[2:, j], 1e-9 if gaussian else 1e-12, (case, k))
        _close(got_theta[:2, j], ref_theta[:2, j], 1e-6, (case, k, "ends of the unit interval"))
    ref_rows = G[f"{case}_rows"]
    got_rows = rows.cpu().numpy()
    assert np.all(got_rows[:, 14:] == 0.0)
    for c, name in enumerate(str(k) for k in G["row_keys"]):
        tol = 1e-9 if name in ("phase", "theta_jn", "dec", "lambda_1", "lambda_2", "chi_2") else 1e-12
        _close(got_rows[2:, c], ref_rows[2:, c], tol, (case, name))
    # the same rows from the sampled parameters (MCMC samplers hold theta, not the unit cube)
    rows2 = torch.empty_like(rows)
    th = torch.from_numpy(ref_theta).cuda()
    _lib.check(h.lib.bb_rows_from_theta_device(h.ptr, th.data_ptr(), n, rows2.data_ptr(), stream))
    torch.cuda.synchronize()
    for c, name in enumerate(str(k) for k in G["row_keys"]):
        tol = 1e-9 if name in ("lambda_1", "lambda_2") else 1e-13
        _close(rows2.cpu().numpy()[:, c], ref_rows[:, c], tol, (case, name, "from theta"))


def test_front_end_refuses_bad_tables():
    from bilby_b200 import _lib
    h = _lib.Handle()
    specs = (Spec * 2)(Spec(1, 2, 0.0, 1.0, 0.0), Spec(1, 2, 0.0, 1.0, 0.0))        # chirp_mass twice
    assert h.lib.bb_set_sampling_priors(h.ptr, 2, ctypes.cast(specs, ctypes.c_void_p), 0, None, None, 0) != 0
    specs = (Spec * 1)(Spec(9, 2, 0.0, 1.0, 0.0))                                    # unknown prior class
    assert h.lib.bb_set_sampling_priors(h.ptr, 1, ctypes.cast(specs, ctypes.c_void_p), 0, None, None, 0) != 0
    assert h.lib.bb_rows_from_theta_device(_lib.Handle().ptr, None, 4, None, None) != 0   # priors not set


def test_likelihood_from_unit_cube_stays_on_device_and_matches_host_path():
    """BatchedLikelihood.log_likelihood_from_unit_cube (device: rescale + conversion + rows + kernels) against the
    host route (numpy rescale -> dict -> conversion -> rows -> kernels) on a distance + phase marginalised likelihood
    with a detector-frame time prior side effect."""
    import torch
    import bilby_b200 as bb
    from bilby_b200.core.prior import PriorDict, Uniform, PowerLaw, Sine, Cosine
    from bilby_b200.core.sampler import BatchedLikelihood
    from bilby_b200.gw.detector import InterferometerList
    from bilby_b200.gw.source import lal_binary_black_hole
    from bilby_b200.workloads import INJECTION
    inj = dict(INJECTION)
    start = inj["geocent_time"] - 2.0
    wa = dict(waveform_approximant="IMRPhenomD", reference_frequency=50.0, minimum_frequency=20.0)
    wfg = bb.gw.WaveformGenerator(duration=4.0, sampling_frequency=2048.0, start_time=start,
                                  frequency_domain_source_model=lal_binary_black_hole, waveform_arguments=dict(wa))
    ifos = InterferometerList(["H1", "L1"])
    ifos.set_strain_data_from_zero_noise(2048.0, 4.0, start)
    ifos.inject_signal(parameters=inj, waveform_generator=wfg)
    priors = PriorDict(dict(
        chirp_mass=Uniform(25.0, 32.0, "chirp_mass"), mass_ratio=Uniform(0.3, 1.0, "mass_ratio"),
        chi_1=Uniform(-0.8, 0.8, "chi_1"), chi_2=Uniform(-0.8, 0.8, "chi_2"),
        luminosity_distance=PowerLaw(2, 500.0, 5000.0, "luminosity_distance"), theta_jn=Sine(name="theta_jn"),
        psi=Uniform(0, np.pi, "psi"), phase=Uniform(0, 2 * np.pi, "phase"), ra=Uniform(0, 2 * np.pi, "ra"),
        dec=Cosine(name="dec"), geocent_time=Uniform(inj["geocent_time"] - 0.05, inj["geocent_time"] + 0.05, "geocent_time")))
    like = bb.gw.GravitationalWaveTransient(ifos, wfg, phase_marginalization=True, distance_marginalization=True,
                                            priors=priors)
    batched = BatchedLikelihood(like, priors)
    assert "phase" not in batched.search_parameter_keys and "luminosity_distance" not in batched.search_parameter_keys
    n = 4096
    u_np = np.random.default_rng(5).uniform(0, 1, (n, batched.ndim))
    u = torch.from_numpy(u_np).cuda()
    lnl_dev = batched.log_likelihood_from_unit_cube(u)
    assert lnl_dev.is_cuda
    lnl_host = batched.log_likelihood_batch(batched.prior_transform_batch(u_np))
    got = lnl_dev.cpu().numpy()
    scale = np.maximum(np.abs(lnl_host), 1.0)
    assert np.max(np.abs(got - lnl_host) / scale) < 1e-9
    rows, theta = batched.rows_from_unit_cube_device(u, return_theta=True)
    lnl_theta = batched.log_likelihood_from_theta_device(theta).cpu().numpy()
    assert np.array_equal(lnl_theta, got)
