"""The batched ensemble sampler (bilby_b200/core/sampler.py B200Ensemble; bilby plugin group "bilby.samplers",
docs/plugins.txt:27-48) on an analytic likelihood - host logic only, no GPU: the sampler must recover a known
Gaussian posterior, hand ARRAYS of points to the likelihood, respect the prior bounds and be registered as a plugin."""
import os
import re

import numpy as np

from bilby_b200.core.prior import PriorDict, Uniform
from bilby_b200.core.sampler import B200Ensemble, BatchedLikelihood, run_sampler

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


class GaussianBatch:
    """lnL = -1/2 sum ((x - mu) / sigma)^2 evaluated for a whole dict of arrays at once."""
    mu = np.array([0.3, -1.2, 2.0])
    sigma = np.array([0.1, 0.25, 0.05])

    def __init__(self):
        self.calls, self.points = 0, 0

    def log_likelihood_ratio_batch(self, p):
        x = np.stack([p["a"], p["b"], p["c"] + 0 * p["a"]], axis=1)
        self.calls += 1
        self.points += len(x)
        return -0.5 * np.sum(((x - self.mu) / self.sigma) ** 2, axis=1)

    def noise_log_likelihood(self):
        return 0.0


def test_ensemble_sampler_recovers_gaussian_posterior_with_batched_calls():
    like = GaussianBatch()
    priors = PriorDict(dict(a=Uniform(-2, 2, "a"), b=Uniform(-3, 3, "b"), c=2.0))
    res = run_sampler(like, priors, nwalkers=512, nsteps=400, seed=3, record_visited=True)
    assert res["search_parameter_keys"] == ["a", "b"]
    s = res["samples"]
    assert s.shape == (200 * 512, 2)
    assert np.all(np.abs(s[:, 0]) <= 2) and np.all(np.abs(s[:, 1]) <= 3)
    assert np.allclose(s.mean(axis=0), like.mu[:2], atol=4 * like.sigma[:2] / np.sqrt(2000))
    assert np.allclose(s.std(axis=0), like.sigma[:2], rtol=0.06)
    assert 0.2 < res["acceptance_fraction"] < 0.9
    # one likelihood call per half-ensemble move (+ the initial ensemble): batches, never single points
    assert like.calls == 1 + 2 * 400
    assert like.points == res["num_likelihood_evaluations"] == len(res["visited_log_likelihood"])
    # the recorded lnL of the visited points are what the likelihood returns for them
    v = res["visited_theta"]
    again = -0.5 * (((v[:, 0] - 0.3) / 0.1) ** 2 + ((v[:, 1] + 1.2) / 0.25) ** 2)
    assert np.allclose(again, res["visited_log_likelihood"], rtol=0, atol=1e-9)


def test_one_point_signature_and_plugin_registration():
    like = GaussianBatch()
    priors = PriorDict(dict(a=Uniform(-2, 2, "a"), b=Uniform(-3, 3, "b"), c=2.0))
    sampler = B200Ensemble(like, priors, nwalkers=8, nsteps=2)
    assert np.isclose(sampler.log_likelihood([0.3, -1.2]), 0.0)
    assert isinstance(sampler.batched, BatchedLikelihood)
    text = open(os.path.join(ROOT, "pyproject.toml")).read()
    m = re.search(r'\[project\.entry-points\."bilby\.samplers"\]\s*\n"b200_ensemble"\s*=\s*"([\w.]+):(\w+)"', text)
    assert m, "entry point missing"
    mod = __import__(m.group(1), fromlist=[m.group(2)])
    assert getattr(mod, m.group(2)) is B200Ensemble


def test_host_prior_rescale_and_conversion_match_the_reference_goldens():
    """The host mirror (core/prior.py rescale, gw/conversion.py, gw/_params.py pack_rows) against vectors generated
    by the UNMODIFIED reference (oracle/tools/make_golden_prior.py); the device kernel is tested against the same
    file in tests/test_gpu_sampling_front_end.py."""
    from bilby_b200.core import prior as P
    from bilby_b200.gw import _params, conversion
    g = np.load(os.path.join(ROOT, "tests", "golden", "prior_transform.npz"))
    make = {0: lambda a, b, c: P.DeltaFunction(a), 1: lambda a, b, c: P.Uniform(a, b),
            2: lambda a, b, c: P.PowerLaw(c, a, b), 3: lambda a, b, c: P.Sine(a, b), 4: lambda a, b, c: P.Cosine(a, b),
            5: lambda a, b, c: P.Gaussian(a, b)}
    for case in (str(c) for c in g["case_names"]):
        keys = [str(k) for k in g[f"{case}_keys"]]
        spec = g[f"{case}_spec"]
        u, ref_theta, ref_rows = g[f"{case}_unit"], g[f"{case}_theta"], g[f"{case}_rows"]
        n = len(u)
        params = {str(k): float(v) for k, v in zip(g[f"{case}_fixed_keys"], g[f"{case}_fixed_values"])}
        for j, k in enumerate(keys):
            pr = make[int(spec[j, 0])](*spec[j, 1:])
            params[k] = pr.rescale(u[:, j])
            np.testing.assert_allclose(params[k][2:], ref_theta[2:, j], rtol=1e-13, atol=0, err_msg=f"{case} {k}")
        convert = (conversion.convert_to_lal_binary_neutron_star_parameters if bool(g[f"{case}_bns"])
                   else conversion.convert_to_lal_binary_black_hole_parameters)
        converted, _ = convert(params)
        rows = _params.pack_rows(converted, n, np)
        for c, name in enumerate(str(k) for k in g["row_keys"]):
            np.testing.assert_allclose(rows[2:, c], ref_rows[2:, c], rtol=1e-12, atol=1e-15, err_msg=f"{case} {name}")


def test_device_front_end_tables():
    from bilby_b200.core import prior as P
    assert BatchedLikelihood.prior_spec(P.PowerLaw(2, 10.0, 500.0)) == (2, 10.0, 500.0, 2.0)
    assert BatchedLikelihood.prior_spec(P.Gaussian(1.0, 0.5)) == (5, 1.0, 0.5, 0.0)
    assert BatchedLikelihood.prior_spec(P.DeltaFunction(3.0)) == (0, 3.0, 0.0, 0.0)
    assert len(BatchedLikelihood.SOURCE_KEYS) == 28 and BatchedLikelihood.SOURCE_KEYS.index("geocent_time") == 22
    text = open(os.path.join(ROOT, "include", "bilby_b200.h")).read()
    for i, key in enumerate(BatchedLikelihood.SOURCE_KEYS):          # the header's enum is the same list
        assert re.search(rf"BB_KEY_{key.upper()} = {i}\b", text), key
