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
