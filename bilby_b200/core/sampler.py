"""Batched sampler adaptor.

The reference's samplers evaluate one point per call: ``Sampler.log_likelihood(theta)`` builds a parameter dict
and calls ``likelihood.log_likelihood_ratio`` (bilby/core/sampler/base_sampler.py:538-563); parallelism is a
``multiprocessing.Pool`` that pickles the whole likelihood into each worker (:772-800) and dynesty's ``pool.map``
over ``queue_size`` live-point proposals (bilby/core/sampler/dynesty.py:40-53, 750-753).  On a B200 one launch
evaluates 1e6 points, so the adaptor below hands *arrays of theta* to ``log_likelihood_ratio_batch``:

* ``BatchedLikelihood``  - theta [n, ndim] (numpy or CUDA tensor) in ``search_parameter_keys`` order + the fixed
  parameters of the prior  ->  lnL [n]; ``log_likelihood(theta)`` keeps the reference's one-point signature.
* ``DeviceBatchPool``    - a ``pool`` object for samplers that fan proposals out with ``pool.map(fn, points)``
  (dynesty's ``pool=``/``queue_size=``): the whole iterable is one batch, ``fn`` is never called.
"""
import numpy as np

from .prior import Prior


class BatchedLikelihood:
    def __init__(self, likelihood, priors, use_ratio=True):
        self.likelihood = likelihood
        self.use_ratio = use_ratio
        self.priors = priors
        # base_sampler.py:297-325 _initialise_parameters: sampled = Prior objects that are not fixed
        self.search_parameter_keys = [k for k, p in priors.items() if isinstance(p, Prior)
                                      and not getattr(p, "is_fixed", False)]
        self.fixed_parameters = {}
        for k, p in priors.items():
            if k in self.search_parameter_keys:
                continue
            self.fixed_parameters[k] = p.peak if isinstance(p, Prior) else p
        self.ndim = len(self.search_parameter_keys)

    def parameters_from_theta(self, theta):
        """theta [n, ndim] -> dict of length-n arrays (+ scalars for the fixed parameters)."""
        if theta.ndim != 2 or theta.shape[1] != self.ndim:
            raise ValueError(f"theta must have shape [n, {self.ndim}]")
        params = dict(self.fixed_parameters)
        for j, key in enumerate(self.search_parameter_keys):
            params[key] = theta[:, j]
        return params

    def log_likelihood_batch(self, theta):
        """numpy [n, ndim] -> numpy [n]; CUDA tensor [n, ndim] -> CUDA tensor [n] (stays on the device)."""
        if isinstance(theta, np.ndarray):
            theta = np.ascontiguousarray(theta, dtype=np.float64)
        lnl = self.likelihood.log_likelihood_ratio_batch(self.parameters_from_theta(theta))
        if not self.use_ratio:
            lnl = lnl + self.likelihood.noise_log_likelihood()
        return lnl

    def log_likelihood(self, theta):
        """One point, the reference's signature (base_sampler.py:538-563)."""
        return float(self.log_likelihood_batch(np.asarray(theta, dtype=np.float64)[None, :])[0])

    def prior_transform_batch(self, u):
        """Unit hypercube [n, ndim] -> theta [n, ndim] (base_sampler.py:495-510 prior_transform, vectorised)."""
        u = np.asarray(u, dtype=np.float64)
        return np.stack([self.priors[k].rescale(u[:, j]) for j, k in enumerate(self.search_parameter_keys)], axis=1)


class DeviceBatchPool:
    """``pool``-like object: ``map(fn, points)`` evaluates every point of the iterable in one device batch."""

    def __init__(self, batched_likelihood, queue_size=4096):
        self.batched = batched_likelihood
        self.size = queue_size          # dynesty reads pool.size for queue_size when it is not given

    def map(self, fn, points):
        pts = np.asarray(list(points), dtype=np.float64)
        if pts.size == 0:
            return []
        return list(self.batched.log_likelihood_batch(pts))

    def close(self):
        pass

    def join(self):
        pass
