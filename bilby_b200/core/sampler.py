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


# ------------------------------------------------------------------------------------------------------------------
# A batched sampler behind bilby's sampler-plugin interface (docs/plugins.txt:27-48, pyproject.toml:37-42 group
# "bilby.samplers"; base class bilby/core/sampler/base_sampler.py:110 Sampler / :908 MCMCSampler)
# ------------------------------------------------------------------------------------------------------------------
try:      # with bilby installed the class is a genuine plugin (entry point "b200_ensemble" in pyproject.toml)
    from bilby.core.sampler.base_sampler import MCMCSampler as _SamplerBase      # pragma: no cover
except Exception:      # bilby is not a dependency of this package: same constructor contract, nothing else
    class _SamplerBase:
        """The part of bilby.core.sampler.base_sampler.Sampler.__init__ (:110-260) the plugin relies on."""
        default_kwargs = {}

        def __init__(self, likelihood, priors, outdir="outdir", label="label", use_ratio=False, plot=False,
                     skip_import_verification=True, injection_parameters=None, meta_data=None, result_class=None,
                     likelihood_benchmark=False, soft_init=False, exit_code=130, npool=1, **kwargs):
            self.likelihood, self.priors = likelihood, priors
            self.outdir, self.label, self.use_ratio = outdir, label, use_ratio
            self.injection_parameters, self.meta_data = injection_parameters, meta_data
            self.kwargs = dict(self.default_kwargs)
            self.kwargs.update(kwargs)


class B200Ensemble(_SamplerBase):
    """Affine-invariant ensemble MCMC (Goodman & Weare stretch move, the algorithm of the reference's `emcee` wrapper,
    bilby/core/sampler/emcee.py) written for a likelihood that evaluates ARRAYS of points: every iteration moves half
    of the walkers with ONE call of ``log_likelihood_ratio_batch`` instead of nwalkers / 2 calls of
    ``log_likelihood(theta)`` (base_sampler.py:538-563) fanned out over a process pool (:772-800).

    kwargs: nwalkers (default 2048), nsteps (500), nburn (nsteps // 2), a (stretch scale, 2.0), seed, thin (1),
    record_visited (keep every evaluated point and its lnL: the parity tests replay them through the oracle).

    ``run_sampler()`` returns a dict with the reference Result's sampler fields: ``samples`` [n, ndim] after burn-in,
    ``search_parameter_keys``, ``log_likelihood_evaluations``, ``log_prior_evaluations``, ``sampling_time``,
    ``num_likelihood_evaluations``, ``walkers`` [nsteps, nwalkers, ndim], ``nburn``, ``acceptance_fraction``.
    """
    sampler_name = "b200_ensemble"
    default_kwargs = dict(nwalkers=2048, nsteps=500, nburn=None, a=2.0, seed=None, thin=1, record_visited=False)

    def __init__(self, likelihood, priors, **kwargs):
        super().__init__(likelihood, priors, **kwargs)
        for k, v in self.default_kwargs.items():
            self.kwargs.setdefault(k, v)
        self.batched = BatchedLikelihood(likelihood, priors, use_ratio=True)
        self.search_parameter_keys = self.batched.search_parameter_keys
        self.ndim = self.batched.ndim

    # ---- vectorised prior (base_sampler.py:522-536 log_prior, one row per walker)
    def log_prior_batch(self, theta):
        lp = np.zeros(len(theta))
        for j, key in enumerate(self.search_parameter_keys):
            lp = lp + self.priors[key].ln_prob(theta[:, j])
        return lp

    def log_likelihood(self, theta):
        """One point with the reference's signature."""
        return self.batched.log_likelihood(theta)

    def _lnpost(self, theta, visited):
        lp = self.log_prior_batch(theta)
        ok = np.isfinite(lp)
        lnl = np.full(len(theta), -np.inf)
        if ok.any():
            lnl[ok] = self.batched.log_likelihood_batch(theta[ok])
            if visited is not None:
                visited.append((theta[ok].copy(), lnl[ok].copy()))
        return lp, lnl

    def run_sampler(self):
        import time
        kw = self.kwargs
        rng = np.random.default_rng(kw["seed"])
        nw, ns, a = int(kw["nwalkers"]), int(kw["nsteps"]), float(kw["a"])
        if nw % 2 or nw < 2 * self.ndim:
            raise ValueError("nwalkers must be even and at least 2 * ndim")
        visited = [] if kw["record_visited"] else None
        t0 = time.time()
        # start from prior draws (base_sampler.py:447-476 get_initial_points_from_prior, vectorised)
        theta = self.batched.prior_transform_batch(rng.uniform(0, 1, (nw, self.ndim)))
        lp, lnl = self._lnpost(theta, visited)
        n_eval = nw
        chain = np.empty((ns, nw, self.ndim))
        chain_lnl = np.empty((ns, nw))
        chain_lp = np.empty((ns, nw))
        accepted = 0
        half = nw // 2
        for it in range(ns):
            for first in (True, False):
                mov = slice(0, half) if first else slice(half, nw)
                oth = slice(half, nw) if first else slice(0, half)
                # stretch move: z ~ g(z) with g(z) ~ 1 / sqrt(z) on [1 / a, a]
                z = ((a - 1.0) * rng.uniform(0, 1, half) + 1.0) ** 2 / a
                partner = theta[oth][rng.integers(0, half, half)]
                prop = partner + z[:, None] * (theta[mov] - partner)
                plp, plnl = self._lnpost(prop, visited)
                n_eval += int(np.isfinite(plp).sum())
                with np.errstate(invalid="ignore"):
                    lnr = (self.ndim - 1) * np.log(z) + (plp + plnl) - (lp[mov] + lnl[mov])
                acc = np.log(rng.uniform(0, 1, half)) < lnr
                acc &= np.isfinite(plp) & np.isfinite(plnl)
                idx = np.arange(mov.start, mov.stop)[acc]
                theta[idx], lp[idx], lnl[idx] = prop[acc], plp[acc], plnl[acc]
                accepted += int(acc.sum())
            chain[it], chain_lnl[it], chain_lp[it] = theta, lnl, lp
        nburn = ns // 2 if kw["nburn"] is None else int(kw["nburn"])
        thin = max(1, int(kw["thin"]))
        out = dict(sampler=self.sampler_name, search_parameter_keys=list(self.search_parameter_keys),
                   samples=chain[nburn::thin].reshape(-1, self.ndim),
                   log_likelihood_evaluations=chain_lnl[nburn::thin].ravel(),
                   log_prior_evaluations=chain_lp[nburn::thin].ravel(), walkers=chain, nburn=nburn,
                   num_likelihood_evaluations=n_eval, sampling_time=time.time() - t0,
                   acceptance_fraction=accepted / (ns * nw))
        if visited is not None:
            out["visited_theta"] = np.concatenate([v[0] for v in visited])
            out["visited_log_likelihood"] = np.concatenate([v[1] for v in visited])
        self.result = out
        return out


def run_sampler(likelihood, priors, sampler="b200_ensemble", **kwargs):
    """The part of bilby.core.sampler.run_sampler (bilby/core/sampler/__init__.py) that picks the sampler class by
    name and runs it; only the batched sampler of this package is known here."""
    if sampler not in ("b200_ensemble", "bilby_b200.b200_ensemble"):
        raise ValueError(f"Sampler {sampler!r} is not implemented here: bilby's own samplers call log_likelihood(theta) "
                         "one point at a time and work unchanged with these likelihood classes")
    return B200Ensemble(likelihood, priors, **kwargs).run_sampler()
