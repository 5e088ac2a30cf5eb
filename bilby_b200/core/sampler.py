"""Batched sampler adaptor.

The reference's samplers evaluate one point per call: ``Sampler.log_likelihood(theta)`` builds a parameter dict
and calls ``likelihood.log_likelihood_ratio`` (bilby/core/sampler/base_sampler.py:538-563); parallelism is a
``multiprocessing.Pool`` that pickles the whole likelihood into each worker (:772-800) and dynesty's ``pool.map``
over ``queue_size`` live-point proposals (bilby/core/sampler/dynesty.py:40-53, 750-753).  On a B200 one launch
evaluates 1e6 points, so the adaptor below hands *arrays of theta* to ``log_likelihood_ratio_batch``:

* ``BatchedLikelihood``  - theta [n, ndim] (numpy or CUDA tensor) in ``search_parameter_keys`` order + the fixed
  parameters of the prior  ->  lnL [n]; ``log_likelihood(theta)`` keeps the reference's one-point signature.
  ``rows_from_unit_cube_device`` / ``rows_from_theta_device`` / ``log_likelihood_from_unit_cube`` keep the points
  on the device: one kernel (csrc/bb_sampling.cuh) does ``PriorDict.rescale``, the waveform generator's parameter
  conversion and the packing into parameter rows, so theta never crosses PCIe.
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


    # ---- device-resident front end (include/bilby_b200.h: bb_set_sampling_priors / bb_rows_from_*_device) -----------
    _PRIOR_KINDS = {"DeltaFunction": 0, "Uniform": 1, "PowerLaw": 2, "LogUniform": 2, "Sine": 3, "Cosine": 4,
                    "Gaussian": 5, "Normal": 5}
    SOURCE_KEYS = ("mass_1", "mass_2", "chirp_mass", "mass_ratio", "total_mass", "symmetric_mass_ratio", "chi_1",
                   "chi_2", "a_1", "a_2", "tilt_1", "tilt_2", "cos_tilt_1", "cos_tilt_2", "luminosity_distance",
                   "theta_jn", "cos_theta_jn", "psi", "phase", "delta_phase", "ra", "dec", "geocent_time",
                   "time_jitter", "lambda_1", "lambda_2", "lambda_tilde", "delta_lambda_tilde")
    _IGNORED_FIXED = ("phi_jl", "phi_12")        # zero for aligned spins (conversion.py:255-257)

    @staticmethod
    def prior_spec(prior):
        """(kind, a, b, c) of include/bilby_b200.h bb_prior_spec for one analytic prior."""
        name = type(prior).__name__
        if name not in BatchedLikelihood._PRIOR_KINDS:
            raise NotImplementedError(f"prior class {name} has no device rescale (Uniform, PowerLaw, LogUniform, Sine, "
                                      "Cosine, Gaussian, DeltaFunction)")
        kind = BatchedLikelihood._PRIOR_KINDS[name]
        if kind == 0:
            return kind, float(prior.peak), 0.0, 0.0
        if kind == 2:
            return kind, float(prior.minimum), float(prior.maximum), float(prior.alpha)
        if kind == 5:
            return kind, float(prior.mu), float(prior.sigma), 0.0
        return kind, float(prior.minimum), float(prior.maximum), 0.0

    def _key_index(self, key):
        like = self.likelihood
        alias = {}
        if getattr(like, "reference_frame", "sky") != "sky":
            alias.update(azimuth="ra", zenith="dec")          # the rows carry (azimuth, zenith) in the sky columns
        if getattr(like, "time_reference", "geocent") != "geocent":
            alias[f"{like.time_reference}_time"] = "geocent_time"
        key = alias.get(key, key)
        if key not in self.SOURCE_KEYS:
            raise NotImplementedError(f"parameter '{key}' is not part of the device sampling front end")
        return self.SOURCE_KEYS.index(key)

    def _device_front_end(self):
        """Upload the prior table once (bb_set_sampling_priors)."""
        if getattr(self, "_front_end_ready", None) is self.likelihood.device_network:
            return self.likelihood.device_network
        import ctypes
        from .. import _lib
        net = self.likelihood.device_network
        if getattr(self.likelihood, "_cal_points", 0):
            raise NotImplementedError("calibration parameters are not part of the device sampling front end")

        class Spec(ctypes.Structure):
            _fields_ = [("kind", ctypes.c_int), ("key", ctypes.c_int), ("a", ctypes.c_double), ("b", ctypes.c_double),
                        ("c", ctypes.c_double)]
        specs = (Spec * max(1, self.ndim))()
        for j, key in enumerate(self.search_parameter_keys):
            kind, a, b, c = self.prior_spec(self.priors[key])
            specs[j] = Spec(kind, self._key_index(key), a, b, c)
        fixed = {k: v for k, v in self.fixed_parameters.items() if k not in self._IGNORED_FIXED}
        fkeys = (ctypes.c_int * max(1, len(fixed)))(*[self._key_index(k) for k in fixed])
        fvals = (ctypes.c_double * max(1, len(fixed)))(*[float(v) for v in fixed.values()])
        wfg = self.likelihood.waveform_generator
        conv = getattr(getattr(wfg, "parameter_conversion", None), "__name__", "")
        src = getattr(getattr(wfg, "frequency_domain_source_model", None), "__name__", "")
        bns = int("neutron_star" in conv or "neutron_star" in src)
        _lib.check(net.lib.bb_set_sampling_priors(net.ptr, self.ndim, ctypes.cast(specs, ctypes.c_void_p), len(fixed),
                                                  ctypes.cast(fkeys, ctypes.c_void_p),
                                                  ctypes.cast(fvals, ctypes.c_void_p), bns))
        self._front_end_ready = net
        return net

    def _check_points(self, x):
        net = self._device_front_end()
        torch = net.torch
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64 and x.dim() == 2
                and x.shape[1] == self.ndim):
            raise ValueError(f"points must be a CUDA float64 tensor of shape [n, {self.ndim}]")
        return net, torch, x.contiguous()

    def rows_from_unit_cube_device(self, u, return_theta=False):
        """Unit hypercube [n, ndim] (CUDA tensor) -> parameter rows [n, 16] (CUDA) in one kernel: PriorDict.rescale
        (core/prior/dict.py:647-666) + the generator's parameter conversion + packing."""
        from .. import _lib
        net, torch, u = self._check_points(u)
        rows = torch.empty((u.shape[0], 16), dtype=torch.float64, device=u.device)
        theta = torch.empty_like(u) if return_theta else None
        _lib.check(net.lib.bb_rows_from_unit_cube_device(net.ptr, u.data_ptr(), u.shape[0],
                                                         theta.data_ptr() if return_theta else None, rows.data_ptr(),
                                                         net._stream()))
        return (rows, theta) if return_theta else rows

    def rows_from_theta_device(self, theta):
        """Sampled parameters [n, ndim] (CUDA tensor, search_parameter_keys order) -> parameter rows [n, 16]."""
        from .. import _lib
        net, torch, theta = self._check_points(theta)
        rows = torch.empty((theta.shape[0], 16), dtype=torch.float64, device=theta.device)
        _lib.check(net.lib.bb_rows_from_theta_device(net.ptr, theta.data_ptr(), theta.shape[0], rows.data_ptr(),
                                                     net._stream()))
        return rows

    def log_likelihood_from_unit_cube(self, u):
        """CUDA unit-cube points in, CUDA lnL out; nothing but the kernels' launches touches the host."""
        lnl = self.likelihood.log_likelihood_ratio_batch(self.rows_from_unit_cube_device(u))
        return lnl if self.use_ratio else lnl + self.likelihood.noise_log_likelihood()

    def log_likelihood_from_theta_device(self, theta):
        lnl = self.likelihood.log_likelihood_ratio_batch(self.rows_from_theta_device(theta))
        return lnl if self.use_ratio else lnl + self.likelihood.noise_log_likelihood()


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
