"""The few analytic priors the marginalised likelihoods need at set-up
(bilby/core/prior/analytical.py: Uniform :186-242, PowerLaw :68-147, Sine/Cosine, DeltaFunction;
bilby/core/prior/dict.py PriorDict - dict semantics + sample()).  Host-side only."""
import numpy as np


class Prior:
    is_fixed = False

    def __init__(self, name=None, latex_label=None, minimum=-np.inf, maximum=np.inf, boundary=None, unit=None):
        self.name = name
        self.latex_label = latex_label
        self.minimum = minimum
        self.maximum = maximum
        self.boundary = boundary
        self.unit = unit

    def sample(self, size=None, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        return self.rescale(rng.uniform(0, 1, size))

    def ln_prob(self, val):
        with np.errstate(divide="ignore"):
            return np.log(self.prob(val))

    def __repr__(self):
        return f"{self.__class__.__name__}(minimum={self.minimum}, maximum={self.maximum}, name={self.name!r})"


class DeltaFunction(Prior):
    is_fixed = True

    def __init__(self, peak, name=None, latex_label=None, unit=None):
        super().__init__(name=name, latex_label=latex_label, minimum=peak, maximum=peak, unit=unit)
        self.peak = peak

    def rescale(self, val):
        return self.peak * np.asarray(val) ** 0

    def prob(self, val):
        return np.where(np.asarray(val) == self.peak, np.inf, 0.0)


class Uniform(Prior):
    def __init__(self, minimum, maximum, name=None, latex_label=None, unit=None, boundary=None):
        super().__init__(name=name, latex_label=latex_label, minimum=minimum, maximum=maximum,
                         boundary=boundary, unit=unit)

    def rescale(self, val):
        return self.minimum + val * (self.maximum - self.minimum)

    def prob(self, val):
        val = np.asarray(val)
        return ((val >= self.minimum) & (val <= self.maximum)) / (self.maximum - self.minimum)


class PowerLaw(Prior):
    def __init__(self, alpha, minimum, maximum, name=None, latex_label=None, unit=None, boundary=None):
        super().__init__(name=name, latex_label=latex_label, minimum=minimum, maximum=maximum,
                         boundary=boundary, unit=unit)
        self.alpha = alpha

    def rescale(self, val):
        if self.alpha == -1:
            return self.minimum * np.exp(val * np.log(self.maximum / self.minimum))
        return (self.minimum ** (1 + self.alpha)
                + val * (self.maximum ** (1 + self.alpha) - self.minimum ** (1 + self.alpha))) ** (
            1. / (1 + self.alpha))

    def prob(self, val):
        val = np.asarray(val, dtype=float)
        inside = (val >= self.minimum) & (val <= self.maximum)
        with np.errstate(divide="ignore", invalid="ignore"):
            if self.alpha == -1:
                return np.nan_to_num(1 / val / np.log(self.maximum / self.minimum)) * inside
            return np.nan_to_num(val ** self.alpha * (1 + self.alpha)
                                 / (self.maximum ** (1 + self.alpha) - self.minimum ** (1 + self.alpha))) * inside


class Gaussian(Prior):
    """bilby/core/prior/analytical.py Gaussian: rescale = mu + erfinv(2 u - 1) sqrt(2) sigma."""

    def __init__(self, mu, sigma, name=None, latex_label=None, unit=None, boundary=None):
        super().__init__(name=name, latex_label=latex_label, unit=unit, boundary=boundary)
        self.mu = mu
        self.sigma = sigma

    def rescale(self, val):
        from scipy.special import erfinv
        return self.mu + erfinv(2 * np.asarray(val) - 1) * 2 ** 0.5 * self.sigma

    def prob(self, val):
        return np.exp(-(self.mu - np.asarray(val)) ** 2 / (2 * self.sigma ** 2)) / (2 * np.pi) ** 0.5 / self.sigma


class Cosine(Prior):
    def __init__(self, minimum=-np.pi / 2, maximum=np.pi / 2, name=None, latex_label=None, unit=None, boundary=None):
        super().__init__(name=name, latex_label=latex_label, minimum=minimum, maximum=maximum,
                         boundary=boundary, unit=unit)

    def rescale(self, val):
        norm = 1 / (np.sin(self.maximum) - np.sin(self.minimum))
        return np.arcsin(val / norm + np.sin(self.minimum))

    def prob(self, val):
        val = np.asarray(val)
        return np.cos(val) / 2 * ((val >= self.minimum) & (val <= self.maximum))


class Sine(Prior):
    def __init__(self, minimum=0, maximum=np.pi, name=None, latex_label=None, unit=None, boundary=None):
        super().__init__(name=name, latex_label=latex_label, minimum=minimum, maximum=maximum,
                         boundary=boundary, unit=unit)

    def rescale(self, val):
        norm = 1 / (np.cos(self.minimum) - np.cos(self.maximum))
        return np.arccos(np.cos(self.minimum) - val / norm)

    def prob(self, val):
        val = np.asarray(val)
        return np.sin(val) / 2 * ((val >= self.minimum) & (val <= self.maximum))


class PriorDict(dict):
    def sample(self, size=None, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        out = {}
        for key, p in self.items():
            if isinstance(p, Prior):
                out[key] = p.sample(size, rng=rng)
            else:
                out[key] = p if size is None else np.full(size, float(p))
        return out

    def copy(self):
        return PriorDict(self)
