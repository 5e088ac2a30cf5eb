"""Likelihood base class: the API contract of bilby/core/likelihood.py:13-70."""


class Likelihood:
    def __init__(self):
        self._marginalized_parameters = []
        self._meta_data = None

    def __repr__(self):
        return self.__class__.__name__ + "()"

    def log_likelihood(self, parameters=None):
        return float("nan")

    def noise_log_likelihood(self):
        return float("nan")

    def log_likelihood_ratio(self, parameters=None):
        return self.log_likelihood(parameters) - self.noise_log_likelihood()

    @property
    def meta_data(self):
        return self._meta_data

    @meta_data.setter
    def meta_data(self, meta_data):
        if isinstance(meta_data, dict) or meta_data is None:
            self._meta_data = meta_data
        else:
            raise ValueError("The meta_data must be an instance of dict")

    @property
    def marginalized_parameters(self):
        return self._marginalized_parameters
