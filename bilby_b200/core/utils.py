"""Host-side helpers mirroring bilby/core/utils (series, constants) - only what the hot path needs."""
import logging

import numpy as np

logger = logging.getLogger("bilby_b200")

# bilby/core/utils/constants.py:3-7
speed_of_light = 299792458.0
parsec = 3.085677581491367e+16
solar_mass = 1.988409870698050731911960804878414216e30
gravitational_constant = 6.6743e-11
radius_of_earth = 6378136.6


def create_frequency_series(sampling_frequency, duration):
    """bilby/core/utils/series.py:115-134."""
    number_of_samples = np.round(duration * sampling_frequency)
    number_of_frequencies = int(np.round(number_of_samples / 2) + 1)
    return np.linspace(0, sampling_frequency / 2, num=number_of_frequencies)


def create_time_series(sampling_frequency, duration, starting_time=0.):
    """bilby/core/utils/series.py:88-112."""
    number_of_samples = int(duration * sampling_frequency)
    return np.linspace(starting_time, duration + starting_time - 1 / sampling_frequency, num=number_of_samples)


def create_white_noise(sampling_frequency, duration, rng):
    """bilby/core/utils/series.py:161-198: frequency-domain white noise with unit one-sided PSD."""
    number_of_samples = duration * sampling_frequency
    number_of_samples = int(np.round(number_of_samples))
    frequencies = create_frequency_series(sampling_frequency, duration)
    norm1 = 0.5 * duration ** 0.5
    re1, im1 = rng.normal(0, norm1, (2, len(frequencies)))
    white_noise = re1 + 1j * im1
    white_noise[0] = 0
    if np.mod(number_of_samples, 2) == 0:
        white_noise[-1] = 0
    return white_noise, frequencies


class random:
    """Mirror of ``bilby.core.utils.random`` (core/utils/random.py:25-70): one process-wide numpy Generator,
    re-seeded with ``random.seed(n)``; read it as ``random.rng`` at the point of use."""
    rng = np.random.default_rng()

    @classmethod
    def seed(cls, seed):
        cls.rng = np.random.default_rng(seed)


def pinned_empty(shape, dtype="float64"):
    """numpy array in page-locked host memory (the storage belongs to a torch tensor the array keeps alive): the host
    entry points copy from / to such buffers directly instead of staging them."""
    import numpy as np
    import torch
    t = torch.empty(tuple(np.atleast_1d(shape)), dtype=getattr(torch, str(np.dtype(dtype)))).pin_memory()
    return t.numpy()
