from . import prior, likelihood, utils  # noqa: F401
