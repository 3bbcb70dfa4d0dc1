from . import multivariate_normal  # noqa: F401
