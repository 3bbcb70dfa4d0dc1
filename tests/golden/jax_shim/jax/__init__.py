"""Stand-in ``jax`` package for fixture generation (see ``_core.py``).  NOT JAX."""
from ._core import JArray as Array, grad, jacfwd, jacrev, jit, vmap  # noqa: F401
from . import numpy, random, scipy  # noqa: F401,E402

__version__ = "0.0-shim"
