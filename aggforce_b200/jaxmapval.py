"""Drop-in name of the reference's ``aggforce.jaxmapval`` (no JAX here: see ``mapval``)."""
from .mapval import *  # noqa: F401,F403
from .mapval import (  # noqa: F401
    mscg_ip,
    random_force_proj,
    random_residual_shift,
    random_uniform_forces,
    rsqpg_forces,
    sq_gaussian_forces,
)
