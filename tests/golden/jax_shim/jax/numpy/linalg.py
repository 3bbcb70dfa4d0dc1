"""``jax.numpy.linalg`` subset."""
import torch

from .._core import JArray, unwrap


def norm(x, axis=None, ord=None):  # noqa: A002
    t = unwrap(x)
    assert ord in (None, 2)
    # sqrt(sum(x*x)) as in jnp.linalg.norm: the derivative at the origin is NaN (reference Q6)
    return JArray(torch.sqrt((t * t).sum() if axis is None else (t * t).sum(dim=axis)))


def cholesky(x):
    return JArray(torch.linalg.cholesky(unwrap(x)))
