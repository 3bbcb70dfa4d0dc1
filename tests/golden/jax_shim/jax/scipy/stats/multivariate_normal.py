"""``jax.scipy.stats.multivariate_normal.logpdf`` (Cholesky form, as JAX evaluates it)."""
import math

import torch

from ..._core import JArray, unwrap


def logpdf(x, mean, cov):
    xt, mt, ct = unwrap(x), unwrap(mean), unwrap(cov)
    n = mt.shape[-1]
    chol = torch.linalg.cholesky(ct)
    y = torch.linalg.solve_triangular(chol, (xt - mt)[..., None], upper=False)[..., 0]
    return JArray(-0.5 * (y * y).sum(-1) - 0.5 * n * math.log(2 * math.pi) - torch.log(torch.diagonal(chol)).sum())
