"""``jax.random`` subset.  The threefry stream is NOT reproduced: keys are integers, draws come
from a seeded torch generator, and every standard-normal draw is appended to ``DRAWS`` so that
fixture scripts can store the exact noise the reference consumed."""
import torch

from ._core import _F, JArray, _torch_dtype, unwrap

DRAWS: list = []


def PRNGKey(seed):  # noqa: N802
    return JArray(torch.tensor([0, int(seed)], dtype=torch.int64))


def split(key, num=2):
    base = int(unwrap(key).reshape(-1)[-1])
    return JArray(torch.tensor([[i + 1, (base * 6364136223846793005 + 1442695040888963407 * (i + 1)) % (1 << 62)]
                                for i in range(num)], dtype=torch.int64))


def multivariate_normal(key, mean, cov, shape=None, dtype=None):
    """mean + L z with L = cholesky(cov), z ~ N(0, I) (what jax.random.multivariate_normal computes with
    its default method="cholesky")."""
    m, c = unwrap(mean), unwrap(cov)
    gen = torch.Generator().manual_seed(int(unwrap(key).reshape(-1)[-1]) % (1 << 62))
    z = torch.randn(m.shape, generator=gen, dtype=torch.float64).to(_F)
    DRAWS.append(z.clone())
    chol = torch.linalg.cholesky(c.reshape(c.shape[-2:]).to(_F))
    out = m.to(_F) + torch.einsum("ij,...j->...i", chol, z)
    td = _torch_dtype(dtype)
    return JArray(out if td is None else out.to(td))
