"""``jax.numpy`` subset on torch CPU tensors (see ``jax/_core.py``)."""
from __future__ import annotations

import math

import numpy as _np
import torch

from .._core import _F, _I, JArray, _torch_dtype, asarray, unwrap, wrap
from . import linalg  # noqa: F401

float32, int32, float64, int64 = _np.float32, _np.int32, _np.float64, _np.int64
ndarray = JArray
pi = math.pi


def _t(x):
    t = unwrap(x)
    return t if isinstance(t, torch.Tensor) else torch.as_tensor(t, dtype=_F if isinstance(t, float) else None)


def array(x, dtype=None):
    return asarray(x, dtype)


def arange(*a, dtype=None):
    return JArray(torch.arange(*a, dtype=_I))


def zeros(shape, dtype=None):
    return JArray(torch.zeros(shape, dtype=_torch_dtype(dtype) or _F))


def ones(shape, dtype=None):
    return JArray(torch.ones(shape, dtype=_torch_dtype(dtype) or _F))


def linspace(start, stop, num):
    # jnp.linspace computes in float32: start + step * iota with the end point pinned
    s, e = float(start), float(stop)
    if num == 1:
        return JArray(torch.tensor([s], dtype=_F))
    step = torch.tensor((e - s) / (num - 1), dtype=_F)
    out = torch.tensor(s, dtype=_F) + step * torch.arange(num, dtype=_F)
    out[-1] = e
    return JArray(out)


def stack(arrs, axis=0, dtype=None):
    out = torch.stack([_t(a) for a in arrs], dim=axis)
    return JArray(out if dtype is None else out.to(_torch_dtype(dtype)))


def vstack(arrs):
    return JArray(torch.cat([_t(a) if _t(a).ndim > 1 else _t(a)[None] for a in arrs], dim=0))


def concatenate(arrs, axis=0):
    return JArray(torch.cat([_t(a) for a in arrs], dim=axis))


def exp(x):
    return JArray(torch.exp(_t(x)))


def sqrt(x):
    return JArray(torch.sqrt(_t(x)))


def log(x):
    return JArray(torch.log(_t(x)))


def ceil(x):
    return JArray(torch.ceil(_t(x)))


def clip(a=None, a_min=None, a_max=None, **kw):
    lo = kw.get("min", a_min)
    hi = kw.get("max", a_max)
    return JArray(torch.clamp(_t(a), min=lo, max=hi))


def einsum(spec, *ops, optimize=None):
    ts = [_t(o) for o in ops]
    dt = torch.promote_types(ts[0].dtype, ts[1].dtype) if len(ts) > 1 else ts[0].dtype
    return JArray(torch.einsum(spec, *[t.to(dt) for t in ts]))


def array_split(arr, n):
    return [JArray(c) for c in torch.tensor_split(_t(arr), int(n))]


def swapaxes(x, a, b):
    return JArray(_t(x).swapaxes(a, b))


def reshape(a, newshape=None, shape=None):
    return JArray(_t(a).reshape(tuple(newshape if newshape is not None else shape)))


def diag(x):
    return JArray(torch.diag(_t(x)))


def repeat(x, repeats):
    t = _t(x)
    if t.ndim == 0:
        t = t[None]
    return JArray(t.repeat_interleave(repeats))


def nan_to_num(x, nan=0.0):
    t = _t(x)
    return JArray(torch.where(torch.isnan(t), torch.as_tensor(nan, dtype=t.dtype), t))


def allclose(a, b, atol=1e-8, rtol=1e-5):
    return bool(torch.allclose(_t(a), _t(b), atol=atol, rtol=rtol))


def triu_indices(n, k=0):
    i = torch.triu_indices(n, n, offset=k)
    return JArray(i[0]), JArray(i[1])


def sum(x, axis=None):  # noqa: A001
    return asarray(x).sum() if axis is None else asarray(x).sum(axis=axis)


def where(c, a, b):
    return JArray(torch.where(_t(c), _t(a), _t(b)))


__all__ = [n for n in dir() if not n.startswith("_")]
_ = wrap
