"""Validation statistics for force maps: projections of mapped forces on random Gaussian force
fields and force-residual shifts (reference ``src/aggforce/jaxmapval.py``, which needs JAX).

The reference builds, per sample, a random pairwise potential on the MAPPED coordinates -- one
Gaussian of every entry of the squared distance matrix, offset drawn from ``randg`` -- gets its
forces by autodiff and reduces ``sum(F * G) / n_frames`` (``random_force_proj``) or
``mean((F - G)**2) - mean(F**2)`` (``random_residual_shift``).  Here the force field has its closed
form and, for the default ``method=rsqpg_forces``, ALL samples are evaluated in one pass over the
frames by ``agf_gauss_field_moments`` (``csrc/mapval.cu``) in float64; offsets are drawn from
``randg`` in the reference's order (one ``randg.random()`` per sample), so seeded runs select the
same force fields.  Any other ``method`` callable runs sample by sample like the reference.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Union

import numpy as np
import numpy.random as r
import torch

from . import _engine, _lib
from .agg import force_smoothness


def random_uniform_forces(positions, scale: float = 1.0, randg: Optional[r.Generator] = None) -> np.ndarray:
    """The same random 3-vector of magnitude ``scale`` on every site of every frame
    (reference ``jaxmapval.py:29-75``)."""
    if randg is None:
        randg = r.default_rng()
    force = 2 * randg.random(size=3) - 1
    force /= ((force**2).sum()) ** 0.5
    force *= scale
    return np.broadcast_to(force[None, None, :], tuple(positions.shape[:2]) + (3,)).copy()


def _is_dev(x) -> bool:
    return isinstance(x, torch.Tensor) and x.is_cuda


def _dev(x) -> torch.Tensor:
    if _is_dev(x):
        t = x
    else:
        t = torch.as_tensor(np.ascontiguousarray(x)).to(_engine.device())
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    return t.contiguous()


def sq_gaussian_forces(positions, offset: float, width: float):
    """Forces of ``E = sum_ij exp(-((|x_j - x_i|**2 - offset) / width)**2)`` for every frame
    (reference ``jaxmapval.py:365-401``).  Same array kind and dtype out as in."""
    x = _dev(positions)
    out = torch.empty_like(x)
    _lib.call("agf_sq_gaussian_forces", _engine.ptr(x), _engine.dtype_code(x), x.shape[0], x.shape[1], float(offset),
              float(width), _engine.ptr(out), _engine.dtype_code(out), _engine.stream_ptr())
    return out if _is_dev(positions) else _engine.to_host(out)


def _draw(inner: float, outer: float, width: float, randg: Optional[r.Generator], sq_args: bool):
    if sq_args:
        outer, inner, width = outer**2, inner**2, width**2
    if randg is None:
        randg = r.default_rng()
    return randg.random() * (outer - inner) + inner, width


def rsqpg_forces(positions, inner: float, outer: float, width: float, randg: Optional[r.Generator] = None,
                 sq_args: bool = True):
    """Forces of a random Gaussian force field: the offset is drawn uniformly between ``inner`` and
    ``outer`` (all three parameters squared first when ``sq_args``), reference ``jaxmapval.py:78-139``."""
    offset, width = _draw(inner, outer, width, randg, sq_args)
    return sq_gaussian_forces(positions, offset, width)


def mscg_ip(forces, funcs) -> float:
    """``sum(funcs * forces) / n_steps`` (reference ``jaxmapval.py:322-360``)."""
    n_steps = forces.shape[0]
    if _is_dev(forces) or _is_dev(funcs):
        return float((_dev(funcs).double() * _dev(forces).double()).sum().item() / n_steps)
    return float((np.asarray(funcs, dtype=np.float64) * np.asarray(forces, dtype=np.float64)).sum() / n_steps)


def _field_moments(coords, forces, n_samples: int, randg, inner: float, outer: float, width: float,
                   sq_args: bool = True) -> np.ndarray:
    """``[n_samples, 2]``: ``sum F.G_s`` and ``sum |G_s|^2`` over all frames and sites (one kernel)."""
    if randg is None:
        randg = r.default_rng()
    drawn = [_draw(inner, outer, width, randg, sq_args) for _ in range(n_samples)]
    offsets = np.asarray([d[0] for d in drawn], dtype=np.float64)
    x, f = _dev(coords), _dev(forces)
    if x.shape != f.shape:
        raise ValueError(f"coords {tuple(x.shape)} and forces {tuple(f.shape)} must have the same shape")
    if f.dtype != x.dtype:
        f = f.to(x.dtype)
    out = torch.zeros((n_samples, 2), dtype=torch.float64, device=x.device)
    if n_samples:
        d_off = torch.as_tensor(offsets, device=x.device)
        _lib.call("agf_gauss_field_moments", _engine.ptr(x), _engine.ptr(f), _engine.dtype_code(x), x.shape[0],
                  x.shape[1], _engine.ptr(d_off), n_samples, float(drawn[0][1]), _engine.ptr(out),
                  _engine.stream_ptr())
    return _engine.to_host(out)


def random_residual_shift(coords, forces, n_samples: int = 1000, randg: Optional[r.Generator] = None,
                          method: Callable = rsqpg_forces, average: bool = False, **kwargs
                          ) -> Union[float, List[float]]:
    """Force-residual difference between random force fields and the zero force field:
    ``force_smoothness(forces - G_s) - force_smoothness(forces)`` per sample (reference
    ``jaxmapval.py:159-237``); their mean when ``average``."""
    if method is rsqpg_forces:
        mom = _field_moments(coords, forces, n_samples, randg, **kwargs)
        count = float(np.prod(forces.shape))
        vals = [float((gg - 2.0 * ip) / count) for ip, gg in mom]
    else:
        if randg is None:
            randg = r.default_rng()
        base = force_smoothness(forces)
        vals = []
        for _ in range(n_samples):
            trial = method(coords, randg=randg, **kwargs)
            diff = (_dev(forces).double() - _dev(trial).double()) if (_is_dev(forces) or _is_dev(trial)) else (
                np.asarray(forces, dtype=np.float64) - np.asarray(trial, dtype=np.float64))
            vals.append(force_smoothness(diff) - base)
    if average:
        return sum(vals) / n_samples
    return vals


def random_force_proj(coords, forces, n_samples: int = 1000, randg: Optional[r.Generator] = None,
                      method: Callable = rsqpg_forces, average: bool = True, **kwargs
                      ) -> Union[float, List[float]]:
    """MSCG-style projections ``sum(forces * G_s) / n_frames`` on ``n_samples`` random force fields
    (reference ``jaxmapval.py:266-319``); their mean when ``average``."""
    if method is rsqpg_forces:
        mom = _field_moments(coords, forces, n_samples, randg, **kwargs)
        vals = [float(ip / forces.shape[0]) for ip, _ in mom]
    else:
        if randg is None:
            randg = r.default_rng()
        vals = [mscg_ip(forces, method(coords, randg=randg, **kwargs)) for _ in range(n_samples)]
    if average:
        return sum(vals) / n_samples
    return vals
