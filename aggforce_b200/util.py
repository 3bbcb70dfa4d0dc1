"""Array primitives and small functional helpers.

``distances`` / ``trjdot`` keep the reference's signatures (src/aggforce/util.py:12-125) but
run on the GPU: ``trjdot`` with a 2-D factor is kernel (d); the remaining flavours
(displacement tensors, per-frame factors) are thin torch expressions on the device, they are
not part of the streamed hot path.  ``Curry`` (src/aggforce/util.py:181-252) bakes keyword
arguments into featurizers and is inspected by ``qp_feat_linear_map`` to recognise
``gb_feat`` configurations.
"""
from __future__ import annotations

from typing import Any, Callable, Generic, Iterable, List, TypeVar, Union

import numpy as np
import torch

from . import _engine

T = TypeVar("T")


def _is_device(x) -> bool:
    return isinstance(x, torch.Tensor) and x.is_cuda


def distances(xyz, cross_xyz=None, return_matrix: bool = True, return_displacements: bool = False):
    """Per-frame distance matrices; semantics of src/aggforce/util.py:12-76.

    Output ``[t, i, j] = |xyz[t, j] - other[t, i]|`` in the input dtype.  The result is
    O(T n^2): use :func:`guess_pairwise_constraints` for statistics over long trajectories.
    """
    if cross_xyz is not None and not return_matrix:
        raise ValueError("Cross distances only supported when return_matrix is truthy.")
    if return_displacements and not return_matrix:
        raise ValueError("Displacements only supported when return_matrix is truthy.")
    host = not _is_device(xyz)
    dev = _engine.device()
    x = torch.as_tensor(xyz).to(dev)
    o = x if cross_xyz is None else torch.as_tensor(cross_xyz).to(dev)
    disp = x[:, None, :, :] - o[:, :, None, :]
    if return_displacements:
        return _engine.to_host(disp) if host else disp
    dist = torch.linalg.vector_norm(disp, dim=-1)
    if not return_matrix:
        i0, i1 = torch.triu_indices(dist.shape[-1], dist.shape[-1], offset=1, device=dev)
        dist = dist[:, i0, i1]
    return _engine.to_host(dist) if host else dist


def trjdot(points, factor):
    """``(points * factor)`` for mdtraj-style arrays; semantics of src/aggforce/util.py:79-125.

    ``factor`` of shape (n_cg, n_fg) runs kernel (d); a per-frame factor (T, n_cg, n_fg) is
    contracted on the device frame by frame.
    """
    f = factor if isinstance(factor, torch.Tensor) else np.asarray(factor)
    if f.ndim == 2:
        frames = _engine.Frames(points)
        fm = f.cpu().numpy() if isinstance(f, torch.Tensor) else f
        cm = _engine.CompiledMap(fm, keep_zero_columns=True)
        out, _, _ = _engine.map_apply(frames, cm, nan_mode=0, nan_atol=0.0)
        if frames.np_dtype == np.float32 and fm.dtype != np.float32:
            pass  # numpy promotion: f32 points x f64 factor -> f64 (already the kernel's output)
        return _engine.to_host(out) if frames.on_host else out
    if f.ndim == 3:
        host = not _is_device(points)
        dev = _engine.device()
        p = torch.as_tensor(points).to(dev)
        ft = torch.as_tensor(f).to(dev)
        dt = torch.promote_types(p.dtype, ft.dtype)
        out = torch.einsum("tfd,tcf->tcd", p.to(dt), ft.to(dt))
        return _engine.to_host(out) if host else out
    raise ValueError("Factor matrix is an incompatible shape.")


def flatten(nested_list: Iterable[Iterable[Any]]) -> List[Any]:
    """``[[1, 2], [3, 4]] -> [1, 2, 3, 4]``."""
    out: List[Any] = []
    for sub in nested_list:
        out.extend(sub)
    return out


def curry(func: Callable[..., T], *args: Any, **kwargs: Any) -> Callable[..., T]:
    """Closure form of :class:`Curry`."""

    def wrapped(*sub_args: Any, **sub_kwargs: Any) -> T:
        return func(*sub_args, *args, **sub_kwargs, **kwargs)

    return wrapped


class Curry(Generic[T]):
    """Callable with trailing positional and keyword arguments baked in.

    ``Curry(f, a, k=v)(x)`` evaluates ``f(x, a, k=v)``.  Attributes ``func``, ``args`` and
    ``kwargs`` are public so the configuration can be inspected (and so that
    ``qp_feat_linear_map`` can route ``Curry(gb_feat, ...)`` to the fused kernel).
    """

    def __init__(self, func: Callable[..., T], *args: Any, **kwargs: Any) -> None:
        self.func = func
        self.args = args
        self.kwargs = kwargs

    def __call__(self, *args: Any, **kwargs: Any) -> T:
        return self.func(*args, *self.args, **kwargs, **self.kwargs)

    def _describe(self) -> List[str]:
        return [
            f"{self.__class__.__name__} instance",
            f"func: {self.func!r}",
            f"args: {self.args!r}",
            f"kwargs: {self.kwargs!r}",
        ]

    def __str__(self) -> str:
        return "\n".join(self._describe())

    def __repr__(self) -> str:
        return "; ".join(self._describe())
