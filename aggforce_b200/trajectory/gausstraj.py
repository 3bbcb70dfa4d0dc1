"""Gaussian augmenters.

``SimpleCondNormal``: isotropic noise on every site, identity pre-map (host numpy; reference
``src/aggforce/trajectory/simplegausstraj.py:13-137``).

``CondNormal``: ``y = A x + eps, eps ~ N(0, var I)`` with ``A`` a ``LinearMap`` -- the B200
replacement for the reference's JAX ``JCondNormal`` (``jaxgausstraj.py:99-402``), whose autodiff
log-gradients have the closed form ``grad_y = -(y - A x)/var``, ``grad_x = A^T (y - A x)/var``.
Noise is drawn on the device (counter-free torch Philox generator, seeded); it cannot match
JAX's threefry stream, so parity tests inject the draw through ``noise=``.
"""
from __future__ import annotations

from typing import Any, Optional, Tuple

import numpy as np
import torch

from .. import _engine
from .augment import Augmenter


class SimpleCondNormal(Augmenter):
    """Adds independent N(0, var) noise to every coordinate (identity pre-map)."""

    def __init__(self, var: float, seed: Optional[int] = None, dtype: Any = None) -> None:
        self.var = var
        self._rng = np.random.default_rng(seed)
        self.dtype = np.dtype(np.float32 if dtype is None else dtype)

    def sample(self, source: np.ndarray) -> np.ndarray:
        noise = np.sqrt(self.var) * self._rng.standard_normal(source.shape, dtype=self.dtype)
        return (source + noise).astype(self.dtype, copy=False)

    def log_gradient(self, source: np.ndarray, generated: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        wrt_generated = (-(1.0 / self.var) * (generated - source)).astype(self.dtype, copy=False)
        return -wrt_generated, wrt_generated

    def astype(self, dtype, *args, **kwargs) -> "SimpleCondNormal":  # noqa: ARG002
        return self.__class__(var=self.var, dtype=dtype)


class CondNormal(Augmenter):
    """``y = premap(x) + N(0, cov)`` with a linear ``premap`` and scalar (isotropic) ``cov``.

    ``premap`` is a ``LinearMap`` (or its bound ``flat_call``); ``None`` means identity.
    ``noise`` optionally injects the standard-normal draw ``(n_frames, n_new, 3)`` used by the
    next ``sample``/``augment`` call (test hook).
    """

    n_dim = 3

    def __init__(self, cov: float, premap=None, source_postmap=None, seed: Optional[int] = None,
                 dtype: Any = None, noise=None) -> None:
        if source_postmap is not None:
            raise NotImplementedError("source_postmap is not supported by the B200 augmenter")
        if not np.isscalar(cov):
            raise NotImplementedError("only scalar (isotropic) covariances are supported")
        owner = getattr(premap, "__self__", None)
        self.premap = owner if owner is not None else premap
        self.cov = float(cov)
        self.seed = int(np.random.default_rng().integers(0, 10**6)) if seed is None else int(seed)
        self.dtype = np.dtype(np.float32 if dtype is None else dtype)
        self._gen: Optional[torch.Generator] = None
        self._noise = noise

    # -- helpers
    def _tdtype(self) -> torch.dtype:
        return torch.float32 if self.dtype == np.float32 else torch.float64

    def _mean(self, source):
        return source if self.premap is None else self.premap(source)

    def _draw(self, shape, device) -> torch.Tensor:
        if self._noise is not None:
            z = torch.as_tensor(self._noise).to(device=device, dtype=self._tdtype())
            self._noise = None
            return z
        if self._gen is None:
            self._gen = torch.Generator(device=device)
            self._gen.manual_seed(self.seed)
        return torch.randn(shape, generator=self._gen, device=device, dtype=self._tdtype())

    def _back(self, eps: torch.Tensor):
        """A^T eps for the pre-map (identity when premap is None)."""
        if self.premap is None:
            return eps
        return self.premap.T(eps)

    # -- Augmenter interface
    def sample(self, source):
        host = not (isinstance(source, torch.Tensor) and source.is_cuda)
        dev = _engine.device()
        mean = torch.as_tensor(self._mean(source)).to(device=dev, dtype=self._tdtype())
        y = mean + np.sqrt(self.cov) * self._draw(mean.shape, dev)
        return _engine.to_host(y) if host else y

    def log_gradient(self, source, generated):
        host = not (isinstance(source, torch.Tensor) and source.is_cuda)
        dev = _engine.device()
        mean = torch.as_tensor(self._mean(source)).to(device=dev, dtype=self._tdtype())
        resid = torch.as_tensor(generated).to(device=dev, dtype=self._tdtype()) - mean
        wrt_generated = -resid / self.cov
        wrt_source = torch.as_tensor(self._back(resid / self.cov)).to(device=dev, dtype=self._tdtype())
        if host:
            return _engine.to_host(wrt_source), _engine.to_host(wrt_generated)
        return wrt_source, wrt_generated

    def augment(self, coords, forces, kbt: float):
        """Augmented ``(coords, forces)`` in one device pass (trajectory/core.py:382-389 of the reference)."""
        host = not (isinstance(coords, torch.Tensor) and coords.is_cuda)
        dev = _engine.device()
        td = self._tdtype()
        x = torch.as_tensor(coords).to(device=dev, dtype=td)
        f = torch.as_tensor(forces).to(device=dev, dtype=td)
        mean = torch.as_tensor(self._mean(x)).to(dtype=td)
        eps = np.sqrt(self.cov) * self._draw(mean.shape, dev)
        back = torch.as_tensor(self._back(eps)).to(dtype=td)
        full_coords = torch.cat([x, mean + eps], dim=1)
        full_forces = torch.cat([f + (kbt / self.cov) * back, (-kbt / self.cov) * eps], dim=1)
        if host:
            return _engine.to_host(full_coords), _engine.to_host(full_forces)
        return full_coords, full_forces

    def astype(self, dtype, *args, **kwargs) -> "CondNormal":  # noqa: ARG002
        return self.__class__(cov=self.cov, premap=self.premap, seed=self.seed, dtype=dtype)


# drop-in name of the reference's JAX class
JCondNormal = CondNormal
