"""Trajectory containers (coordinates and/or forces over frames).

Mirrors the reference's ``src/aggforce/trajectory/core.py``.  Arrays may be numpy arrays
(host) or torch CUDA tensors (device resident, required once data outgrows host memory);
the containers only check shapes and never copy.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Any, Callable, NoReturn, Optional, Tuple, TypeVar

import numpy as np
import torch

from .augment import Augmenter

A = TypeVar("A")


def _copy(x):
    return x.clone() if isinstance(x, torch.Tensor) else x.copy()


def _astype(x, *args, **kwargs):
    if isinstance(x, torch.Tensor):
        dt = np.dtype(args[0] if args else kwargs["dtype"])
        return x.to({np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[dt])
    return x.astype(*args, **kwargs)


def _cat(a, b):
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        ta = torch.as_tensor(a)
        tb = torch.as_tensor(b).to(ta.device)
        dt = torch.promote_types(ta.dtype, tb.dtype)
        return torch.cat([ta.to(dt), tb.to(dt)], dim=1)
    return np.concatenate([a, b], axis=1)


def _need_slice(index) -> None:
    if not isinstance(index, slice):
        raise ValueError("Only slices are allowed for indexing.")


class ForcesTrajectory:
    """Forces ``(n_frames, n_sites, n_dim)`` without positions."""

    def __init__(self, *, forces) -> None:
        if len(forces.shape) != 3:
            raise ValueError("forces must have 3 dimensions.")
        self.forces = forces

    @property
    def n_sites(self) -> int:
        return self.forces.shape[1]

    @property
    def n_dim(self) -> int:
        return self.forces.shape[2]

    def __len__(self) -> int:
        return len(self.forces)

    def __getitem__(self, index: slice) -> "ForcesTrajectory":
        _need_slice(index)
        return self.__class__(forces=self.forces[index])

    def copy(self) -> "ForcesTrajectory":
        return self.__class__(forces=_copy(self.forces))

    def astype(self, *args, **kwargs) -> "ForcesTrajectory":
        return self.__class__(forces=_astype(self.forces, *args, **kwargs))


class CoordsTrajectory:
    """Positions ``(n_frames, n_sites, n_dim)`` without forces."""

    def __init__(self, *, coords) -> None:
        if len(coords.shape) != 3:
            raise ValueError("coords must have 3 dimensions.")
        self.coords = coords

    @property
    def n_sites(self) -> int:
        return self.coords.shape[1]

    @property
    def n_dim(self) -> int:
        return self.coords.shape[2]

    def __len__(self) -> int:
        return len(self.coords)

    def __getitem__(self, index: slice) -> "CoordsTrajectory":
        _need_slice(index)
        return self.__class__(coords=self.coords[index])

    def copy(self) -> "CoordsTrajectory":
        return self.__class__(coords=_copy(self.coords))

    def astype(self, *args, **kwargs) -> "CoordsTrajectory":
        return self.__class__(coords=_astype(self.coords, *args, **kwargs))


class Trajectory(CoordsTrajectory, ForcesTrajectory):
    """Coordinates and forces of the same shape ``(n_frames, n_sites, n_dim)``."""

    def __init__(self, *, coords, forces) -> None:
        if tuple(coords.shape) != tuple(forces.shape) or len(coords.shape) != 3:
            raise ValueError("coords and forces must be of same shape.")
        CoordsTrajectory.__init__(self, coords=coords)
        ForcesTrajectory.__init__(self, forces=forces)

    def __getitem__(self, index: slice) -> "Trajectory":
        _need_slice(index)
        return Trajectory(coords=self.coords[index], forces=self.forces[index])

    def copy(self) -> "Trajectory":
        return Trajectory(coords=_copy(self.coords), forces=_copy(self.forces))

    def astype(self, *args, **kwargs) -> "Trajectory":
        return self.__class__(
            coords=_astype(self.coords, *args, **kwargs), forces=_astype(self.forces, *args, **kwargs)
        )


class AugmentedTrajectory(Trajectory):
    """Trajectory over ``(x, y)``: real sites ``x`` followed by sites ``y ~ g(.|x)`` drawn by an
    ``Augmenter``; forces are ``kbt * grad log[g(y|x) f(x)]`` (reference core.py:227-303):

        forces_y = kbt * grad_y log g(y|x)
        forces_x = real forces + kbt * grad_x log g(y|x)
    """

    def __init__(
        self,
        *,
        coords,
        forces,
        augmenter: Augmenter,
        kbt: float,
        override_first_augment: Optional[Tuple[Any, Any]] = None,
    ) -> None:
        self.augmenter = augmenter
        self.kbt = kbt
        self._real_forces = forces
        self._real_n_sites = coords.shape[1]
        if override_first_augment is None:
            ext_coords, ext_forces = self._augment(coords, forces)
        else:
            ext_coords, ext_forces = override_first_augment
        super().__init__(coords=ext_coords, forces=ext_forces)

    def _augment(self, coords, forces) -> Tuple[Any, Any]:
        fused = getattr(self.augmenter, "augment", None)
        if fused is not None:  # device augmenters build both arrays in one pass
            return fused(coords, forces, self.kbt)
        aug_coords = self.augmenter.sample(coords)
        real_lgrad, aug_lgrad = self.augmenter.log_gradient(coords, aug_coords)
        return (
            _cat(coords, aug_coords),
            _cat(forces + self.kbt * real_lgrad, self.kbt * aug_lgrad),
        )

    @property
    def real_coords(self):
        return self.coords[:, : self._real_n_sites, :]

    @real_coords.setter
    def real_coords(self, value: Any) -> NoReturn:  # noqa: ARG002
        raise ValueError("real_positions cannot be reassigned.")

    @property
    def real_forces(self):
        """Forces on the real sites BEFORE augmentation (``forces`` holds the corrected ones)."""
        return self._real_forces

    @real_forces.setter
    def real_forces(self, value: Any) -> NoReturn:  # noqa: ARG002
        raise ValueError("real_forces cannot be reassigned.")

    @property
    def n_real_sites(self) -> int:
        return self._real_n_sites

    @property
    def n_aug_sites(self) -> int:
        return self.coords.shape[1] - self._real_n_sites

    @property
    def real_slice(self) -> slice:
        return slice(0, self.n_real_sites)

    @property
    def aug_slice(self) -> slice:
        return slice(self.n_real_sites, self.n_real_sites + self.n_aug_sites)

    def refresh(self) -> None:
        """Redraw the augmenting sites."""
        self.coords, self.forces = self._augment(coords=self.real_coords, forces=self.real_forces)

    def __getitem__(self, index: slice) -> "AugmentedTrajectory":
        _need_slice(index)
        return AugmentedTrajectory(
            coords=self.real_coords[index],
            forces=self.real_forces[index],
            augmenter=self.augmenter,
            kbt=self.kbt,
            override_first_augment=(self.coords[index], self.forces[index]),
        )

    def copy(self) -> "AugmentedTrajectory":
        return self.__class__(
            coords=_copy(self.real_coords),
            forces=_copy(self.real_forces),
            augmenter=deepcopy(self.augmenter),
            kbt=self.kbt,
            override_first_augment=(_copy(self.coords), _copy(self.forces)),
        )

    def astype(self, *args, **kwargs) -> "AugmentedTrajectory":
        return self.__class__(
            coords=_astype(self.real_coords, *args, **kwargs),
            forces=_astype(self.real_forces, *args, **kwargs),
            augmenter=self.augmenter.astype(*args, **kwargs),
            kbt=self.kbt,
            override_first_augment=(_astype(self.coords, *args, **kwargs), _astype(self.forces, *args, **kwargs)),
        )

    def pullback(self, C: Callable[["AugmentedTrajectory"], A], array: bool = False) -> Callable:
        """Callable that augments its input (with this instance's augmenter/kbt) then applies ``C``."""
        if array:

            def from_arrays(coords, forces) -> A:
                return C(self.__class__(coords=coords, forces=forces, augmenter=self.augmenter, kbt=self.kbt))

            return from_arrays

        def from_traj(t: Trajectory) -> A:
            return C(self.__class__(coords=t.coords, forces=t.forces, augmenter=self.augmenter, kbt=self.kbt))

        return from_traj

    @classmethod
    def from_trajectory(cls, t: Trajectory, kbt: float, augmenter: Augmenter) -> "AugmentedTrajectory":
        return cls(coords=t.coords, forces=t.forces, kbt=kbt, augmenter=augmenter)
