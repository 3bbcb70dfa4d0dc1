"""Trajectory-level maps (coordinates and forces together).

API of the reference's ``src/aggforce/map/tmap.py``.  These are thin compositions; the work
happens in ``LinearMap`` / ``CLAMap`` (kernel (d)) and in the augmenters.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Callable, Final, Iterable, Optional, Tuple, TypeVar
from warnings import warn

import numpy as np
import torch

from ..trajectory import AugmentedTrajectory, Augmenter, CoordsTrajectory, ForcesTrajectory, Trajectory
from .core import CLAMap, LinearMap

ArrayTransform = Callable[[Any], Any]
_T_TMap = TypeVar("_T_TMap", bound="TMap")


def _cat_sites(a, b):
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        return torch.cat([torch.as_tensor(a), torch.as_tensor(b).to(torch.as_tensor(a).device)], dim=1)
    return np.concatenate([a, b], axis=1)


class TMap(ABC):
    """Maps a ``Trajectory`` to a new ``Trajectory``."""

    @abstractmethod
    def __init__(self) -> None:
        """Initialize."""

    @abstractmethod
    def __call__(self, t: Trajectory) -> Trajectory:
        """Map a trajectory."""

    def map_arrays(self, coords, forces) -> Tuple[Any, Any]:
        """Map a (coords, forces) pair of arrays; wraps ``__call__``."""
        derived = self(Trajectory(coords=coords, forces=forces))
        return (derived.coords, derived.forces)

    @abstractmethod
    def astype(self: _T_TMap, *args, **kwargs) -> _T_TMap:
        """Convert the map to a numpy precision."""


def _astype_pair(obj, *args, **kwargs):
    try:
        return obj.__class__(
            coord_map=obj.coord_map.astype(*args, **kwargs),
            force_map=obj.force_map.astype(*args, **kwargs),
        )
    except AttributeError as e:
        raise TypeError("Underlying coord_map and/or force_map do not support astype.") from e


class SeperableTMap(TMap):
    """Independent array maps for coordinates and forces (two kernel (d) launches)."""

    def __init__(self, coord_map: ArrayTransform, force_map: ArrayTransform) -> None:
        self.coord_map = coord_map
        self.force_map = force_map

    def __call__(self, t: Trajectory) -> Trajectory:
        return Trajectory(coords=self.coord_map(t.coords), forces=self.force_map(t.forces))

    def astype(self, *args, **kwargs) -> "SeperableTMap":
        return _astype_pair(self, *args, **kwargs)


class CLAFTMap(TMap):
    """Linear coordinate map plus configuration-dependent (``CLAMap``) force map."""

    def __init__(self, coord_map: ArrayTransform, force_map: CLAMap) -> None:
        self.coord_map = coord_map
        self.force_map = force_map

    def __call__(self, t: Trajectory) -> Trajectory:
        return Trajectory(
            coords=self.coord_map(t.coords),
            forces=self.force_map(points=t.forces, copoints=t.coords),
        )

    def astype(self, *args, **kwargs) -> "CLAFTMap":
        return _astype_pair(self, *args, **kwargs)


class AugmentedTMap(TMap):
    """Augment the input (fresh noise on every call), then apply ``aug_tmap``."""

    def __init__(self, aug_tmap: TMap, augmenter: Augmenter, kbt: float) -> None:
        self.tmap: Final = aug_tmap
        self.augmenter: Final = augmenter
        self.kbt: Final = kbt

    def __call__(self, t: Trajectory) -> Trajectory:
        new_draw = getattr(self.augmenter, "new_draw", None)
        if (new_draw is not None and isinstance(self.tmap, SeperableTMap) and isinstance(self.tmap.coord_map, LinearMap)
                and isinstance(self.tmap.force_map, LinearMap)):
            # device augmenter + linear maps: augmented coordinates / forces are generated slab by slab
            # in front of the apply kernel (same noise realisation for both), never materialised
            from ..trajectory.gausstraj import AugmentedFrames

            draw = new_draw()
            coords = AugmentedFrames(t.coords, self.augmenter, self.kbt, draw, "coords")
            forces = AugmentedFrames(t.forces, self.augmenter, self.kbt, draw, "forces")
            return Trajectory(coords=self.tmap.coord_map(coords), forces=self.tmap.force_map(forces))
        return self.tmap(AugmentedTrajectory.from_trajectory(t=t, kbt=self.kbt, augmenter=self.augmenter))

    def astype(self, *args, **kwargs) -> "AugmentedTMap":
        return self.__class__(
            aug_tmap=self.tmap.astype(*args, **kwargs),
            augmenter=self.augmenter.astype(*args, **kwargs),
            kbt=self.kbt,
        )


class ComposedTMap(TMap):
    """Composition of TMaps; the right-most entry of ``submaps`` is applied first."""

    def __init__(self, submaps: Iterable[TMap]) -> None:
        self.submaps: Final = list(submaps)

    def __call__(self, t: Trajectory) -> Trajectory:
        for mapping in self.submaps[::-1]:
            t = mapping(t)
        return t

    def __getitem__(self, idx: int, /) -> TMap:
        return self.submaps[idx]

    def astype(self, *args, **kwargs) -> "ComposedTMap":
        return self.__class__(submaps=[m.astype(*args, **kwargs) for m in self.submaps])


class NullForcesTMap(TMap):
    """Attach (or overwrite) forces with a constant fill, ``nan`` by default."""

    def __init__(self, warn_input_forces: bool = True, fill_value: Any = np.nan) -> None:
        self.warn_input_forces = warn_input_forces
        self.fill_value = fill_value

    def __call__(self, t: CoordsTrajectory) -> Trajectory:
        if isinstance(t, ForcesTrajectory) and self.warn_input_forces:
            warn("Discarding forces on input trajectory.", stacklevel=0)
        return Trajectory(coords=t.coords, forces=self.fill_value * t.coords)

    def map_arrays(self, coords, forces: Optional[Any] = None) -> Tuple[Any, Any]:
        t = CoordsTrajectory(coords=coords) if forces is None else Trajectory(coords=coords, forces=forces)
        derived = self(t)
        return (derived.coords, derived.forces)

    def astype(self, *args, **kwargs) -> "NullForcesTMap":  # noqa: ARG002
        return self.__class__(warn_input_forces=self.warn_input_forces, fill_value=self.fill_value)


class RATMap:
    """Map the real particles of an ``AugmentedTrajectory``; augmented particles pass through."""

    def __init__(self, tmap: TMap) -> None:
        self.tmap = tmap

    def __call__(self, t: AugmentedTrajectory) -> Trajectory:
        coords, forces = self.tmap.map_arrays(t.coords[:, t.real_slice, :], t.forces[:, t.real_slice, :])
        return Trajectory(
            coords=_cat_sites(coords, t.coords[:, t.aug_slice, :]),
            forces=_cat_sites(forces, t.forces[:, t.aug_slice, :]),
        )
