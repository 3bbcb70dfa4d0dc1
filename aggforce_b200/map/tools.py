"""Helpers that build special-purpose maps (reference ``src/aggforce/map/tools.py``)."""
from __future__ import annotations

from itertools import combinations
from typing import Iterable, Union

import numpy as np

from ..trajectory import AugmentedTrajectory
from .core import LinearMap


def lmap_augvariables(aug: AugmentedTrajectory) -> LinearMap:
    """Slice map isolating the augmenting sites (the trailing ``n_aug_sites``) of ``aug``."""
    return LinearMap([[s] for s in range(aug.n_real_sites, aug.n_sites)], n_fg_sites=aug.n_sites)


def smear_map(site_groups: Iterable[Iterable[int]], n_sites: int,
              return_mapping_matrix: bool = False) -> Union[LinearMap, np.ndarray]:
    """(n_sites, n_sites) map replacing every listed group of sites by the group mean.

    Float32 matrix as in the reference (tools.py:96-100); groups must be disjoint.
    """
    groups = [sorted(set(g)) for g in site_groups]
    for a, b in combinations(groups, 2):
        if set(a).intersection(b):
            raise ValueError("Site definitions in site_groups overlap; merge before passing.")
    matrix = np.eye(n_sites, dtype=np.float32)
    for g in groups:
        matrix[np.ix_(g, g)] = 1 / len(g)
    return matrix if return_mapping_matrix else LinearMap(mapping=matrix)
