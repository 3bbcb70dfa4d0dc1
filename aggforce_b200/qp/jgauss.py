"""Noised ("Gaussian") force maps (reference ``src/aggforce/qp/jgauss.py:27-140``).

``joptgauss_map``: augment the trajectory with one noise site per bead (``y = A x + eps``), fit
an optimal LINEAR force map on the augmented trajectory whose coordinate map isolates the
noise sites (kernel (a) on ``n_fg + n_cg`` columns), and wrap it so that applying it to a plain
trajectory first re-augments with fresh noise.  The augmentation runs on the device
(``trajectory.gausstraj.CondNormal``); the reference's JAX ``JCondNormal`` is replaced by its
closed form.
"""
from __future__ import annotations

from typing import Optional

from ..constraints import Constraints
from ..map import AugmentedTMap, LinearMap, lmap_augvariables
from ..trajectory import AugmentedTrajectory, CondNormal, Trajectory
from .qplinear import qp_linear_map


def joptgauss_map(
    traj: Trajectory,
    coord_map: LinearMap,
    var: float,
    kbt: float,
    constraints: Optional[Constraints] = None,
    seed: Optional[int] = None,
    noise=None,
    **kwargs,
) -> AugmentedTMap:
    """Optimised Gaussian map.

    ``var``: variance of the isotropic noise added to the mapped positions; ``kbt``: thermal
    energy turning log-density gradients into forces; ``constraints`` refer to the real sites
    (noise sites are appended after them, so the indices stay valid); ``seed`` seeds the noise;
    ``**kwargs`` go to ``qp_linear_map`` on the augmented trajectory.  ``noise`` (extra, for
    tests): standard-normal array ``(n_frames, n_cg, 3)`` used for the FIT's augmentation instead
    of a random draw.

    The returned map is stochastic: every application draws new noise.
    """
    augmenter = CondNormal(cov=var, premap=coord_map, seed=seed, noise=noise)
    aug_traj = AugmentedTrajectory.from_trajectory(t=traj, augmenter=augmenter, kbt=kbt)
    aug_coord_map = lmap_augvariables(aug_traj)
    aug_tmap = qp_linear_map(traj=aug_traj, coord_map=aug_coord_map, constraints=constraints, **kwargs)
    return AugmentedTMap(aug_tmap=aug_tmap, augmenter=augmenter, kbt=kbt)
