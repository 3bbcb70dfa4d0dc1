"""Noised ("Gaussian") force maps (reference ``src/aggforce/qp/jgauss.py:27-140``).

``joptgauss_map``: augment the trajectory with one noise site per bead (``y = A x + eps``), fit
an optimal LINEAR force map on the augmented trajectory whose coordinate map isolates the
noise sites (kernel (a) on ``n_fg + n_cg`` columns), and wrap it so that applying it to a plain
trajectory first re-augments with fresh noise.  The augmentation is one fused kernel
(``agf_gauss_augment``; the reference's JAX ``JCondNormal`` is replaced by its closed form) that
runs slab by slab in front of the Gram / apply kernels, so the augmented arrays are never
materialised.
"""
from __future__ import annotations

from typing import Optional

from types import SimpleNamespace

from ..constraints import Constraints
from ..map import AugmentedTMap, LinearMap
from ..trajectory import CondNormal, Trajectory
from ..trajectory.gausstraj import AugmentedFrames
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
    # the reference materialises AugmentedTrajectory.from_trajectory(traj) (jgauss.py:129-133); here the
    # augmented forces are generated slab by slab while the Gram kernel consumes them
    n_real, n_new = coord_map.n_fg_sites, coord_map.n_cg_sites
    aug_forces = AugmentedFrames(traj.forces, augmenter, kbt, augmenter.new_draw(), "forces")
    aug_coord_map = LinearMap([[s] for s in range(n_real, n_real + n_new)], n_fg_sites=n_real + n_new)
    aug_tmap = qp_linear_map(traj=SimpleNamespace(forces=aug_forces), coord_map=aug_coord_map,
                             constraints=constraints, **kwargs)
    return AugmentedTMap(aug_tmap=aug_tmap, augmenter=augmenter, kbt=kbt)
