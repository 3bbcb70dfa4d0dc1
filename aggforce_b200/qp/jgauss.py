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

import warnings
from types import SimpleNamespace

import numpy as np
import torch

from ..constraints import Constraints
from ..map import AugmentedTMap, ComposedTMap, LinearMap, NullForcesTMap, RATMap, SeperableTMap
from ..trajectory import AugmentedTrajectory, CondNormal, CoordsTrajectory, Trajectory
from ..trajectory.gausstraj import AugmentedFrames
from .basicagg import constraint_aware_uni_map
from .qplinear import qp_linear_map
from .solver import DEFAULT_SOLVER_OPTIONS, SolverOptions


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


# --------------------------------------------------------------------------------------
# staged Gaussian maps (reference jgauss.py:143-650): host re-compositions of
# kernel (a) (fits), kernel (d) (maps) and agf_gauss_augment.  Each returns a ComposedTMap whose
# last entry is a deterministic pre-map (coarse-grain once, save) and whose first entry adds the
# noise to the already mapped data.  ``noise`` (extra, tests) injects the fit's draw.
# --------------------------------------------------------------------------------------
def _noise_site_map(n_sites: int, n_aug: int) -> LinearMap:
    return LinearMap([[i] for i in range(n_sites - n_aug, n_sites)], n_fg_sites=n_sites)


def _pre_tmap(traj, coord_map, force_map, constraints, l2, solver_args) -> SeperableTMap:
    if force_map is None:
        return qp_linear_map(traj=traj, coord_map=coord_map, constraints=constraints, l2_regularization=l2,
                             solver_args=solver_args)
    return SeperableTMap(coord_map=coord_map, force_map=force_map)


def stagedjoptgauss_map(
    traj: Trajectory,
    coord_map: LinearMap,
    var: float,
    kbt: float,
    force_map: Optional[LinearMap] = None,
    constraints: Optional[Constraints] = None,
    seed: Optional[int] = None,
    premap_l2_regularization: float = 0.0,
    premap_solver_args: SolverOptions = DEFAULT_SOLVER_OPTIONS,
    noise=None,
    **kwargs,
) -> ComposedTMap:
    """Optimised Gaussian map with an explicit linear pre-map (reference ``jgauss.py:143-312``).

    1. optimise a noise-free force map (or take ``force_map``); 2. augment the full trajectory;
    3. map its real sites with the pre-map (``RATMap``); 4. optimise a linear map on the result
    whose coordinate map isolates the noise sites; 5. compose.  ``result[1]`` is the pre-map,
    ``result[0]`` the noising map, whose augmenter carries ``source_postmap = force_map @
    coord_map.T`` so that it corrects forces of an already mapped trajectory.
    """
    pre_tmap = _pre_tmap(traj, coord_map, force_map, constraints, premap_l2_regularization, premap_solver_args)
    augmenter = CondNormal(cov=var, premap=pre_tmap.coord_map, seed=seed, noise=noise)
    aug_traj = AugmentedTrajectory.from_trajectory(t=traj, augmenter=augmenter, kbt=kbt)
    pmapped_traj = RATMap(tmap=pre_tmap)(aug_traj)
    pmapped_coord_map = _noise_site_map(pmapped_traj.n_sites, aug_traj.n_aug_sites)
    # constraints have been mapped away by the pre-map (jgauss.py:259-264)
    pmapped_tmap = qp_linear_map(traj=pmapped_traj, coord_map=pmapped_coord_map, constraints=set(), **kwargs)
    pmapped_augmenter = CondNormal(cov=var, source_postmap=(pre_tmap.force_map @ pre_tmap.coord_map.T), seed=seed)
    post_tmap = AugmentedTMap(aug_tmap=pmapped_tmap, augmenter=pmapped_augmenter, kbt=kbt)
    return ComposedTMap(submaps=[post_tmap, pre_tmap])


def stagedjslicegauss_map(
    traj: CoordsTrajectory,
    coord_map: LinearMap,
    var: float,
    kbt: float,
    seed: Optional[int] = None,
    constraints: Optional[Constraints] = None,  # noqa: ARG001
    warn_input_forces: bool = True,
    noise=None,
) -> ComposedTMap:
    """Gaussian map whose reported forces come from the noise alone (reference ``jgauss.py:315-446``).

    Three stages: ``result[2]`` attaches NaN forces (so force-free input works), ``result[1]``
    maps the coordinates, ``result[0]`` noises them and reports the noise-site forces.
    """
    naforce_traj = NullForcesTMap(warn_input_forces=warn_input_forces)(traj)
    augmenter = CondNormal(cov=var, premap=coord_map, seed=seed, noise=noise)
    aug_traj = AugmentedTrajectory.from_trajectory(t=naforce_traj, augmenter=augmenter, kbt=kbt)
    null_fmap = LinearMap(mapping=np.ones_like(np.asarray(coord_map.standard_matrix)), handle_nans=False)
    pre_tmap = SeperableTMap(coord_map=coord_map, force_map=null_fmap)
    pmapped_traj = RATMap(tmap=pre_tmap)(aug_traj)
    pmapped_coord_map = _noise_site_map(pmapped_traj.n_sites, aug_traj.n_aug_sites)
    pmapped_tmap = constraint_aware_uni_map(traj=pmapped_traj, coord_map=pmapped_coord_map, constraints=set())
    post_tmap = AugmentedTMap(aug_tmap=pmapped_tmap, augmenter=CondNormal(cov=var, seed=seed), kbt=kbt)
    return ComposedTMap(submaps=[post_tmap, pre_tmap, NullForcesTMap(warn_input_forces=False)])


def stagedjforcegauss_map(
    traj: Trajectory,
    coord_map: LinearMap,
    var: float,
    kbt: float,
    force_map: Optional[LinearMap] = None,
    constraints: Optional[Constraints] = None,
    seed: Optional[int] = None,
    premap_l2_regularization: float = 0.0,
    premap_solver_args: SolverOptions = DEFAULT_SOLVER_OPTIONS,
    contribution_tolerance: float = 1e-6,
    noise=None,
    **kwargs,
) -> ComposedTMap:
    """Gaussian map that lets as little noise-derived force as possible into the mapped forces
    (reference ``jgauss.py:449-650``): as ``stagedjoptgauss_map``, but the second map is optimised
    on a trajectory whose REAL forces are zeroed, so only noise contributions are minimised; warns
    when their mean square stays above ``contribution_tolerance``.
    """
    pre_tmap = _pre_tmap(traj, coord_map, force_map, constraints, premap_l2_regularization, premap_solver_args)
    augmenter = CondNormal(cov=var, premap=pre_tmap.coord_map, seed=seed, noise=noise)
    zeros = torch.zeros_like(traj.forces) if isinstance(traj.forces, torch.Tensor) else np.zeros_like(traj.forces)
    aug_traj = AugmentedTrajectory.from_trajectory(t=Trajectory(coords=traj.coords, forces=zeros), augmenter=augmenter,
                                                   kbt=kbt)
    pmapped_traj = RATMap(tmap=pre_tmap)(aug_traj)
    pmapped_coord_map = _noise_site_map(pmapped_traj.n_sites, aug_traj.n_aug_sites)
    pmapped_tmap = qp_linear_map(traj=pmapped_traj, coord_map=pmapped_coord_map, constraints=set(), **kwargs)
    mapped = pmapped_tmap(pmapped_traj).forces
    remaining = float((mapped.double() ** 2).mean().item()) if isinstance(mapped, torch.Tensor) else float(
        np.mean(np.asarray(mapped) ** 2))
    if remaining > contribution_tolerance:
        warnings.warn(f"Unable to remove all noise contributions in forces. Remaining contribution: {remaining}.",
                      stacklevel=0)
    pmapped_augmenter = CondNormal(cov=var, source_postmap=(pre_tmap.force_map @ pre_tmap.coord_map.T), seed=seed)
    post_tmap = AugmentedTMap(aug_tmap=pmapped_tmap, augmenter=pmapped_augmenter, kbt=kbt)
    return ComposedTMap(submaps=[post_tmap, pre_tmap])
