"""Optimal static linear force map -- kernel (a) plus the host QP.

Drop-in for the reference's ``src/aggforce/qp/qplinear.py``.
"""
from __future__ import annotations

import functools
from typing import Union

import numpy as np
import scipy.sparse as ss
import torch

from .. import _engine
from ..constraints import Constraints, constraint_lookup_dict, merged_groups
from ..map import LinearMap, SeperableTMap
from ..trajectory import ForcesTrajectory
from .solver import DEFAULT_SOLVER_OPTIONS, SolverOptions, solve, solve_equality_qp_device


def reduced_columns(n_sites: int, constraints: Constraints) -> np.ndarray:
    """Reduced-coefficient column of every site (``int64[n_sites]``).

    Sites constrained together share one coefficient.  Column order follows the reference
    (``qplinear.py:157-163``): walking the sites in increasing order, every site that is not a
    dependent member of a merged constraint group opens the next column; dependents reuse
    their anchor's (the smallest index of their group).  Pure index bookkeeping: memoised on the
    constraint set (a fit, the uniform map and the CV folds all ask for the same table).
    """
    try:
        key = frozenset(constraints)
    except TypeError:  # unhashable members: compute directly
        return _reduced_columns(n_sites, constraints)
    return _reduced_columns_cached(n_sites, key).copy()


@functools.lru_cache(maxsize=16)
def _reduced_columns_cached(n_sites: int, constraints: frozenset) -> np.ndarray:
    cols = _reduced_columns(n_sites, constraints)
    cols.setflags(write=False)
    return cols


def _reduced_columns(n_sites: int, constraints: Constraints) -> np.ndarray:
    anchor_of = constraint_lookup_dict(merged_groups(constraints))
    cols = np.full(n_sites, -1, dtype=np.int64)
    free = [s for s in range(n_sites) if s not in anchor_of]
    cols[free] = np.arange(len(free))
    for site, anchor in anchor_of.items():
        cols[site] = cols[anchor]
    return cols


def make_bond_constraint_matrix(n_sites: int, constraints: Constraints) -> np.ndarray:
    """One-hot ``(n_sites, n_reduced)`` matrix expanding reduced coefficients to all sites."""
    cols = reduced_columns(n_sites, constraints)
    mat = np.zeros((n_sites, int(cols.max()) + 1 if n_sites else 0))
    mat[np.arange(n_sites), cols] = 1
    return mat


def qp_form(target: np.ndarray) -> np.ndarray:
    """``(n_steps, n_sites, 3) -> (n_steps*3, n_sites)`` with rows ordered (step, dim)."""
    mixed = np.swapaxes(target, 1, 2)
    return np.reshape(mixed, (mixed.shape[0] * mixed.shape[1], -1))


def force_gram(forces, n_sites: int, constraints: Constraints):
    """``(gram (n_red, n_red) float64 on the host, columns)`` -- the QP objective of
    ``qplinear.py:66-71``, accumulated on the GPU and all-reduced across frame shards."""
    cols = reduced_columns(n_sites, constraints)
    n_red = int(cols.max()) + 1
    gram = _engine.gram_linear(_engine.Frames(forces), cols, n_red)
    return _engine.to_host_overlapped(gram), cols


# reduced problems at least this large are solved on the device (cuSOLVER through torch.linalg);
# below it the host solve is faster than the launch + synchronisation overhead
_DEVICE_SOLVE_MIN = 512


def qp_linear_map(
    traj: ForcesTrajectory,
    coord_map: LinearMap,
    constraints: Union[None, Constraints] = None,
    l2_regularization: float = 0.0,
    solver_args: SolverOptions = DEFAULT_SOLVER_OPTIONS,
) -> SeperableTMap:
    """Linear force map minimising the mean squared mapped force.

    Same contract as the reference: constrained sites share coefficients, the map satisfies
    ``coord_map @ W.T = I`` on the reduced coefficients, ``l2_regularization`` penalises the
    expanded coefficient vector and is relative to the *unnormalised* frame sum.
    """
    if constraints is None:
        constraints = set()
    n_fg = coord_map.n_fg_sites
    if traj.forces.shape[1] != n_fg:
        raise ValueError("coord_map and forces disagree on the number of fine-grained sites.")
    backend = dict(solver_args or {}).get("backend", "exact")
    cols = reduced_columns(n_fg, constraints)
    n_red = int(cols.max()) + 1
    on_device = backend == "exact" and n_red >= _DEVICE_SOLVE_MIN
    if on_device:
        qp_mat = _engine.gram_linear(_engine.Frames(traj.forces), cols, n_red)  # stays on the device
        _engine.run_deferred()
    else:
        qp_mat, cols = force_gram(traj.forces, n_fg, constraints)
    group_size = np.bincount(cols, minlength=n_red).astype(np.float64)
    if l2_regularization > 0.0:  # l2 * C'C
        if on_device:
            qp_mat.diagonal().add_(torch.as_tensor(l2_regularization * group_size, device=qp_mat.device))
        else:
            qp_mat[np.diag_indices(n_red)] += l2_regularization * group_size
    # A = coord_map @ C : sum the coordinate-map columns of every group
    cmat = np.asarray(coord_map.standard_matrix, dtype=np.float64)
    if n_fg * coord_map.n_cg_sites <= (1 << 16):  # small: a scatter-add beats building a sparse one-hot
        a_mat = np.zeros((coord_map.n_cg_sites, n_red))
        np.add.at(a_mat.T, cols, cmat.T)
    else:
        onehot = ss.csr_matrix((np.ones(n_fg), (np.arange(n_fg), cols)), shape=(n_fg, n_red))
        a_mat = np.asarray((onehot.T @ cmat.T).T)
    if backend == "exact":
        sol = None
        if on_device:
            sol = solve_equality_qp_device(qp_mat, a_mat, np.eye(coord_map.n_cg_sites))
            if sol is None:  # not numerically positive definite: host path with its null-space fallback
                qp_mat = _engine.to_host(qp_mat)
        if sol is None:
            sol = solve(qp_mat, a_mat, np.eye(coord_map.n_cg_sites), solver_args)
        if sol is None:
            raise ValueError("Map optimization failed.")
        reduced = sol.T
    else:
        rows = []
        for bead in range(coord_map.n_cg_sites):
            target = np.zeros(coord_map.n_cg_sites)
            target[bead] = 1
            x = solve(qp_mat, a_mat, target, solver_args)
            if x is None:
                raise ValueError("Map optimization failed.")
            rows.append(x)
        reduced = np.stack(rows)
    force_map = LinearMap(reduced[:, cols])
    force_map._column_labels = cols  # sites of one reduced column share their matrix column by construction
    return SeperableTMap(coord_map=coord_map, force_map=force_map)
