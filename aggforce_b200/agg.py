"""Orchestration: fit a force map and apply it (reference ``src/aggforce/agg.py``).

``project_forces`` keeps the reference's signature and result dictionary.  Every per-frame
reduction behind it runs on the GPU:
  constraint detection (kernel c) -> fit (kernel a + host QP, or kernel b) -> apply coords and
  forces (kernel d, the residual ``mean(F_mapped**2)`` fused into the force application).
Under ``aggforce_b200.frame_sharding()`` the arrays are this rank's frame slice; the fit
all-reduces its accumulators and the residual is the global mean.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Final, Union

import numpy as np
import torch

from . import _engine
from .constraints import Constraints, guess_pairwise_constraints
from .map import LinearMap, SeperableTMap, TMap
from .qp import qp_linear_map
from .trajectory import Trajectory

PROJECT_FORCES_CNSTR_AUTO: Final = "auto"

SCORES_KNAME: Final = "scores"
SDS_KNAME: Final = "sds"
NRUNS_KNAME: Final = "n_runs"

PROJFORCES_KNAME: Final = "mapped_forces"
PROJCOORDS_KNAME: Final = "mapped_coords"
TMAP_KNAME: Final = "tmap"
RESIDUAL_KNAME: Final = "residual"
CONSTRAINTS_KNAME: Final = "constraints"


def force_smoothness(array) -> float:
    """Mean squared element of ``array`` (reference agg.py:291-297)."""
    if isinstance(array, torch.Tensor):
        return float((array.double() ** 2).mean().item())
    return float(np.mean(np.asarray(array) ** 2))


def _global_mean_sq(sumsq: float, shape) -> float:
    local = np.array([sumsq, float(np.prod(shape))])
    tot = np.sum(_engine.allgather_host(local), axis=0)
    return float(tot[0] / tot[1])


def project_forces(
    coords,
    forces,
    coord_map: LinearMap,
    constrained_inds: Union[Constraints, str, None] = PROJECT_FORCES_CNSTR_AUTO,
    method: Callable[..., TMap] = qp_linear_map,
    **kwargs,
) -> Dict[str, Any]:
    """Produce an optimised force map and the mapped trajectory.

    Arguments as in the reference: ``coords``/``forces`` of shape ``(n_steps, n_sites, 3)``
    (numpy arrays, or torch CUDA tensors for device-resident data), ``coord_map`` the
    configurational ``LinearMap``, ``constrained_inds`` a set of frozensets or ``"auto"``
    (``guess_pairwise_constraints`` on all of ``coords``), ``method`` any callable
    ``(traj, coord_map, constraints, **kwargs) -> TMap``.

    Returns a dict with keys ``mapped_coords``, ``mapped_forces``, ``tmap``, ``residual``
    (mean squared mapped force, not held out) and ``constraints``.
    """
    auto = isinstance(constrained_inds, str) and constrained_inds == PROJECT_FORCES_CNSTR_AUTO
    if auto and coords is None:
        raise ValueError(f"If constrained_inds is {PROJECT_FORCES_CNSTR_AUTO}, coords cannot be None.")
    t = Trajectory(coords=coords, forces=forces)
    # one device upload per array, shared by constraint detection, the fit and the application
    coords_in, forces_in = _engine.Frames(coords), _engine.Frames(forces)
    if auto:
        constrained_inds = guess_pairwise_constraints(coords_in)
    traj_map: TMap = method(
        traj=Trajectory(coords=_Shared(coords, coords_in), forces=_Shared(forces, forces_in)),
        coord_map=coord_map,
        constraints=constrained_inds,
        **kwargs,
    )
    if (isinstance(traj_map, SeperableTMap) and isinstance(traj_map.force_map, LinearMap)
            and isinstance(traj_map.coord_map, LinearMap)):
        # both applications are enqueued back to back; one read returns NaN flags + residual sum
        cm, fm = traj_map.coord_map, traj_map.force_map
        fc, oc, sc = cm._launch(coords_in)
        host_c = _engine.start_d2h(oc) if fc.on_host else None  # downloads overlap the force upload
        ff, of, sf = fm._launch(forces_in, want_sumsq=True)
        host_f = _engine.start_d2h(of) if ff.on_host else None
        # [flags_c(2), sumsq_c, flags_f(2), sumsq_f, n_elements]; under frame sharding the residual
        # numerator / denominator are summed over ranks on the device before the single read
        count = torch.tensor([float(np.prod(of.shape))], dtype=torch.float64, device=of.device)
        packed = torch.cat([sc, sf, count])
        if _engine.sharded():
            tail = packed[5:7].clone()
            _engine.allreduce_sum_(tail)
            packed = torch.cat([packed[:5], tail])
        status = packed.cpu().numpy()
        mapped_coords = cm._finish(fc, oc, status[0:3], host_c)
        mapped_forces = fm._finish(ff, of, status[3:6], host_f)
        residual = float(status[5] / status[6])
    else:
        mapped = traj_map(t)
        mapped_coords, mapped_forces = mapped.coords, mapped.forces
        residual = force_smoothness(mapped_forces)
    return {
        PROJCOORDS_KNAME: mapped_coords,
        PROJFORCES_KNAME: mapped_forces,
        TMAP_KNAME: traj_map,
        RESIDUAL_KNAME: residual,
        CONSTRAINTS_KNAME: constrained_inds,
    }


class _Shared:
    """Array stand-in handed to fit methods: exposes ``shape`` like the original array and
    carries the already-uploaded ``Frames`` so kernels do not upload it a second time."""

    def __init__(self, original, frames) -> None:
        self.original = original
        self.frames = frames

    @property
    def shape(self):
        return self.original.shape

    def __len__(self) -> int:
        return len(self.original)

    def __getattr__(self, name):
        return getattr(self.original, name)

    def __getitem__(self, idx):
        return self.original[idx]

    def __array__(self, *args, **kwargs):
        return np.asarray(self.original, *args, **kwargs)
