"""aggforce_b200 -- B200-native force-aggregation maps for coarse-graining MD trajectories.

Keeps the public names of ``noegroup/aggforce`` (``project_forces``, ``LinearMap``,
``guess_pairwise_constraints``, ``qp_linear_map``, ``constraint_aware_uni_map``, ``Trajectory``;
featurised and Gaussian maps under ``aggforce_b200.qp``) and runs every per-frame reduction in
hand-written sm_100a kernels (``csrc/``) behind the C ABI in ``include/agf_b200.h``.
"""
from .trajectory import Trajectory  # noqa: F401
from .agg import project_forces, project_forces_grid_cv  # noqa: F401
from .constraints import guess_pairwise_constraints  # noqa: F401
from .qp import (  # noqa: F401
    qp_linear_map,
    constraint_aware_uni_map,
    joptgauss_map,
    stagedjoptgauss_map,
    stagedjslicegauss_map,
    stagedjforcegauss_map,
)
from .map import LinearMap  # noqa: F401
from ._engine import frame_sharding, Frames  # noqa: F401
