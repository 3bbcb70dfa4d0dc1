"""Quadratic-programming based force-map optimisation."""
from .qplinear import qp_linear_map, qp_form, make_bond_constraint_matrix  # noqa: F401
from .basicagg import constraint_aware_uni_map  # noqa: F401
from .solver import DEFAULT_SOLVER_OPTIONS, solve_equality_qp  # noqa: F401
from .featlinearmap import (  # noqa: F401
    FeatZipper,
    Multifeaturize,
    GeneralizedFeatures,
    GeneralizedFeaturizer,
    multifeaturize,
    qp_feat_linear_map,
    id_feat,
)
from .gbfeat import gb_feat, GbSpec  # noqa: F401
from .jgauss import (  # noqa: F401
    joptgauss_map,
    stagedjforcegauss_map,
    stagedjoptgauss_map,
    stagedjslicegauss_map,
)
