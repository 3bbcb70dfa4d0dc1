"""Tools and definitions related to molecular constraints."""
from .hints import Constraints  # noqa: F401
from .constfinder import guess_pairwise_constraints  # noqa: F401
from .tools import reduce_constraint_sets, constraint_lookup_dict, merged_groups  # noqa: F401
