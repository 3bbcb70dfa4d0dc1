"""Maps between resolutions."""
from .core import LinearMap, CLAMap, trjdot  # noqa: F401
from .tmap import (  # noqa: F401
    TMap,
    SeperableTMap,
    CLAFTMap,
    AugmentedTMap,
    ComposedTMap,
    NullForcesTMap,
    RATMap,
)
from .tools import lmap_augvariables, smear_map  # noqa: F401
