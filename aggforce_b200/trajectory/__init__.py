"""Trajectory containers and augmenters."""
from .core import ForcesTrajectory, CoordsTrajectory, Trajectory, AugmentedTrajectory  # noqa: F401
from .augment import Augmenter  # noqa: F401
from .gausstraj import SimpleCondNormal, CondNormal, JCondNormal  # noqa: F401
