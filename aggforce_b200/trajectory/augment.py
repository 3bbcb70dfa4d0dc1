"""Interface of objects that stochastically extend a trajectory with new sites.

Mirrors the reference's ``src/aggforce/trajectory/augment.py:13-110``.
"""
from abc import ABC, abstractmethod
from typing import Any, Tuple, TypeVar

_T_Augmenter = TypeVar("_T_Augmenter", bound="Augmenter")


class Augmenter(ABC):
    """Samples augmenting sites ``y ~ g(.|x)`` and evaluates ``grad log g`` w.r.t. x and y."""

    @abstractmethod
    def __init__(self) -> None:
        """Initialize."""

    @abstractmethod
    def sample(self, source: Any) -> Any:
        """Augmenting coordinates ``(n_frames, n_new_sites, 3)`` for ``source`` coordinates."""

    @abstractmethod
    def log_gradient(self, source: Any, generated: Any) -> Tuple[Any, Any]:
        """``(d log g / d source, d log g / d generated)``."""

    @abstractmethod
    def astype(self: _T_Augmenter, *args, **kwargs) -> _T_Augmenter:
        """Instance producing arrays of the given dtype."""
