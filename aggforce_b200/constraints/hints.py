"""Type of a constraint collection (mirrors src/aggforce/constraints/hints.py:7)."""
from typing import FrozenSet, Set

Constraints = Set[FrozenSet[int]]
