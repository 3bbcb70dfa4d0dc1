"""Constraint inference from pair-distance fluctuations -- kernel (c).

Drop-in for ``src/aggforce/constraints/constfinder.py:14-57`` of the reference.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .. import _engine
from .hints import Constraints


def guess_pairwise_constraints(xyz, cross_xyz=None, threshold: float = 1e-3) -> Constraints:
    """Find pairs of sites whose distance standard deviation over frames is below ``threshold``.

    Arguments and return value as in the reference: ``xyz`` is ``(n_steps, n_sites, 3)``;
    without ``cross_xyz`` the result is a set of 2-member frozensets, with ``cross_xyz`` a set
    of ordered tuples ``(i over cross_xyz, j over xyz)``.

    The statistics are evaluated in float64 on the GPU (the reference evaluates them in the
    input dtype; for float32 input that differs only when a pair sits within float32
    rounding of the threshold or the trajectory is long enough for float32 accumulation to
    drift).  All pairs are screened on a short prefix of frames; pairs whose partial
    sum of squared deviations already exceeds ``threshold**2 * n_steps`` can never qualify and
    are dropped, the survivors are streamed over every frame.
    """
    x = _engine.Frames(xyz)
    o = None if cross_xyz is None else _engine.Frames(cross_xyz)
    pairs, sd = _engine.pair_constraints(x, o, threshold)
    with np.errstate(invalid="ignore"):
        hit = sd < threshold
    found = pairs[hit].tolist()  # Python ints in one pass
    if o is None:
        return {frozenset(p) for p in found}
    return {(p[0], p[1]) for p in found}
