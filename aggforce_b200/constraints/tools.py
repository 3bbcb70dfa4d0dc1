"""Set algebra on molecular constraints (host side; tiny, never on the hot path).

Same results as the reference's ``src/aggforce/constraints/tools.py``:
``reduce_constraint_sets`` (:7-77) and ``constraint_lookup_dict`` (:80-116).
"""
from __future__ import annotations

from typing import Dict, Iterable, List

from .hints import Constraints


def _components(sets: List[frozenset]) -> List[int]:
    """Connected-component id of every input set (sets sharing a member are connected)."""
    parent = list(range(len(sets)))

    def find(i: int) -> int:
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    owner: Dict[int, int] = {}
    for idx, grp in enumerate(sets):
        for member in grp:
            if member in owner:
                a, b = find(owner[member]), find(idx)
                if a != b:
                    parent[b] = a
            else:
                owner[member] = idx
    return [find(i) for i in range(len(sets))]


def reduce_constraint_sets(constraints: Constraints) -> Constraints:
    """Merge overlapping constraint sets transitively into disjoint frozensets.

    ``{{1,2},{2,3},{4,5}} -> {{1,2,3},{4,5}}``.

    The partition is computed with a union-find.  The *iteration order* of the returned
    Python set is observable downstream (it fixes the id-feature label order, reference
    ``featlinearmap.py:602``), so the result set is populated the way the
    reference populates its own: seeds are drawn with ``set.pop()`` from a copy of the input
    and each merged group is inserted when its seed comes up.  ``tests/test_sets.py`` checks the
    resulting label vectors against the reference's on random inputs.
    """
    pool = set(constraints)
    if len(constraints) <= 1:
        return pool
    members = list(pool)
    comp = _components([frozenset(m) for m in members])
    merged: Dict[int, frozenset] = {}
    parts: Dict[int, List] = {}
    for m, cid in zip(members, comp):
        parts.setdefault(cid, []).append(m)
    for cid, items in parts.items():
        merged[cid] = frozenset().union(*items)
    comp_of = {m: cid for m, cid in zip(members, comp)}
    out: set = set()
    while pool:
        seed = pool.pop()
        cid = comp_of[seed]
        pool.difference_update(parts[cid])
        out.add(merged[cid])
    return out


def constraint_lookup_dict(constraints: Iterable[Iterable[int]]) -> Dict[int, int]:
    """Map every non-anchor member of each group to the group's smallest member."""
    lookup: Dict[int, int] = {}
    for grp in constraints:
        ordered = sorted(grp)
        for member in ordered[1:]:
            lookup[member] = ordered[0]
    return lookup
