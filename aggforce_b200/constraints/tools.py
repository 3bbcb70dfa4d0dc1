"""Set algebra on molecular constraints (host side; tiny, never on the hot path).

Same results as the reference's ``src/aggforce/constraints/tools.py``:
``reduce_constraint_sets`` (:7-77) and ``constraint_lookup_dict`` (:80-116).
"""
from __future__ import annotations

import functools
from typing import Dict, Iterable, Tuple

from .hints import Constraints


def reduce_constraint_sets(constraints: Constraints) -> Constraints:
    """Merge overlapping constraint sets transitively into disjoint frozensets.

    ``{{1,2},{2,3},{4,5}} -> {{1,2,3},{4,5}}``.

    The partition is the set of connected components.  The *iteration order* of the returned
    Python set is observable downstream -- it fixes the id-feature label order (reference
    ``featlinearmap.py:602``) and with it which gb channel is dropped (Q5) -- and that
    order depends on CPython's set internals: ``set.pop()`` walks the hash table from a moving
    finger and ``difference_update`` rebuilds the table whenever more than a quarter of it is
    tombstones, even when called with nothing to remove.  To stay label-compatible the flood below
    therefore issues exactly the reference's sequence of set operations (one ``difference_update``
    per growth round, two empty ones before a component is closed, seeds drawn with ``pop``).
    ``tests/test_host_logic.py`` pins the resulting label vectors against the reference.
    """
    pool = set(constraints)
    if len(constraints) <= 1:
        return pool
    out: set = set()
    merged = frozenset(pool.pop())
    confirmed = False
    while True:
        hits = [grp for grp in pool if merged.intersection(grp)]
        merged = merged.union(*hits)
        pool.difference_update(hits)
        if hits:
            continue
        out.add(merged)
        if not confirmed:
            confirmed = True
            continue
        confirmed = False
        if not pool:
            return out
        merged = frozenset(pool.pop())


@functools.lru_cache(maxsize=64)
def _merged_groups_cached(key: frozenset) -> Tuple[Tuple[int, ...], ...]:
    parent: Dict[int, int] = {}

    def find(a: int) -> int:
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for grp in key:
        members = list(grp)
        for v in members:
            parent.setdefault(v, v)
        for v in members[1:]:
            ra, rb = find(members[0]), find(v)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
    comps: Dict[int, list] = {}
    for v in parent:
        comps.setdefault(find(v), []).append(v)
    return tuple(sorted((tuple(sorted(int(x) for x in c)) for c in comps.values()), key=lambda g: g[0]))


def merged_groups(constraints: Iterable[Iterable[int]]) -> Tuple[Tuple[int, ...], ...]:
    """The partition ``reduce_constraint_sets`` computes, as sorted tuples ordered by anchor.

    Union-find, memoised on the constraint set: the fit / map builders call this on every
    ``project_forces`` and only need the partition, not the reference's set iteration order.
    """
    return _merged_groups_cached(frozenset(frozenset(g) for g in constraints))


def constraint_lookup_dict(constraints: Iterable[Iterable[int]]) -> Dict[int, int]:
    """Map every non-anchor member of each group to the group's smallest member."""
    lookup: Dict[int, int] = {}
    for grp in constraints:
        ordered = sorted(grp)
        for member in ordered[1:]:
            lookup[member] = ordered[0]
    return lookup
