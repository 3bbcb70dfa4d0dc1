"""CPU oracle for the aggforce hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in ``aggforce_b200`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker or the timed CPU baseline.

Parity status (see DESIGN.md "Oracle"):
  * kernels (a) linear Gram, (c) pair moments, (d) map apply, the uniform map, the
    id-feature Gram and the equality-QP solve are PINNED: ``tests/golden/make_golden.py``
    ran the unmodified reference (``/root/reference/src`` behind a ``qpsolvers`` shim)
    and the committed fixtures under ``tests/golden/`` hold its outputs.
  * ``gb_feat`` (JAX) and ``joptgauss_map`` (JAX) could not be executed in the build
    container (no jax): for those two the oracle is a line-by-line restatement whose
    derivative is checked against finite differences -- **parity unpinned**.
"""
from .ref_numpy import *  # noqa: F401,F403
