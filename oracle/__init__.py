"""CPU oracle for the aggforce hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in ``aggforce_b200`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker or the timed CPU baseline.

Parity status (see DESIGN.md "Oracle"):
  * kernels (a) linear Gram, (c) pair moments, (d) map apply, the uniform map, the
    id-feature Gram and the equality-QP solve are PINNED: ``tests/golden/make_golden.py``
    ran the unmodified reference (``/root/reference/src`` behind a ``qpsolvers`` shim)
    and the committed fixtures under ``tests/golden/`` hold its outputs.
  * ``gb_feat``, ``JCondNormal`` / ``joptgauss_map`` and the ``jaxmapval`` projections (the JAX half)
    are PINNED to the reference's own modules executed unmodified behind a torch-backed ``jax``
    stand-in (``tests/golden/make_golden_jax.py``, ``tests/golden/jax_shim``; real jax/jaxlib are
    not installable here): fixtures ``ref_gbfeat.npz``, ``ref_jcondnormal.npz``, ``ref_mapval.npz``.
    The stand-in runs float32 on torch CPU, so those comparisons carry a float32 tolerance, and the
    reference's threefry noise stream is replaced by recorded draws.
"""
from .ref_numpy import *  # noqa: F401,F403
