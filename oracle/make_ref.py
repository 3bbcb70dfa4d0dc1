"""Recipe for ``oracle/_ref``: a runnable copy of the reference's own Python package.

TEST / BENCH INFRASTRUCTURE ONLY.  ``python oracle/make_ref.py`` (called by
``__graft_entry__.build()`` when ``/root/reference`` is present, i.e. in the build container)
copies ``/root/reference/src/aggforce`` UNMODIFIED into ``oracle/_ref/aggforce``.  ``oracle/_ref/``
is git-ignored (reference sources never enter the history) but travels to the GPU box with the
snapshot, where ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg time the
reference's own ``guess_pairwise_constraints`` / ``project_forces`` on the host cores.

The reference imports ``qpsolvers`` (absent here, unpinned in its ``setup.cfg``); ``load()`` installs
``oracle/qpsolvers_shim.py`` -- the exact equality-QP solve -- under that name before importing it.
Its JAX-guarded modules drop out on import (``try/except ImportError`` in the reference's own
``__init__`` files); the timed path is numpy only.
"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/src/aggforce")
DEST = HERE / "_ref"


def make() -> Path | None:
    """Copy the reference package; returns the destination or ``None`` when the mount is absent."""
    if not REF_SRC.is_dir():
        return None
    if DEST.exists():
        shutil.rmtree(DEST)
    DEST.mkdir(parents=True)
    shutil.copytree(REF_SRC, DEST / "aggforce", ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    (DEST / "README").write_text(
        "Unmodified copy of /root/reference/src/aggforce made by oracle/make_ref.py; git-ignored, not product.\n")
    return DEST


def available() -> bool:
    return (DEST / "aggforce" / "__init__.py").exists()


def load():
    """Import the reference package from ``oracle/_ref`` (behind the qpsolvers shim); ``None`` if absent."""
    if not available():
        return None
    if "qpsolvers" not in sys.modules:
        from . import qpsolvers_shim

        sys.modules["qpsolvers"] = qpsolvers_shim
    if str(DEST) not in sys.path:
        sys.path.insert(0, str(DEST))
    import aggforce  # noqa: PLC0415  (the reference)

    assert str(DEST) in str(Path(aggforce.__file__).resolve()), aggforce.__file__
    return aggforce


if __name__ == "__main__":
    print(make())
