"""Synthetic trajectories of the shapes named in BASELINE.json (benchmarks and tests).

No dataset of the reference is redistributable at scale (its cln025 trajectory is missing
from the reference checkout altogether), so benchmarks replicate the *shapes*: chignolin
(cln025: 175 atoms, 10 C-alpha beads, 78 X-H bonds -> 97 constraint groups) and generic
"protein-like" systems of any size.  Model: see ``csrc/synth.cu``.

``synth_trajectory_host`` is the numpy generator used by CPU tests and fixtures;
``synth_trajectory_device`` runs the counter-based CUDA generator (any frame range of any
rank reproduces the same frames).  The two use different random streams by design.
"""
from __future__ import annotations

import ctypes as C
import json
from dataclasses import dataclass
from pathlib import Path
from typing import List, Set, Tuple

import numpy as np

_DATA = Path(__file__).resolve().parent / "data"


@dataclass
class Topology:
    n_sites: int
    ref_pos: np.ndarray        # (n_sites, 3) float32
    parent: np.ndarray         # (n_sites,) int32, -1 for heavy atoms, else the bonded heavy atom
    bond_len: np.ndarray       # (n_sites,) float32 (0 for heavy atoms)
    bead_atoms: List[int]      # one fine-grained site per coarse-grained bead (slice map)

    @property
    def xh_constraints(self) -> Set[frozenset]:
        return {frozenset((int(p), int(h))) for h, p in enumerate(self.parent) if p >= 0}


def chignolin_topology() -> Topology:
    """cln025 topology derived from the reference's ``tests/data/cln025.pdb``."""
    topo = json.loads((_DATA / "cln025_topology.json").read_text())
    n = topo["n_atoms"]
    pos = np.asarray(topo["positions_angstrom"], dtype=np.float32)
    parent = np.full(n, -1, dtype=np.int32)
    for heavy, h in topo["xh_pairs"]:
        parent[h] = heavy
    bond = np.zeros(n, dtype=np.float32)
    hs = np.nonzero(parent >= 0)[0]
    bond[hs] = np.linalg.norm(pos[hs] - pos[parent[hs]], axis=1)
    return Topology(n, pos, parent, bond, list(topo["ca_indices"]))


def protein_like_topology(n_residues: int, atoms_per_residue: int = 10, heavy_per_residue: float = 5.2,
                          box: float = 60.0, seed: int = 1) -> Topology:
    """Generic system: ``n_residues`` residues of ``atoms_per_residue`` atoms; the first
    ``round(heavy_per_residue * n_residues)`` ... heavy atoms are spread so that exactly
    ``round(n_residues * heavy_per_residue)`` atoms are heavy and every hydrogen is bonded to a
    heavy atom of its own residue; the bead of a residue is its first heavy atom.
    (5000 atoms = 500 x 10 with 2600 heavy / 2400 H.)"""
    rng = np.random.default_rng(seed)
    n = n_residues * atoms_per_residue
    n_heavy_total = int(round(n_residues * heavy_per_residue))
    base, extra = divmod(n_heavy_total, n_residues)
    pos = np.zeros((n, 3), dtype=np.float32)
    parent = np.full(n, -1, dtype=np.int32)
    bond = np.zeros(n, dtype=np.float32)
    beads = []
    for r in range(n_residues):
        n_heavy = base + (1 if r < extra else 0)
        first = r * atoms_per_residue
        centre = rng.uniform(0.0, box, size=3)
        beads.append(first)
        for k in range(atoms_per_residue):
            a = first + k
            if k < n_heavy:
                pos[a] = centre + rng.normal(0.0, 1.5, size=3)
            else:
                par = first + (k % n_heavy)
                v = rng.normal(size=3)
                v /= np.linalg.norm(v)
                parent[a] = par
                bond[a] = 1.09
                pos[a] = pos[par] + 1.09 * v
    return Topology(n, pos, parent, bond, beads)


POS_SIGMA, FORCE_SIGMA, H_COUPLING, H_WOBBLE = 0.3, 300.0, 0.8, 0.35


def synth_trajectory_host(topo: Topology, n_frames: int, seed: int = 1234) -> Tuple[np.ndarray, np.ndarray]:
    """(coords, forces) float32 ``(n_frames, n_sites, 3)`` on the host (numpy generator)."""
    rng = np.random.default_rng(seed)
    n = topo.n_sites
    heavy = topo.parent < 0
    hs = np.nonzero(~heavy)[0]
    coords = np.empty((n_frames, n, 3), dtype=np.float32)
    jitter = rng.normal(0.0, POS_SIGMA, size=(n_frames, n, 3)).astype(np.float32)
    coords[:, heavy] = topo.ref_pos[heavy] + jitter[:, heavy]
    direction = topo.ref_pos[hs] - topo.ref_pos[topo.parent[hs]]
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    wob = direction[None] + H_WOBBLE * rng.normal(size=(n_frames, hs.size, 3))
    wob /= np.linalg.norm(wob, axis=2, keepdims=True)
    coords[:, hs] = (coords[:, topo.parent[hs]].astype(np.float64) + topo.bond_len[hs][None, :, None] * wob).astype(
        np.float32
    )
    forces = rng.normal(0.0, FORCE_SIGMA, size=(n_frames, n, 3)).astype(np.float32)
    forces[:, hs] -= np.float32(H_COUPLING) * forces[:, topo.parent[hs]]
    return coords, forces


def synth_trajectory_device(topo: Topology, n_frames: int, seed: int = 1234, frame0: int = 0,
                            want_coords: bool = True, want_forces: bool = True):
    """(coords, forces) float32 CUDA tensors for global frames [frame0, frame0 + n_frames)."""
    import torch

    from . import _engine, _lib

    dev = _engine.device()
    ref = torch.as_tensor(topo.ref_pos, device=dev).contiguous()
    par = torch.as_tensor(topo.parent, device=dev).contiguous()
    bl = torch.as_tensor(topo.bond_len, device=dev).contiguous()
    coords = torch.empty((n_frames, topo.n_sites, 3), dtype=torch.float32, device=dev) if want_coords else None
    forces = torch.empty((n_frames, topo.n_sites, 3), dtype=torch.float32, device=dev) if want_forces else None
    _lib.call("agf_synth_frames", _engine.ptr(ref), _engine.ptr(par), _engine.ptr(bl), topo.n_sites, int(frame0),
              int(n_frames), C.c_uint64(seed), POS_SIGMA, FORCE_SIGMA, H_COUPLING, _engine.ptr(coords),
              _engine.ptr(forces), _engine.stream_ptr())
    return coords, forces


class SynthFrames:
    """Virtual ``(n_frames, n_sites, 3)`` frame source backed by the counter-based generator: slabs of
    frames are produced by ``agf_synth_frames`` as a kernel asks for them and never stored, so a
    rank's share of config 4 (1.25 M frames x 5 000 atoms: 75 GB per array) streams through HBM once
    per pass.  Behaves like ``_engine.Frames`` (use ``make_synth_frames``)."""


def make_synth_frames(topo: Topology, n_frames: int, which: str, seed: int = 1234, frame0: int = 0,
                      slab_bytes: int = 1 << 30):
    """``which``: ``"coords"`` or ``"forces"``.  Global frames ``[frame0, frame0 + n_frames)``."""
    import numpy as _np
    import torch

    from . import _engine

    assert which in ("coords", "forces")

    class _SynthFrames(_engine.Frames, SynthFrames):
        def __init__(self, source=None) -> None:
            if "n_frames" in self.__dict__:  # _engine.Frames(x) passes an existing instance through
                return
            self._dev, self._host = None, None
            self.n_frames, self.n_sites = int(n_frames), int(topo.n_sites)

        @property
        def on_host(self) -> bool:
            return False

        @property
        def np_dtype(self):
            return _np.float32

        def pieces(self, start: int = 0, stop=None):
            stop = self.n_frames if stop is None else min(stop, self.n_frames)
            per = max(4, (slab_bytes // (self.n_sites * 12)) // 4 * 4)
            for a in range(start, stop, per):
                b = min(a + per, stop)
                c, f = synth_trajectory_device(topo, b - a, seed=seed, frame0=frame0 + a,
                                               want_coords=which == "coords", want_forces=which == "forces")
                yield a, (c if which == "coords" else f)

        def resident(self) -> "torch.Tensor":
            return torch.cat([p for _, p in self.pieces()])

        def prefix(self, n: int) -> "torch.Tensor":
            return torch.cat([p for _, p in self.pieces(0, min(n, self.n_frames))])

        def gather(self, frame_indices) -> "torch.Tensor":
            idx = _np.asarray(frame_indices, dtype=_np.int64)
            rows = [synth_trajectory_device(topo, 1, seed=seed, frame0=frame0 + int(i), want_coords=which == "coords",
                                            want_forces=which == "forces")[0 if which == "coords" else 1] for i in idx]
            return torch.cat(rows)

    return _SynthFrames()
