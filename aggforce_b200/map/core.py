"""Array maps: ``LinearMap`` (kernel (d)) and ``CLAMap``.

API of the reference's ``src/aggforce/map/core.py`` (LinearMap :46-317, CLAMap :320-430).
``LinearMap.__call__`` is one fused kernel launch: matrix product, NaN probe, NaN protocol.
Host (numpy) input returns numpy, CUDA tensors return CUDA tensors.
"""
from __future__ import annotations

import hashlib
from typing import Callable, Dict, Final, List, Literal, Optional, Tuple, Union

import numpy as np
import torch

from .. import _engine
from ..util import trjdot


_COMPILED_BY_CONTENT: Dict[Tuple[bytes, int], "_engine.CompiledMap"] = {}


class _Taggable:
    """Holds a free-form ``tags`` dictionary (fit logs such as feature coefficients)."""

    def __init__(self, tags: Union[None, Dict[str, str]]) -> None:
        self.tags = {} if tags is None else tags


_WEIGHTS: dict = {}


def _fingerprint(m: np.ndarray) -> bytes:
    """Change detection for a live matrix.  Small ones (every call of a cln025-sized map): the interpreter's
    64-bit keyed hash of the bytes, a few microseconds.  Large ones (a 500 x 5 000 map is 20 MB; a 128-bit
    cryptographic digest of it costs 50 ms per call): the wrapping 64-bit sum of the words and their
    position-weighted sum -- any single edit changes the first, any exchange of unequal entries the second --
    about 7 ms."""
    c = np.ascontiguousarray(m)
    if c.nbytes <= (1 << 16):
        return hash(c.tobytes()).to_bytes(8, "little", signed=True)
    if c.nbytes % 8:
        return hashlib.blake2b(c.tobytes(), digest_size=16).digest()
    v = c.reshape(-1).view(np.uint64)
    w = _WEIGHTS.get(v.size)
    if w is None:
        if len(_WEIGHTS) > 4:
            _WEIGHTS.clear()
        w = _WEIGHTS[v.size] = np.arange(1, v.size + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        return int(v.sum()).to_bytes(8, "little") + int(np.dot(v, w)).to_bytes(8, "little")


class LinearMap:
    """Linear fine-grained -> coarse-grained map defined by its ``standard_matrix``.

    ``mapping`` is either a 2-D array ``(n_cg, n_fg)`` of coefficients, or a list of lists of
    fine-grained indices (each bead the unweighted mean of its sites; needs ``n_fg_sites``).

    ``handle_nans`` (True | False | "safe"): with NaN handling on, NaN entries of the input are
    ignored wherever the map gives them zero weight and a ``ValueError`` is raised if the
    result would depend on them (tolerance ``nan_check_threshold``).  Unlike the reference the
    caller's array is never modified, i.e. True behaves like "safe".
    """

    n_dim: Final = 3

    def __init__(
        self,
        mapping: Union[List[List[int]], np.ndarray],
        n_fg_sites: Union[int, None] = None,
        handle_nans: Union[bool, Literal["safe"]] = True,
        nan_check_threshold: float = 1e-6,
    ) -> None:
        if isinstance(mapping, torch.Tensor):
            mapping = mapping.detach().cpu().numpy()
        if isinstance(mapping, np.ndarray) and mapping.ndim == 2:
            if n_fg_sites is not None:
                raise ValueError("Cannot specify n_fg_sites when mapping is ArrayLike. Let it be inferred.")
            matrix = mapping
        elif hasattr(mapping, "__iter__"):
            if n_fg_sites is None:
                raise ValueError("n_fg_sites is required when mapping is a list of index lists.")
            groups = [list(g) for g in mapping]
            matrix = np.zeros((len(groups), n_fg_sites))
            for bead, members in enumerate(groups):
                row = np.zeros(n_fg_sites)
                row[members] = 1 / len(members)
                matrix[bead, :] = row
        else:
            raise ValueError(f"Cannot understand mapping {mapping}.")
        self._matrix: Optional[np.ndarray] = matrix
        self._shape: Tuple[int, int] = tuple(matrix.shape)  # type: ignore[assignment]
        self._pending = None  # a fit whose result still lives on the device (see from_device_fit)
        self.handle_nans = handle_nans
        if self.handle_nans and not np.all(np.isfinite(matrix)):
            raise ValueError("NaN checking can only be performed if standard_matrix is itself finite.")
        self.nan_check_threshold = nan_check_threshold
        self._compiled: Optional[Tuple[bytes, _engine.CompiledMap]] = None
        self._frozen_digest: Optional[bytes] = None
        self._column_labels: Optional[np.ndarray] = None  # set by fits that know the column structure

    @classmethod
    def from_device_fit(cls, pending, n_cg: int, column_labels: np.ndarray, compiled: "_engine.CompiledMap",
                        handle_nans: Union[bool, Literal["safe"]] = True,
                        nan_check_threshold: float = 1e-6) -> "LinearMap":
        """Map fitted ON THE DEVICE (``agf_qp_equality_small``): ``compiled`` already holds the
        coefficients kernel (d) reads, so the map can be applied without the host ever seeing them;
        ``standard_matrix`` is downloaded (and the solver's status checked) on first access through
        ``pending.matrix()``."""
        self = cls.__new__(cls)
        self._matrix = None
        self._shape = (int(n_cg), int(np.asarray(column_labels).size))
        self._pending = pending
        self.handle_nans = handle_nans
        self.nan_check_threshold = nan_check_threshold
        self._compiled = (b"device-fit", compiled)
        self._frozen_digest = None
        self._column_labels = np.asarray(column_labels)
        return self

    # ------------------------------------------------------------------ descriptors
    @property
    def _standard_matrix(self) -> np.ndarray:
        if self._matrix is None:  # first host access to a device fit: one synchronising read
            self._matrix = self._pending.matrix()
            self._adopt_device_fit()
        return self._matrix

    def _adopt_device_fit(self) -> None:
        """The host copy of a device fit has just been materialised: fingerprint it NOW, while it still equals
        the coefficients on the device, so that an in-place edit made before the next application is seen."""
        if self._matrix is not None and self._compiled is not None and self._compiled[0] == b"device-fit":
            m = self._matrix
            digest = _fingerprint(m) + repr((m.shape, str(m.dtype), bool(self.handle_nans))).encode()
            self._compiled = (digest, self._compiled[1])

    @_standard_matrix.setter
    def _standard_matrix(self, value: np.ndarray) -> None:
        self._matrix = value

    @property
    def standard_matrix(self) -> np.ndarray:
        return self._standard_matrix

    @property
    def n_cg_sites(self) -> int:
        return self._shape[0]

    @property
    def n_fg_sites(self) -> int:
        return self._shape[1]

    @property
    def participating_fg(self) -> List[List[int]]:
        """For every bead, the fine-grained sites with a positive coefficient."""
        table: List[List[int]] = [[] for _ in range(self.n_cg_sites)]
        for cg, fg in zip(*np.nonzero(self._standard_matrix > 0)):
            table[cg].append(fg)
        return table

    def close_to_identity(self, threshold: float = 1e-8) -> bool:
        m = self._standard_matrix
        if m.shape[0] != m.shape[1]:
            return False
        return bool(np.sqrt(((np.identity(m.shape[0], dtype=m.dtype) - m) ** 2).sum()) <= threshold)

    # ------------------------------------------------------------------ application
    def _compile(self) -> _engine.CompiledMap:
        if self._matrix is None:  # device fit nobody has looked at (or edited) yet
            return self._compiled[1]  # type: ignore[index]
        m = self._standard_matrix
        # ``standard_matrix`` hands out the live array and the reference re-reads it on every call
        # (core.py:240), so in-place edits must be seen: digest of the WHOLE matrix (microseconds at
        # cln025, 20 ms for a 500 x 5000 map), taken once only for arrays nobody can write to
        if m.flags.writeable or self._frozen_digest is None:
            digest = _fingerprint(m)
            if not m.flags.writeable:
                self._frozen_digest = digest
        else:
            digest = self._frozen_digest
        digest += repr((m.shape, str(m.dtype), bool(self.handle_nans))).encode()
        if self._compiled is not None and self._compiled[0] != digest:
            self._column_labels = None  # matrix was edited in place: the fit's column structure is stale
        if self._compiled is None or self._compiled[0] != digest:
            # maps with the same content (the coordinate map of every call, the uniform map rebuilt from
            # the same constraints) share ONE device form: compiled maps are immutable
            key = (digest, _engine.device().index)
            shared = _COMPILED_BY_CONTENT.get(key) if self._column_labels is None else None
            if shared is None:
                # plain mode keeps all-zero columns so that 0 * NaN = NaN exactly as numpy computes it
                shared = _engine.CompiledMap(self._standard_matrix, keep_zero_columns=not self.handle_nans,
                                             column_labels=self._column_labels)
                if self._column_labels is None:
                    if len(_COMPILED_BY_CONTENT) > 64:
                        _COMPILED_BY_CONTENT.clear()
                    _COMPILED_BY_CONTENT[key] = shared
            self._compiled = (digest, shared)
        return self._compiled[1]

    def _launch(self, points, want_sumsq: bool = False, status: Optional[torch.Tensor] = None,
                slots: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, download: Optional[list] = None):
        """Enqueue the kernel; returns ``(frames, out_dev, status_dev)`` without synchronising.
        ``status_dev`` is a float64[2] device tensor ``[sum(out**2), flags]`` whose second slot holds
        the two int32 flags ``(saw_nan, nan_violation)``.  ``slots = (sum_slot, flag_slot)`` -- two
        zeroed one-element float64 views of a larger buffer -- makes several launches report through
        ONE read (then ``status_dev`` is ``None``).  ``download``: see ``_engine.map_apply``."""
        frames = _engine.Frames(points)
        if slots is None:
            if status is None:
                status = torch.zeros(2, dtype=torch.float64, device=_engine.device())
            slots = (status[0:1], status[1:2])
        out, _, _ = _engine.map_apply(
            frames, self._compile(), nan_mode=1 if self.handle_nans else 0,
            nan_atol=self.nan_check_threshold, want_sumsq=want_sumsq,
            sumsq=slots[0] if want_sumsq else None, flags=slots[1].view(torch.int32), download=download,
        )
        return frames, out, status

    @staticmethod
    def _decode_flags(flag_slot_host: np.ndarray) -> Tuple[int, int]:
        """``(saw_nan, nan_violation)`` from a downloaded flag slot (one float64 holding two int32)."""
        flags = np.ascontiguousarray(flag_slot_host).reshape(-1)[:1].view(np.int32)
        return int(flags[0]), int(flags[1])

    def _finish(self, frames, out, flag_slot_host, host_copy=None):
        """``host_copy``: pinned tensor of an already started download (``_engine.start_d2h``)."""
        if self.handle_nans and self._decode_flags(flag_slot_host)[1] != 0:
            raise ValueError(
                "NaN handling is on and results seem to depend on NaN "
                "positions in input array. Check input and standard_matrix."
            )
        if not frames.on_host:
            return out
        if host_copy is not None:
            _engine.finish_d2h()
            return host_copy.numpy()
        return _engine.to_host(out)

    def _apply(self, points, want_sumsq: bool = False):
        dl: list = []
        frames, out, status = self._launch(points, want_sumsq, download=dl)
        if _engine.sharded():  # every rank must raise (or not) together: the next collective would hang
            _engine.allreduce_max_(status[1:2].view(torch.int32))
        host = _engine.to_host(status)  # one synchronising read: NaN flags and the residual sum together
        return self._finish(frames, out, host[1:2], dl[0] if dl else None), float(host[0])

    def __call__(self, points):
        """Map ``(n_steps, n_fg_sites, 3)`` points to ``(n_steps, n_cg_sites, 3)``."""
        return self._apply(points)[0]

    def apply_with_sumsq(self, points):
        """``(mapped, sum(mapped**2))`` from one kernel launch (the residual of agg.py:291-297)."""
        return self._apply(points, want_sumsq=True)

    def flat_call(self, flattened):
        """Apply to ``(n_frames, n_fg_sites*3)`` input, returning ``(n_frames, n_cg_sites*3)``."""
        shape = tuple(flattened.shape)
        if len(shape) == 3:
            raise ValueError(f"Expected array of rank 2; got array with shape {shape}.")
        if shape[1] % self.n_dim != 0:
            raise ValueError(f"Array of shape {shape} can't be reshaped with dim of {self.n_dim}.")
        mapped = self(flattened.reshape((shape[0], shape[1] // self.n_dim, self.n_dim)))
        return mapped.reshape((mapped.shape[0], mapped.shape[1] * mapped.shape[2]))

    # ------------------------------------------------------------------ algebra
    def _like(self, matrix: np.ndarray) -> "LinearMap":
        return self.__class__(mapping=matrix, handle_nans=self.handle_nans,
                              nan_check_threshold=self.nan_check_threshold)

    @property
    def T(self) -> "LinearMap":
        return LinearMap(mapping=self._standard_matrix.T, handle_nans=self.handle_nans,
                         nan_check_threshold=self.nan_check_threshold)

    def __matmul__(self, lm: "LinearMap", /) -> "LinearMap":
        return LinearMap(mapping=self._standard_matrix @ lm.standard_matrix, handle_nans=self.handle_nans,
                         nan_check_threshold=self.nan_check_threshold)

    def __rmul__(self, c: float, /) -> "LinearMap":
        return LinearMap(mapping=c * self._standard_matrix, handle_nans=self.handle_nans,
                         nan_check_threshold=self.nan_check_threshold)

    def __add__(self, lm: "LinearMap", /) -> "LinearMap":
        return LinearMap(mapping=self._standard_matrix + lm.standard_matrix, handle_nans=self.handle_nans,
                         nan_check_threshold=self.nan_check_threshold)

    def astype(self, *args, **kwargs) -> "LinearMap":
        """Instance whose matrix is cast with ``numpy.astype(*args, **kwargs)``."""
        return self._like(self._standard_matrix.astype(*args, **kwargs))


class CLAMap(_Taggable):
    """Co-local affine map ``x_t -> A(y_t) x_t + b(y_t)`` (output of featurised fits).

    ``scale(copoints) -> (n_steps, n_cg, n_fg)`` and ``trans(copoints) -> (n_steps, n_cg, 3)``
    are callables; with ``zeroes_check`` they are probed once on a zero frame to validate /
    infer ``n_cg_sites`` (reference core.py:381-394).
    """

    n_dim: Final = 3

    def __init__(
        self,
        scale: Callable,
        trans: Callable,
        n_fg_sites: int,
        n_cg_sites: Optional[int] = None,
        zeroes_check: bool = True,
        tags: Optional[Dict[str, str]] = None,
    ) -> None:
        super().__init__(tags=tags)
        if zeroes_check:
            probe = np.zeros((1, n_fg_sites, self.n_dim))
            mapped = trjdot(probe, scale(probe)) + trans(probe)
            if n_cg_sites is None:
                n_cg_sites = mapped.shape[1]
            elif n_cg_sites != mapped.shape[1]:
                raise ValueError("n_cg_sites did not match results from zero test")
        elif n_cg_sites is None:
            raise ValueError("If n_cg_sites is not set, zeroes_check must be truthy.")
        self._n_cg_sites: Final = n_cg_sites
        self._n_fg_sites: Final = n_fg_sites
        self.scale: Final = scale
        self.trans: Final = trans

    @property
    def n_cg_sites(self) -> int:
        return self._n_cg_sites

    @property
    def n_fg_sites(self) -> int:
        return self._n_fg_sites

    def __call__(self, points, copoints):
        return trjdot(points, self.scale(copoints)) + self.trans(copoints)
