"""Gaussian augmenters.

``SimpleCondNormal``: isotropic noise on every site, identity pre-map (host numpy; reference
``src/aggforce/trajectory/simplegausstraj.py:13-137``).

``CondNormal``: ``y = A x + eps, eps ~ N(0, var I)`` with ``A`` a ``LinearMap`` -- the B200
replacement for the reference's JAX ``JCondNormal`` (``jaxgausstraj.py:99-402``), whose autodiff
log-gradients have the closed form ``grad_y = -(y - A x)/var``, ``grad_x = A^T (y - A x)/var``.
Sampling and both augmented arrays come from one fused kernel (``agf_gauss_augment``,
csrc/augment.cu); noise is drawn in-kernel (Philox keyed by seed / global frame / bead / draw id, so
frame slabs and ranks reproduce it) and cannot match JAX's threefry stream: parity tests inject
the draw through ``noise=``.
"""
from __future__ import annotations

from typing import Any, Optional, Tuple

import numpy as np
import torch

from .. import _engine
from .augment import Augmenter


class SimpleCondNormal(Augmenter):
    """Adds independent N(0, var) noise to every coordinate (identity pre-map)."""

    def __init__(self, var: float, seed: Optional[int] = None, dtype: Any = None) -> None:
        self.var = var
        self._rng = np.random.default_rng(seed)
        self.dtype = np.dtype(np.float32 if dtype is None else dtype)

    def sample(self, source: np.ndarray) -> np.ndarray:
        noise = np.sqrt(self.var) * self._rng.standard_normal(source.shape, dtype=self.dtype)
        return (source + noise).astype(self.dtype, copy=False)

    def log_gradient(self, source: np.ndarray, generated: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        wrt_generated = (-(1.0 / self.var) * (generated - source)).astype(self.dtype, copy=False)
        return -wrt_generated, wrt_generated

    def astype(self, dtype, *args, **kwargs) -> "SimpleCondNormal":  # noqa: ARG002
        return self.__class__(var=self.var, dtype=dtype)


class CondNormal(Augmenter):
    """``y = premap(x) + N(0, cov)`` with a linear ``premap``.  ``cov``: a scalar (isotropic noise: the fused
    kernel ``agf_gauss_augment``) or, as the reference's ``JCondNormal`` accepts, the full covariance matrix of
    the flattened ``(n_new * 3)`` noise vector (sites major, xyz fastest: ``jaxgausstraj.py:331-346``), which
    takes a dense device path (Cholesky factor for the draw, precision matrix for the log-gradients).

    ``premap`` is a ``LinearMap`` (or its bound ``flat_call``); ``None`` means identity.
    ``noise`` optionally injects the standard-normal draw ``(n_frames, n_new, 3)`` used by the
    next ``sample``/``augment`` call (test hook).
    """

    n_dim = 3

    def __init__(self, cov: float, premap=None, source_postmap=None, seed: Optional[int] = None,
                 dtype: Any = None, noise=None) -> None:
        self._cov_matrix = None
        if not np.isscalar(cov):
            mat = np.asarray(cov, dtype=np.float64)
            if mat.ndim != 2 or mat.shape[0] != mat.shape[1] or mat.shape[0] % self.n_dim:
                raise ValueError("cov must be a scalar or a square (n_new * 3, n_new * 3) matrix")
            self._cov_matrix = mat
            if dtype is None and isinstance(cov, np.ndarray) and cov.dtype in (np.float32, np.float64):
                dtype = cov.dtype  # the reference takes the working dtype from the matrix
        owner = getattr(premap, "__self__", None)
        self.premap = owner if owner is not None else premap
        # linear map applied to the log-gradient w.r.t. the source sites (jaxgausstraj.py:263-284):
        # the staged maps pass (force_map @ coord_map.T) so that the noise force on an ALREADY mapped
        # trajectory is the mapped back-projected noise force
        post_owner = getattr(source_postmap, "__self__", None)
        self.source_postmap = post_owner if post_owner is not None else source_postmap
        self.cov = float(cov) if self._cov_matrix is None else self._cov_matrix
        self.seed = int(np.random.default_rng().integers(0, 10**6)) if seed is None else int(seed)
        self.dtype = np.dtype(np.float32 if dtype is None else dtype)
        self._noise = noise

    # -- helpers
    def _tdtype(self) -> torch.dtype:
        return torch.float32 if self.dtype == np.float32 else torch.float64

    def _mean(self, source):
        return source if self.premap is None else self.premap(source)

    def _back(self, eps: torch.Tensor):
        """[source_postmap] A^T eps for the pre-map A (identity when premap is None)."""
        out = eps if self.premap is None else self.premap.T(eps)
        return out if self.source_postmap is None else self.source_postmap(out)

    # -- full covariance (dense device path)
    def _factors(self):
        """(Cholesky factor L, precision matrix) of the covariance as float64 device tensors, cached."""
        cached = getattr(self, "_cov_factors", None)
        if cached is None:
            sigma = torch.as_tensor(self._cov_matrix, device=_engine.device())
            chol, info = torch.linalg.cholesky_ex(sigma)
            if int(info.item()) != 0:
                raise ValueError("cov is not positive definite")
            prec = torch.cholesky_solve(torch.eye(sigma.shape[0], dtype=torch.float64, device=sigma.device), chol)
            cached = self._cov_factors = (chol, prec)
        return cached

    def _precision_times(self, resid: torch.Tensor) -> torch.Tensor:
        """``cov^-1 r`` for every frame (float64 arithmetic, result in the working dtype)."""
        _, prec = self._factors()
        flat = resid.reshape(resid.shape[0], -1).to(torch.float64)
        if flat.shape[1] != prec.shape[0]:
            raise ValueError(f"cov is {tuple(prec.shape)} but the augmenting sites have {flat.shape[1]} coordinates")
        return (flat @ prec).reshape(resid.shape).to(resid.dtype)

    _NOISE_BLOCK = 4096

    def _standard_normal(self, draw: "NoiseDraw", frame0: int, n_frames: int, n_new: int) -> torch.Tensor:
        """Standard-normal draw for global frames ``frame0 .. frame0 + n_frames``: generated in fixed blocks of
        frames, each from its own generator keyed by (seed, draw, block), so that any slab decomposition of the
        same draw sees the same noise."""
        dev = _engine.device()
        first, last = frame0 // self._NOISE_BLOCK, (frame0 + n_frames - 1) // self._NOISE_BLOCK
        parts = []
        for blk in range(first, last + 1):
            gen = torch.Generator(device=dev)
            gen.manual_seed((self.seed * 1_000_003 + draw.index * 7919 + blk) % (2**63 - 1))
            z = torch.randn((self._NOISE_BLOCK, n_new, self.n_dim), dtype=torch.float64, device=dev, generator=gen)
            lo = max(frame0 - blk * self._NOISE_BLOCK, 0)
            hi = min(frame0 + n_frames - blk * self._NOISE_BLOCK, self._NOISE_BLOCK)
            parts.append(z[lo:hi])
        return torch.cat(parts)

    def _augment_full_cov(self, coords, forces, kbt: float, draw: "NoiseDraw", frame0: int):
        td = self._tdtype()
        ref = coords if coords is not None else forces
        n_frames, n_sites = int(ref.shape[0]), int(ref.shape[1])
        n_new = self.n_new_sites(n_sites)
        chol, prec = self._factors()
        if chol.shape[0] != n_new * self.n_dim:
            raise ValueError(f"cov is {tuple(chol.shape)} but the augmenting sites have {n_new * self.n_dim} coordinates")
        if draw.noise is not None:
            z = draw.noise[frame0 : frame0 + n_frames].to(torch.float64)
            if tuple(z.shape) != (n_frames, n_new, self.n_dim):
                raise ValueError(f"injected noise has shape {tuple(draw.noise.shape)}; frames {frame0}.."
                                 f"{frame0 + n_frames} of (n_frames, {n_new}, 3) are needed")
        else:
            z = self._standard_normal(draw, frame0, n_frames, n_new)
        flat_z = z.reshape(n_frames, -1)
        eps = (flat_z @ chol.T).reshape(n_frames, n_new, self.n_dim)           # cov-distributed noise
        scaled = torch.linalg.solve_triangular(chol.T.contiguous(), flat_z.T.contiguous(), upper=True).T  # cov^-1 eps = L^-T z
        scaled = scaled.reshape(n_frames, n_new, self.n_dim)
        oc = of = None
        if coords is not None:
            x = coords.to(td)
            mean = torch.as_tensor(self._mean(x)).to(device=x.device, dtype=torch.float64)
            oc = torch.cat([x, (mean + eps).to(td)], dim=1)
        if forces is not None:
            f = forces.to(td)
            back = torch.as_tensor(self._back(scaled)).to(device=f.device, dtype=torch.float64)
            of = torch.cat([(f.to(torch.float64) + kbt * back).to(td), (-kbt * scaled).to(td)], dim=1)
        return oc, of

    # -- Augmenter interface
    def sample(self, source):
        host = not (isinstance(source, torch.Tensor) and source.is_cuda)
        dev = _engine.device()
        x = torch.as_tensor(source).to(device=dev)
        y = self.augment_device(x, None, 0.0, self.new_draw())[0][:, x.shape[1]:].contiguous()
        return _engine.to_host(y) if host else y

    def log_gradient(self, source, generated):
        host = not (isinstance(source, torch.Tensor) and source.is_cuda)
        dev = _engine.device()
        mean = torch.as_tensor(self._mean(source)).to(device=dev, dtype=self._tdtype())
        resid = torch.as_tensor(generated).to(device=dev, dtype=self._tdtype()) - mean
        scaled = resid / self.cov if self._cov_matrix is None else self._precision_times(resid)
        wrt_generated = -scaled
        wrt_source = torch.as_tensor(self._back(scaled)).to(device=dev, dtype=self._tdtype())
        if host:
            return _engine.to_host(wrt_source), _engine.to_host(wrt_generated)
        return wrt_source, wrt_generated

    # -- fused device path (csrc/augment.cu)
    def new_draw(self) -> "NoiseDraw":
        """Token identifying ONE noise realisation: every ``augment_device`` call made with it --
        whatever the frame slab -- sees the same noise (Philox keyed by seed / frame / bead / draw
        id, or the injected array)."""
        noise = None
        if self._noise is not None:
            noise = torch.as_tensor(self._noise).to(device=_engine.device(), dtype=self._tdtype()).contiguous()
            self._noise = None
        self._n_draws = getattr(self, "_n_draws", 0) + 1
        return NoiseDraw(self._n_draws, noise)

    def n_new_sites(self, n_sites: int) -> int:
        return n_sites if self.premap is None else int(self.premap.n_cg_sites)

    def _map_csr(self, n_sites: int):
        """Device CSR of the pre-map rows and of its transpose."""
        cached = getattr(self, "_csr", None)
        if cached is not None and cached[0] == n_sites:
            return cached[1]
        m = np.eye(n_sites) if self.premap is None else np.asarray(self.premap.standard_matrix, dtype=np.float64)
        if m.shape[1] != n_sites:
            raise ValueError(f"premap expects {m.shape[1]} sites but the array has {n_sites}")

        def csr(mat):
            rows, cols = np.nonzero(mat)
            ptr_ = np.zeros(mat.shape[0] + 1, dtype=np.int32)
            np.cumsum(np.bincount(rows, minlength=mat.shape[0]), out=ptr_[1:])
            if rows.size == 0:  # keep the device arrays non-empty
                return _engine.dev_i32(ptr_), _engine.dev_i32([0]), _engine.dev_f64([0.0])
            return _engine.dev_i32(ptr_), _engine.dev_i32(cols), _engine.dev_f64(mat[rows, cols])

        corr = np.ascontiguousarray(m.T)  # (n_sites, n_new): back-projection of the noise force
        if self.source_postmap is not None:
            post = np.asarray(self.source_postmap.standard_matrix, dtype=np.float64)
            if post.shape != (n_sites, n_sites):
                raise ValueError(f"source_postmap must map {n_sites} sites to {n_sites} sites; got {post.shape}")
            corr = post @ corr
        arrays = (*csr(m), *csr(corr), m.shape[0])
        self._csr = (n_sites, arrays)
        return arrays

    def augment_device(self, coords: Optional[torch.Tensor], forces: Optional[torch.Tensor], kbt: float,
                       draw: "NoiseDraw", frame0: int = 0):
        """Augmented device arrays for a slab of frames starting at global frame ``frame0``; either
        input may be ``None`` (its output is then ``None``)."""
        from .. import _lib

        if self._cov_matrix is not None:
            return self._augment_full_cov(coords, forces, kbt, draw, frame0)
        td = self._tdtype()
        ref = coords if coords is not None else forces
        n_frames, n_sites = int(ref.shape[0]), int(ref.shape[1])
        bp, bs, bw, sp, sb, sw, n_new = self._map_csr(n_sites)
        prep = lambda t: None if t is None else t.to(td).contiguous()  # noqa: E731
        coords, forces = prep(coords), prep(forces)
        mk = lambda t: None if t is None else torch.empty((n_frames, n_sites + n_new, 3), dtype=td, device=t.device)  # noqa: E731
        oc, of = mk(coords), mk(forces)
        noise = None
        if draw.noise is not None:
            noise = draw.noise[frame0 : frame0 + n_frames]
            if tuple(noise.shape) != (n_frames, n_new, 3):
                raise ValueError(f"injected noise has shape {tuple(draw.noise.shape)}; frames {frame0}.."
                                 f"{frame0 + n_frames} of (n_frames, {n_new}, 3) are needed")
        p = _engine.ptr
        _lib.call("agf_gauss_augment", p(coords), p(forces), _lib.F32 if td == torch.float32 else _lib.F64,
                  n_frames, n_sites, p(bp), p(bs), p(bw), n_new, p(sp), p(sb), p(sw), self.cov, float(kbt), p(noise),
                  self.seed, draw.index, int(frame0), p(oc), p(of), _engine.stream_ptr())
        return oc, of

    def augment(self, coords, forces, kbt: float):
        """Augmented ``(coords, forces)`` in one device pass (trajectory/core.py:382-389 of the reference)."""
        host = not (isinstance(coords, torch.Tensor) and coords.is_cuda)
        dev = _engine.device()
        x = torch.as_tensor(coords).to(device=dev)
        f = torch.as_tensor(forces).to(device=dev)
        full_coords, full_forces = self.augment_device(x, f, kbt, self.new_draw())
        if host:
            return _engine.to_host(full_coords), _engine.to_host(full_forces)
        return full_coords, full_forces

    def astype(self, dtype, *args, **kwargs) -> "CondNormal":  # noqa: ARG002
        return self.__class__(cov=self.cov, premap=self.premap, source_postmap=self.source_postmap, seed=self.seed,
                              dtype=dtype)


class NoiseDraw:
    """One noise realisation of a ``CondNormal`` (see ``CondNormal.new_draw``)."""

    def __init__(self, index: int, noise: Optional[torch.Tensor]) -> None:
        self.index = int(index)
        self.noise = noise


_AUG_SLAB_BYTES = 512 << 20


class AugmentedFrames(_engine.Frames):
    """Virtual ``(n_frames, n_fg + n_new, 3)`` frame source: the augmented coordinates OR forces of
    ``CondNormal``, generated slab by slab by ``agf_gauss_augment`` while a kernel consumes them.
    The augmented arrays of config 5 (2 000 atoms + 200 noise sites, 2 M frames: 53 GB each next
    to 48 GB each of input) are never materialised."""

    def __init__(self, source, augmenter: Optional[CondNormal] = None, kbt: float = 0.0,
                 draw: Optional[NoiseDraw] = None, which: str = "coords") -> None:
        if "_src" in self.__dict__:  # _engine.Frames(aug) passes an existing instance through
            return
        assert augmenter is not None and draw is not None and which in ("coords", "forces")
        self._src = _engine.Frames(source)
        self._augmenter, self._kbt, self._draw, self._which = augmenter, float(kbt), draw, which
        self._dev, self._host = None, None
        self.n_frames = self._src.n_frames
        self.n_sites = self._src.n_sites + augmenter.n_new_sites(self._src.n_sites)

    @property
    def on_host(self) -> bool:
        return self._src.on_host

    @property
    def np_dtype(self):
        return self._augmenter.dtype.type

    def pieces(self, start: int = 0, stop: Optional[int] = None):
        frame_bytes = self.n_sites * 3 * self._augmenter.dtype.itemsize
        per = max(4, (_AUG_SLAB_BYTES // frame_bytes) // 4 * 4)
        for t0, piece in self._src.pieces(start, stop):
            for a in range(0, piece.shape[0], per):
                sub = piece[a : a + per]
                c, f = (sub, None) if self._which == "coords" else (None, sub)
                oc, of = self._augmenter.augment_device(c, f, self._kbt, self._draw, frame0=t0 + a)
                yield t0 + a, (oc if self._which == "coords" else of)

    def resident(self) -> torch.Tensor:
        return torch.cat([p for _, p in self.pieces()])

    def gather(self, frame_indices) -> torch.Tensor:
        raise NotImplementedError("augmented frames are generated in slabs; gather from a materialised trajectory")

    def prefix(self, n: int) -> torch.Tensor:
        return torch.cat([p for _, p in self.pieces(0, n)])


# drop-in name of the reference's JAX class
JCondNormal = CondNormal
