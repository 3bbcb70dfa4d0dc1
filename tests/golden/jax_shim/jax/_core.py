"""A minimal stand-in for the part of JAX that the reference uses -- FIXTURE GENERATION ONLY.

``tests/golden/make_golden.py`` puts this directory on ``sys.path`` so that the UNMODIFIED
reference modules ``aggforce/qp/jaxfeat.py``, ``aggforce/jaxutil.py``,
``aggforce/map/jaxlinearmap.py``, ``aggforce/trajectory/jaxgausstraj.py``, ``aggforce/qp/jgauss.py``
and ``aggforce/jaxmapval.py`` import and run in a container without jax/jaxlib.  Arrays are torch CPU
tensors behind a thin wrapper; ``jit`` is the identity; ``grad``/``jacrev``/``jacfwd``/``vmap`` are
``torch.func`` transforms.  Nothing in ``aggforce_b200``, the oracle or the tests imports this.

Semantics reproduced on purpose (because fixtures depend on them):
  * default dtype float32 / int32 (JAX with x64 disabled): float64 inputs are narrowed;
  * Python scalars are weakly typed (a float32 array times a Python float stays float32);
  * ``linalg.norm`` is ``sqrt(sum(x*x))`` so that its derivative at 0 is NaN as in JAX
    (torch's own ``linalg.norm`` returns the 0 subgradient);
  * ``x.at[idx].set(y)`` is functional, and **returns ``x`` unchanged when the indexed slice is
    empty, whatever the shape of ``y``** -- JAX's scatter has this early exit
    (``jax/_src/ops/scatter.py:_scatter_impl``: "Avoid calling scatter if the slice shape is empty",
    before ``y`` is broadcast).  This is what makes ``channel_allocate``
    (reference ``qp/jaxfeat.py:343-349,361-366``) silently drop the block of the largest label
    (SURVEY Q5): the slice ``[n_feats*max_channels, n_feats*(max_channels+1))`` of an array that is
    ``n_feats*max_channels`` wide clamps to an empty slice.  Set ``STRICT_EMPTY_SCATTER = True`` to
    get the other reading (raise like numpy would on the shape mismatch).
Not reproduced: XLA's float32 rounding (torch CPU kernels round differently at the 1e-7 level) and
the threefry random stream (``random.multivariate_normal`` draws from a seeded torch generator and
RECORDS the standard-normal draw so that fixtures can carry the noise).
"""
from __future__ import annotations

import builtins
from typing import Any

import numpy as np
import torch

STRICT_EMPTY_SCATTER = False

_F = torch.float32
_I = torch.int64  # torch indexes with int64; reported as int32 where it matters


def _narrow(t: torch.Tensor) -> torch.Tensor:
    if t.dtype == torch.float64:
        return t.to(_F)
    if t.dtype in (torch.int32, torch.int16, torch.int8, torch.uint8):
        return t.to(_I)
    return t


def unwrap(x: Any) -> Any:
    """JArray / numpy / python scalar -> torch tensor (python scalars stay python: weak typing)."""
    if isinstance(x, JArray):
        return x._t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        return _narrow(torch.from_numpy(np.ascontiguousarray(x)))
    if isinstance(x, np.generic):
        return x.item()
    if isinstance(x, (list, tuple)) and x and all(isinstance(v, JArray) for v in x):
        return torch.stack([v._t for v in x])
    if isinstance(x, (list, tuple)):
        return _narrow(torch.as_tensor(np.asarray(x)))
    return x


def wrap(t: Any) -> Any:
    if isinstance(t, torch.Tensor):
        return JArray(t)
    if isinstance(t, tuple):
        return tuple(wrap(v) for v in t)
    if isinstance(t, list):
        return [wrap(v) for v in t]
    return t


def unwrap_tree(x: Any) -> Any:
    if isinstance(x, tuple):
        return tuple(unwrap_tree(v) for v in x)
    if isinstance(x, list):
        return [unwrap_tree(v) for v in x]
    return unwrap(x)


def _index(idx: Any) -> Any:
    if isinstance(idx, tuple):
        return tuple(_index(i) for i in idx)
    if isinstance(idx, JArray):
        return idx._t
    if isinstance(idx, np.ndarray):
        return torch.from_numpy(idx)
    if isinstance(idx, list):
        return torch.as_tensor(idx)
    return idx


def _axis(kw: dict) -> dict:
    if "axis" in kw:
        ax = kw.pop("axis")
        kw["dim"] = tuple(ax) if isinstance(ax, (list, tuple)) else ax
    return kw


class _At:
    def __init__(self, arr: "JArray") -> None:
        self.arr = arr

    def __getitem__(self, idx: Any) -> "_AtIdx":
        return _AtIdx(self.arr, _index(idx))


class _AtIdx:
    def __init__(self, arr: "JArray", idx: Any) -> None:
        self.arr, self.idx = arr, idx

    def set(self, y: Any) -> "JArray":
        base = self.arr._t
        target = base[self.idx]
        if target.numel() == 0 and not STRICT_EMPTY_SCATTER:
            return self.arr  # JAX: empty slice shape -> operand returned as is (see module docstring)
        out = base.clone()
        out[self.idx] = unwrap(y)
        return JArray(out)

    def add(self, y: Any) -> "JArray":
        base = self.arr._t
        if base[self.idx].numel() == 0 and not STRICT_EMPTY_SCATTER:
            return self.arr
        out = base.clone()
        out[self.idx] = out[self.idx] + unwrap(y)
        return JArray(out)


class JArray:
    """The shim's ``jax.Array``."""

    __array_priority__ = 100.0

    def __init__(self, t: torch.Tensor) -> None:
        self._t = t

    # ---- introspection
    @property
    def shape(self) -> tuple:
        return tuple(self._t.shape)

    @property
    def ndim(self) -> int:
        return self._t.ndim

    @property
    def size(self) -> int:
        return self._t.numel()

    @property
    def dtype(self) -> np.dtype:
        return np.dtype(str(self._t.dtype).replace("torch.", "")) if self._t.dtype != _I else np.dtype("int32")

    def __len__(self) -> int:
        return self._t.shape[0]

    def __iter__(self):
        return (JArray(v) for v in self._t)

    def __array__(self, dtype=None, copy=None):
        out = self._t.detach().cpu().numpy()
        if self._t.dtype == _I:
            out = out.astype(np.int32)
        return out if dtype is None else out.astype(dtype)

    def __float__(self) -> float:
        return float(self._t)

    def __int__(self) -> int:
        return int(self._t)

    def __bool__(self) -> bool:
        return bool(self._t)

    def __repr__(self) -> str:
        return f"JArray({self._t!r})"

    def item(self):
        return self._t.item()

    # ---- indexing
    def __getitem__(self, idx: Any) -> "JArray":
        return JArray(self._t[_index(idx)])

    @property
    def at(self) -> _At:
        return _At(self)

    # ---- arithmetic
    def _bin(self, other: Any, fn) -> "JArray":
        return JArray(fn(self._t, unwrap(other)))

    def _rbin(self, other: Any, fn) -> "JArray":
        o = unwrap(other)
        if not isinstance(o, torch.Tensor):
            o = torch.as_tensor(o, dtype=self._t.dtype if self._t.is_floating_point() or isinstance(o, int) else _F)
        return JArray(fn(o, self._t))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._rbin(o, torch.add)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._rbin(o, torch.sub)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._rbin(o, torch.mul)
    def __truediv__(self, o): return self._bin(o, torch.true_divide)
    def __rtruediv__(self, o): return self._rbin(o, torch.true_divide)
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __rpow__(self, o): return self._rbin(o, torch.pow)
    def __matmul__(self, o): return self._bin(o, torch.matmul)
    def __rmatmul__(self, o): return self._rbin(o, torch.matmul)
    def __neg__(self): return JArray(-self._t)
    def __lt__(self, o): return self._bin(o, torch.lt)
    def __le__(self, o): return self._bin(o, torch.le)
    def __gt__(self, o): return self._bin(o, torch.gt)
    def __ge__(self, o): return self._bin(o, torch.ge)

    # ---- methods the reference calls
    def sum(self, *a, **kw): return JArray(self._t.sum(*[tuple(v) if isinstance(v, list) else v for v in a], **_axis(kw)))
    def mean(self, *a, **kw):
        kw = {k: v for k, v in _axis(kw).items() if v is not None and k in ("dim", "keepdim")}
        return JArray(self._t.mean(*a, **kw))
    def max(self, *a, **kw): return JArray(self._t.amax(*a, **_axis(kw))) if (a or kw) else JArray(self._t.max())
    def min(self, *a, **kw): return JArray(self._t.amin(*a, **_axis(kw))) if (a or kw) else JArray(self._t.min())

    def reshape(self, *shape, **kw) -> "JArray":
        if "newshape" in kw:
            shape = (kw["newshape"],)
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return JArray(self._t.reshape(*shape))

    def astype(self, dtype) -> "JArray":
        return JArray(self._t.to(_torch_dtype(dtype)))

    @property
    def T(self) -> "JArray":
        return JArray(self._t.T)

    def swapaxes(self, a: int, b: int) -> "JArray":
        return JArray(self._t.swapaxes(a, b))


def _torch_dtype(dtype: Any):
    if dtype is None:
        return None
    if isinstance(dtype, torch.dtype):
        return dtype
    name = np.dtype(dtype).name
    return {"float32": _F, "float64": _F, "int32": _I, "int64": _I, "bool": torch.bool}[name]


def asarray(x: Any, dtype: Any = None) -> JArray:
    t = unwrap(x)
    if not isinstance(t, torch.Tensor):
        t = _narrow(torch.as_tensor(t))
    td = _torch_dtype(dtype)
    if td is not None:
        t = t.to(td)
    return JArray(t)


# ------------------------------------------------------------------ transforms
def jit(fn=None, **_):
    if fn is None:
        return lambda f: f
    return fn


def _argnums(argnums):
    return (argnums,) if isinstance(argnums, int) else tuple(argnums)


def _transform(tf, fn, argnums, **tf_kw):
    single = isinstance(argnums, int)
    nums = _argnums(argnums)

    def wrapped(*args, **kwargs):
        args = list(args)

        def inner(*diff):
            full = list(args)
            for i, d in zip(nums, diff):
                full[i] = JArray(d)
            return unwrap_tree(fn(*full, **kwargs))

        prim = [t if isinstance(t, torch.Tensor) else torch.as_tensor(t, dtype=_F) for t in (unwrap(args[i]) for i in nums)]
        out = tf(inner, argnums=0 if single else tuple(range(len(nums))), **tf_kw)(*prim)
        return wrap(out)

    return wrapped


def grad(fn, argnums=0):
    return _transform(torch.func.grad, fn, argnums)


def jacrev(fn, argnums=0):
    return _transform(torch.func.jacrev, fn, argnums)


def jacfwd(fn, argnums=0):
    return _transform(torch.func.jacfwd, fn, argnums)


def vmap(fn, in_axes=0, out_axes=0):
    def wrapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        mapped = [i for i, a in enumerate(axes) if a is not None]

        def inner(*batched):
            full = list(args)
            for i, b in zip(mapped, batched):
                full[i] = JArray(b)
            return unwrap_tree(fn(*full))

        out = torch.func.vmap(inner, in_dims=tuple(axes[i] for i in mapped), out_dims=out_axes)(
            *[unwrap(args[i]) for i in mapped])
        return wrap(out)

    return wrapped


_ = builtins  # keep linters quiet about the unused import guard
