"""Gaussian-binned distance features (``gb_feat``).

The reference implements these with JAX (``src/aggforce/qp/jaxfeat.py``).  Here

* inside ``qp_feat_linear_map`` the configuration ``Multifeaturize([id_feat, Curry(gb_feat, ...)])``
  is recognised and never materialised: kernel (b) evaluates the features in shared memory
  (``csrc/featgram.cu``);
* calling ``gb_feat`` directly keeps the reference's contract -- a dict of per-bead feature /
  divergence arrays (generators when ``lazy``) -- computed on the GPU with the closed-form
  divergence (float32 results like the reference's; this is a convenience path, not the hot one).

Per bead ``c``, frame ``t``, site ``a`` with constraint-group label ``l(a)``:
``feat[t, a, l(a)*n_basis + k] = g_k(|p~_{l(a)} - R_c|)`` where ``p~`` is the group mean position,
``R_c`` the mapped bead position and ``g_k(d) = max(exp(-((d-mu_k)/width)^2), clip) - clip``;
``div[t, l*n_basis + k, :] = m_l g_k'(d_l) (p~_l - R_c)/d_l`` (bead held fixed).  As in the reference
(``jaxfeat.py:115``) the feature array is ``n_basis * max(labels)`` wide, so the block of the LAST
label is dropped; ``drop_last_channel=False`` keeps it.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Final, Iterable, Union

import numpy as np
import torch

from .. import _engine
from ..constraints import Constraints
from ..map import LinearMap

DIVMETHOD_REORDER: Final = "reorder"
DIVMETHOD_BASIC: Final = "basic"


@dataclass(frozen=True)
class GbSpec:
    """Parameters of one ``gb_feat`` configuration."""

    outer: float
    inner: float = 0.0
    n_basis: int = 10
    width: float = 1.0
    dist_power: float = 0.5
    clip: float = 1e-3
    drop_last_channel: bool = True

    def centers(self) -> np.ndarray:
        """Gaussian centres: evenly spaced in ``d**dist_power`` (jaxfeat.py:235-236)."""
        grid = np.linspace(float(self.inner) ** self.dist_power, float(self.outer) ** self.dist_power, self.n_basis)
        return grid ** (1.0 / self.dist_power)

    def n_channels(self, n_labels: int) -> int:
        return n_labels - 1 if self.drop_last_channel else n_labels


def _materialise(points: torch.Tensor, cmap_row: torch.Tensor, labels: torch.Tensor, n_labels: int, spec: GbSpec,
                 div_method: str = DIVMETHOD_BASIC):
    """(feats (T, n, n_ch*nb), divs (T, n_ch*nb, 3)) in float64 on the device for one bead."""
    T, n, _ = points.shape
    nb, n_ch = spec.n_basis, spec.n_channels(n_labels)
    x = points.to(torch.float64)
    sums = torch.zeros((T, n_labels, 3), dtype=torch.float64, device=x.device).index_add_(1, labels, x)
    size = torch.zeros(n_labels, dtype=torch.float64, device=x.device).index_add_(
        0, labels, torch.ones(n, dtype=torch.float64, device=x.device))
    mean = sums / size[None, :, None]
    bead = torch.einsum("f,tfd->td", cmap_row.to(torch.float64), x)
    disp = mean - bead[:, None, :]
    dist = torch.linalg.vector_norm(disp, dim=-1)
    mu = torch.as_tensor(spec.centers(), device=x.device)
    z = (dist[..., None] - mu) / spec.width
    e = torch.exp(-(z * z))
    g = torch.clamp(e, min=spec.clip) - spec.clip
    gp = torch.where(e > spec.clip, -2.0 * z / spec.width * e, torch.zeros_like(e))
    feats = torch.zeros((T, n, n_ch * nb), dtype=torch.float64, device=x.device)
    site = torch.arange(n, device=x.device)
    keep = labels < n_ch
    for k in range(nb):
        feats[:, site[keep], labels[keep] * nb + k] = g[:, labels[keep], k]
    unit = disp / dist[..., None]
    divs = (size[None, :n_ch, None, None] * gp[:, :n_ch, :, None] * unit[:, :n_ch, None, :]).reshape(T, n_ch * nb, 3)
    if div_method == DIVMETHOD_REORDER:
        # reverse-mode autodiff of the reference (jaxfeat.py:544-565): the NaN cotangent of a group that
        # coincides with the bead (d = 0) meets the zeros of the smear matrix, 0 * NaN = NaN, and the
        # whole frame's divergence is NaN (pinned by tests/golden/ref_gbfeat.npz); forward mode ("basic")
        # confines it to the coincident group's own channel
        divs[(dist == 0.0).any(dim=1)] = float("nan")
    return feats, divs


def gb_feat(
    points,
    cmap: LinearMap,
    constraints: Constraints,
    outer: float,
    inner: float = 0,
    n_basis: int = 10,
    width: float = 1.0,
    dist_power: float = 0.5,
    batch_size: Union[None, int] = None,
    lazy: bool = True,
    div_method: str = DIVMETHOD_REORDER,
    drop_last_channel: bool = True,
) -> Dict[str, Union[Iterable, None]]:
    """Featurise every site by Gaussian bins of its (constraint-smeared) distance to each bead.

    Signature and return value as the reference's (``jaxfeat.py:20-184``): ``{"feats": per-bead
    (n_frames, n_fg, n_feat) arrays, "divs": per-bead (n_frames, n_feat, 3) arrays, "names": None}``,
    generators when ``lazy``.  ``batch_size`` is accepted for compatibility (frames are processed
    on the device in one pass); both ``div_method`` values give the same closed form and differ only
    in how far the NaN of a group coinciding with the bead spreads (see ``_materialise``).

    ``drop_last_channel=True`` is the reference's behaviour (Q5): ``channel_allocate`` scatters the
    block of the largest label into an empty slice, which JAX's scatter ignores; it is pinned by the
    reference run behind the jax stand-in (``tests/golden/ref_gbfeat.npz``), not by real jaxlib.
    """
    if div_method not in (DIVMETHOD_REORDER, DIVMETHOD_BASIC):
        raise ValueError("Unknown method for jacobian calculation.")
    from .featlinearmap import id_feat

    spec = GbSpec(outer=outer, inner=inner, n_basis=n_basis, width=width, dist_power=dist_power,
                  drop_last_channel=drop_last_channel)
    labels_np = id_feat(points, cmap, constraints, return_ids=True)
    dev = _engine.device()
    host = not (isinstance(points, torch.Tensor) and points.is_cuda)
    pts = torch.as_tensor(np.asarray(points) if host else points).to(dev)
    labels = torch.as_tensor(labels_np.astype(np.int64), device=dev)
    n_labels = int(labels_np.max()) + 1
    cm = torch.as_tensor(np.asarray(cmap.standard_matrix, dtype=np.float64), device=dev)

    def one(bead: int, which: int):
        out = _materialise(pts, cm[bead], labels, n_labels, spec, div_method)[which].to(torch.float32)
        return _engine.to_host(out) if host else out

    feats = (one(c, 0) for c in range(cmap.n_cg_sites))
    divs = (one(c, 1) for c in range(cmap.n_cg_sites))
    if not lazy:
        feats, divs = list(feats), list(divs)
    return {"feats": feats, "divs": divs, "names": None}
