"""Host-side equality-constrained QP:  min 1/2 x'Px  s.t.  Ax = b.

The reference hands this problem to ``qpsolvers.solve_qp(..., solver="osqp")`` once per bead
(``qplinear.py:83-85``, ``featlinearmap.py:375-381``).  It has no inequality constraints and no
linear term, so the minimiser is available in closed form; solving it exactly (one Cholesky
factorisation shared by all beads / right-hand sides) is what the parity tolerance "weights
within 1e-6" is defined against.  ``qpsolvers`` is used instead when the caller
asks for it with ``solver_args={"backend": "qpsolvers", ...}`` and the package is importable.
"""
from __future__ import annotations

from typing import Any, Mapping, Optional

import numpy as np
import scipy.linalg as sl
import scipy.sparse as ss

DEFAULT_SOLVER_OPTIONS = {
    "solver": "osqp",
    "eps_abs": 1e-7,
    "max_iter": int(1e3),
    "polish": True,
    "polish_refine_iter": 10,
}
SolverOptions = Mapping[str, Any]


def _dense(m) -> np.ndarray:
    return m.toarray() if ss.issparse(m) else np.asarray(m, dtype=np.float64)


def solve_equality_qp(P, A, b) -> Optional[np.ndarray]:
    """Exact minimiser; ``b`` may be a vector or a matrix with one column per problem.

    ``x = P^-1 A' (A P^-1 A')^+ b`` when ``P`` is positive definite, otherwise a null-space
    solve (``x = x0 + Z w`` with ``Z`` spanning ``ker A``).  Returns ``None`` if the problem is
    infeasible or non-finite, mirroring ``solve_qp``'s failure value.
    """
    P, A, b = _dense(P), _dense(A), np.asarray(b, dtype=np.float64)
    if not (np.isfinite(P).all() and np.isfinite(A).all() and np.isfinite(b).all()):
        return None
    try:
        cf = sl.cho_factor(P, lower=True, check_finite=False)
        pia = sl.cho_solve(cf, A.T, check_finite=False)
        s = A @ pia
        try:  # A P^-1 A' is small (n_cg x n_cg) and SPD when A has full row rank
            lam = sl.cho_solve(sl.cho_factor(s, lower=True, check_finite=False), b, check_finite=False)
            if not np.isfinite(lam).all():
                raise sl.LinAlgError("singular Schur complement")
        except (sl.LinAlgError, np.linalg.LinAlgError, ValueError):
            lam = np.linalg.lstsq(s, b, rcond=None)[0]
        x = pia @ lam
    except (sl.LinAlgError, np.linalg.LinAlgError):
        _, sv, vt = np.linalg.svd(A, full_matrices=True)
        rank = int((sv > sv.max() * 1e-12).sum()) if sv.size else 0
        x0 = np.linalg.lstsq(A, b, rcond=None)[0]
        z = vt[rank:].T
        if z.shape[1] == 0:
            x = x0
        else:
            w = np.linalg.lstsq(z.T @ P @ z, -(z.T @ (P @ x0)), rcond=None)[0]
            x = x0 + z @ w
    resid = A @ x - b
    if not np.isfinite(x).all() or np.abs(resid).max(initial=0.0) > 1e-6 * max(1.0, np.abs(b).max(initial=0.0)):
        return None
    return x


def solve_equality_qp_device(P, A, b, keep_on_device: bool = False):
    """Same closed form on the GPU for large problems: ``P`` is a float64 CUDA tensor
    (the Gram never leaves the device), one cuSOLVER Cholesky + multi-right-hand-side solves through
    ``torch.linalg``.  Returns ``None`` when ``P`` or the Schur complement is not numerically
    positive definite (or the solution misses the equality constraints) -- the caller then falls back
    to :func:`solve_equality_qp` on the host.  A singular Schur complement (redundant equality rows, as
    the featurised fit produces) is handled like the host path: least squares on the small system.
    ``keep_on_device``: return the solution as a CUDA tensor (checked like the host copy) instead of
    downloading it -- the fitted map of a large problem is applied from where it is."""
    import torch

    a = A if isinstance(A, torch.Tensor) else torch.as_tensor(np.asarray(A, dtype=np.float64), device=P.device)
    rhs = torch.as_tensor(np.asarray(b, dtype=np.float64), device=P.device)
    vec = rhs.ndim == 1
    if vec:
        rhs = rhs[:, None]
    chol, info = torch.linalg.cholesky_ex(P)
    if int(info.item()) != 0:
        return None
    pia = torch.cholesky_solve(a.T.contiguous(), chol)  # P^-1 A'
    schur = a @ pia
    chol_s, info_s = torch.linalg.cholesky_ex(schur)
    if int(info_s.item()) == 0:
        lam = torch.cholesky_solve(rhs, chol_s)
    else:  # redundant equality rows: the small (n_rows x n_rows) system goes to the host's least squares
        lam_h = np.linalg.lstsq(schur.cpu().numpy(), rhs.cpu().numpy(), rcond=None)[0]
        lam = torch.as_tensor(lam_h, device=P.device)
    x = pia @ lam
    resid = float((a @ x - rhs).abs().max().item()) if x.numel() else 0.0
    if not bool(torch.isfinite(x).all().item()) or resid > 1e-6 * max(1.0, float(rhs.abs().max().item())):
        return None
    if keep_on_device:
        return x[:, 0] if vec else x
    out = x.cpu().numpy()
    return out[:, 0] if vec else out


def solve_equality_qp_device_batched(P, A, b) -> Optional[np.ndarray]:
    """:func:`solve_equality_qp_device` for a batch of problems of one shape -- the beads of a featurised fit:
    ``P (B, n, n)``, ``A (B, m, n)``, ``b (B, m)`` float64 CUDA tensors.  Batched Cholesky of ``P``, batched
    Schur complement; a singular Schur complement (redundant equality rows are the rule there) is solved in the
    minimum-norm least-squares sense through its symmetric eigen-decomposition.  One read at the
    end checks every member (positive definite, finite, equality rows met); ``None`` if any fails -- the caller
    then solves bead by bead with the fallbacks of the single-problem path.  Returns ``(B, n)`` on the host."""
    import torch

    chol, info = torch.linalg.cholesky_ex(P)
    pia = torch.cholesky_solve(A.transpose(1, 2).contiguous(), chol)  # P^-1 A'
    schur = A @ pia
    schur = 0.5 * (schur + schur.transpose(1, 2))
    # Redundant equality rows are the rule in a featurised fit, so the (m x m) Schur complements are singular:
    # minimum-norm least squares through the symmetric eigen-decomposition with numpy lstsq's cutoff (eps * m times
    # the largest singular value) -- what the single-problem path computes by SVD on the host.  (Measured at B = 10,
    # m = 200: 16 ms here, 22 ms for ten lstsq calls on the host, worse with one host thread per member.)
    w, v = torch.linalg.eigh(schur)
    m = schur.shape[-1]
    cutoff = torch.finfo(torch.float64).eps * m * w.abs().amax(dim=-1, keepdim=True)
    inv = torch.where(w.abs() > cutoff, 1.0 / w, torch.zeros_like(w))
    rhs = b.unsqueeze(-1)
    lam = v @ (inv.unsqueeze(-1) * (v.transpose(1, 2) @ rhs))
    x = (pia @ lam).squeeze(-1)
    resid = (A @ x.unsqueeze(-1) - rhs).abs().amax()
    bad = torch.stack([info.ne(0).any().to(torch.float64), (~torch.isfinite(x)).any().to(torch.float64), resid,
                       rhs.abs().amax()])
    bad_h = bad.cpu().numpy()
    if bad_h[0] != 0 or bad_h[1] != 0 or not bad_h[2] <= 1e-6 * max(1.0, bad_h[3]):
        return None
    return x.cpu().numpy()


_WARNED = [False]


def warn_if_solver_ignored(solver_args: Optional[SolverOptions]) -> None:
    """The reference forwards ``solver_args`` to ``qpsolvers.solve_qp`` (``solver="osqp"``, tolerances,
    ...).  Here the equality QP is solved exactly unless ``backend="qpsolvers"`` is given: say so once
    when a caller passes its own solver options, instead of silently ignoring them."""
    if solver_args is None or solver_args is DEFAULT_SOLVER_OPTIONS or _WARNED[0]:
        return
    opts = dict(solver_args)
    if opts.get("backend", "exact") == "exact" and any(k != "backend" for k in opts):
        import warnings

        _WARNED[0] = True
        warnings.warn("aggforce_b200 solves the equality-constrained QP exactly (closed form); the given "
                      f"solver options {sorted(k for k in opts if k != 'backend')} are not used.  Pass "
                      'solver_args={"backend": "qpsolvers", ...} to hand the problem to qpsolvers as the '
                      "reference does.", stacklevel=3)


def solve(P, A, b, solver_args: Optional[SolverOptions] = None) -> Optional[np.ndarray]:
    """Dispatch: exact solve by default, ``qpsolvers`` on request (vector ``b`` only)."""
    opts = dict(solver_args or {})
    if opts.pop("backend", "exact") == "qpsolvers":
        from qpsolvers import solve_qp  # type: ignore[import-not-found]

        return solve_qp(P=P, q=np.zeros(P.shape[0]), A=A, b=b, **opts)
    return solve_equality_qp(P, A, b)
