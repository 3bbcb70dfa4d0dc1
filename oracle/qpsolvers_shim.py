"""Stand-in for the third-party ``qpsolvers`` module the reference imports (TEST / BENCH ONLY).

The reference calls ``solve_qp(P, q=0, A=..., b=..., solver="osqp", ...)`` for
``min 1/2 x'Px  s.t.  Ax = b`` (src/aggforce/qp/qplinear.py:83-85, qp/featlinearmap.py:375-381).
``qpsolvers``/``osqp`` are not installed in this image and are unpinned in the reference's
``setup.cfg``; the problem has the closed form ``x = P^-1 A' (A P^-1 A')^+ b``.
"""
import numpy as np
import scipy.sparse as ss


def solve_qp(P, q, A=None, b=None, **_):  # noqa: N803
    p = P.toarray() if ss.issparse(P) else np.asarray(P, float)
    a = A.toarray() if ss.issparse(A) else np.asarray(A, float)
    pi_at = np.linalg.solve(p, a.T)
    return pi_at @ np.linalg.lstsq(a @ pi_at, np.asarray(b, float), rcond=None)[0]
