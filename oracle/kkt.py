"""Solver-independent fp64 KKT checker for  min 1/2 z'Pz + q'z  s.t.  l <= Az <= u.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is how parity is pinned
in the absence of the real OSQP binary: any (z, y) returned by any
solver is judged by
  r_prim = || Az - clip(Az, l, u) ||_inf
  r_dual = || Pz + q + A'y ||_inf
  r_comp = max_i  [ y_i^+ (u_i - (Az)_i) , y_i^- ((Az)_i - l_i) ]  (sign + complementarity)
When no multipliers are available, ``dual_from_primal`` recovers them by a
least-squares fit on the active set, which certifies optimality of z alone.
"""
import numpy as np


def residuals(P, q, A, l, u, z, y):
    Az = A @ z
    r_prim = float(np.max(np.abs(Az - np.clip(Az, l, u)))) if len(Az) else 0.0
    r_dual = float(np.max(np.abs(P @ z + q + A.T @ y)))
    yp, ym = np.maximum(y, 0), np.maximum(-y, 0)
    gu = np.where(np.isfinite(u), u - Az, 0.0)
    gl = np.where(np.isfinite(l), Az - l, 0.0)
    bad_sign = np.max(np.where(~np.isfinite(u), yp, 0.0), initial=0.0) + np.max(np.where(~np.isfinite(l), ym, 0.0), initial=0.0)
    r_comp = float(max(np.max(yp * np.abs(gu), initial=0.0), np.max(ym * np.abs(gl), initial=0.0), bad_sign))
    return dict(r_prim=r_prim, r_dual=r_dual, r_comp=r_comp)


def objective(P, q, z):
    return float(0.5 * z @ (P @ z) + q @ z)


def dual_from_primal(P, q, A, l, u, z, tol=1e-6):
    """Multipliers supported on rows within ``tol`` of a bound, sign-constrained NNLS."""
    from scipy.optimize import nnls
    Az = A @ z
    eq = np.isfinite(l) & np.isfinite(u) & (np.abs(u - l) < 1e-12)
    up = (~eq) & np.isfinite(u) & (u - Az < tol)
    lo = (~eq) & np.isfinite(l) & (Az - l < tol)
    g = P @ z + q
    cols, rows, sgn = [], [], []
    for i in np.where(eq)[0]:
        cols += [A[i], -A[i]]
        rows += [i, i]
        sgn += [1.0, -1.0]
    for i in np.where(up)[0]:
        cols.append(A[i]); rows.append(i); sgn.append(1.0)
    for i in np.where(lo)[0]:
        cols.append(-A[i]); rows.append(i); sgn.append(-1.0)
    y = np.zeros(A.shape[0])
    if not cols:
        return y
    M = np.array(cols).T
    c, _ = nnls(M, -g, maxiter=50 * M.shape[1])
    for k, i in enumerate(rows):
        y[i] += sgn[k] * c[k]
    return y
