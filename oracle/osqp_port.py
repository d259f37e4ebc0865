"""ctypes front-end of oracle/osqp_port.c.  TEST INFRASTRUCTURE ONLY.

Mirrors the call the reference makes at PredictiveControllers.py:269-283:
``OSQP().setup(P, q, A, l, u, polish=True)`` followed by ``solve()``, cold.
"""
import ctypes as C
import os
import subprocess
import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Settings(C.Structure):
    _fields_ = [("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
                ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("delta", C.c_double),
                ("max_iter", C.c_int), ("check_every", C.c_int), ("scaling_iters", C.c_int),
                ("adaptive_rho", C.c_int), ("adaptive_interval", C.c_int),
                ("adaptive_tol", C.c_double), ("polish", C.c_int), ("polish_refine", C.c_int),
                ("polish_strict", C.c_int)]


class Info(C.Structure):
    _fields_ = [("iters", C.c_int), ("status", C.c_int), ("polished", C.c_int), ("rho_updates", C.c_int),
                ("pri_res", C.c_double), ("dua_res", C.c_double), ("obj", C.c_double), ("rho_final", C.c_double)]


def _cpu_stamp():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def ensure_built():
    """(Re)build libosqp_port.so with -march=native for THIS host when needed."""
    so = os.path.join(_HERE, "libosqp_port.so")
    stamp = os.path.join(_HERE, "_ref", "build_cpu.txt")
    want = _cpu_stamp()
    have = open(stamp).read().strip() if os.path.exists(stamp) else None
    src_newer = os.path.exists(so) and os.path.getmtime(os.path.join(_HERE, "osqp_port.c")) > os.path.getmtime(so)
    if not os.path.exists(so) or have != want or src_newer:
        subprocess.run(["make", "-C", _HERE, "-B", "libosqp_port.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        os.makedirs(os.path.dirname(stamp), exist_ok=True)
        with open(stamp, "w") as f:
            f.write(want)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(ensure_built())
        L.osqp_port_default_settings.argtypes = [C.POINTER(Settings)]
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        L.osqp_port_solve.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, ip, ip, dp, dp, dp,
                                      C.POINTER(Settings), dp, dp, C.POINTER(Info)]
        L.osqp_port_solve.restype = C.c_int
        L.osqp_port_solve_batch.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, dp, dp, ip, ip, dp, dp, dp,
                                            C.POINTER(Settings), C.c_int, dp, dp, C.POINTER(Info)]
        L.osqp_port_solve_batch.restype = C.c_int
        L.osqp_port_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def default_settings(**kw):
    s = Settings()
    lib().osqp_port_default_settings(C.byref(s))
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def csc_pattern(P_dense_mask, A_dense_mask):
    """Fixed sparsity pattern (upper-tri P, full A) from boolean masks."""
    Pu = sp.csc_matrix(np.triu(P_dense_mask).astype(float))
    Ac = sp.csc_matrix(A_dense_mask.astype(float))
    return (Pu.indptr.astype(np.int32), Pu.indices.astype(np.int32)), (Ac.indptr.astype(np.int32), Ac.indices.astype(np.int32))


def gather_values(M, indptr, indices):
    """Values of dense M at a CSC pattern."""
    cols = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    return np.ascontiguousarray(M[indices, cols], dtype=np.float64)


def solve(P, q, A, l, u, settings=None, **kw):
    """Dense or scipy-sparse P (full symmetric), A.  Returns (x, info dict, y)."""
    s = settings if settings is not None else default_settings(**kw)
    Pu = sp.triu(sp.csc_matrix(P), format="csc")
    Pu.sort_indices()
    Ac = sp.csc_matrix(A)
    Ac.sort_indices()
    n, m = Pu.shape[0], Ac.shape[0]
    Pp, Pi, Px = Pu.indptr.astype(np.int32), Pu.indices.astype(np.int32), Pu.data.astype(np.float64)
    Ap, Ai, Ax = Ac.indptr.astype(np.int32), Ac.indices.astype(np.int32), Ac.data.astype(np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    l = np.ascontiguousarray(l, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    x = np.zeros(n)
    y = np.zeros(m)
    info = Info()
    lib().osqp_port_solve(n, m, _ip(Pp), _ip(Pi), _dp(Px), _dp(q), _ip(Ap), _ip(Ai), _dp(Ax), _dp(l), _dp(u),
                          C.byref(s), _dp(x), _dp(y), C.byref(info))
    return x, _info_dict(info), y


def _info_dict(i):
    return dict(iters=i.iters, status=i.status, polished=i.polished, rho_updates=i.rho_updates,
                pri_res=i.pri_res, dua_res=i.dua_res, obj=i.obj, rho_final=i.rho_final)


def solve_batch(pat_P, pat_A, Px, q, Ax, l, u, settings=None, nthreads=0, want_y=False, **kw):
    """Batch of QPs with one shared pattern.  Px[B,nnzP], q[B,n], Ax[B,nnzA], l/u[B,m]."""
    s = settings if settings is not None else default_settings(**kw)
    (Pp, Pi), (Ap, Ai) = pat_P, pat_A
    B, n = q.shape
    m = l.shape[1]
    Px, q, Ax, l, u = [np.ascontiguousarray(a, dtype=np.float64) for a in (Px, q, Ax, l, u)]
    x = np.zeros((B, n))
    y = np.zeros((B, m)) if want_y else None
    infos = (Info * B)()
    lib().osqp_port_solve_batch(B, n, m, _ip(Pp), _ip(Pi), _dp(Px), _dp(q), _ip(Ap), _ip(Ai), _dp(Ax), _dp(l), _dp(u),
                                C.byref(s), nthreads, _dp(x), _dp(y) if want_y else None, infos)
    return x, [_info_dict(i) for i in infos], y


def max_threads():
    return lib().osqp_port_max_threads()


def reference_qp(P, q, A, l, u):
    """QP back-end with the reference's settings (PC.py:275: defaults + polish)."""
    x, info, _ = solve(P, q, A, l, u)
    return x, info


def tight_qp(P, q, A, l, u):
    """Same algorithm driven to tight tolerances — the parity oracle for xPred/uPred."""
    x, info, _ = solve(P, q, A, l, u, eps_abs=1e-9, eps_rel=1e-9, max_iter=400000, polish_strict=1)
    return x, info
