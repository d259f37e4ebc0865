"""Whole-batch oracle runs for the parity tests (TEST INFRASTRUCTURE: imports oracle/).

The GPU path solves thousands of QPs per launch; comparing a handful of them with the oracle leaves the rest
unchecked.  These helpers push EVERY instance of a workload through the oracle:
  * the reference's matrix assembly (oracle/ftocp.py = PC.py:166-257,340-416) and, for LMPC steps, its k-NN
    regression and safe-set selection (oracle/ltv_model.py = PM.py:48-197), in a process pool (pure Python/NumPy);
  * the OSQP-algorithm C port driven to 1e-9 (oracle/osqp_port.c, OpenMP over the batch) for the optimum.
"""
import multiprocessing as mp
import os

import numpy as np

from oracle import ftocp, ltv_model, osqp_port


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def _stack_problems(Ps, qs, As, ls, us):
    maskP = np.any(np.array([p != 0 for p in Ps]), axis=0)
    maskA = np.any(np.array([a != 0 for a in As]), axis=0)
    patP, patA = osqp_port.csc_pattern(maskP, maskA)
    Px = np.stack([osqp_port.gather_values(p, *patP) for p in Ps])
    Ax = np.stack([osqp_port.gather_values(a, *patA) for a in As])
    return patP, patA, Px, np.stack(qs), Ax, np.stack(ls), np.stack(us)


def tight_batch(prob, nthreads=0):
    """Optimum of every QP of a stacked problem set: the OSQP algorithm driven to 1e-9 with the strict polish
    (the same settings as osqp_port.tight_qp).  Returns z[B,n], list of info dicts."""
    patP, patA, Px, q, Ax, l, u = prob
    z, infos, _ = osqp_port.solve_batch(patP, patA, Px, q, Ax, l, u, nthreads=nthreads or host_threads(),
                                        eps_abs=1e-9, eps_rel=1e-9, max_iter=400000, polish_strict=1)
    return z, infos


# ---------------------------------------------------------------------------------------------- configs[1]
def ltv_problem_set(x0, uold, abc, N, vt=0.8):
    """The B LTV-MPC QPs (126 vars / 174 rows at N = 12) exactly as the reference assembles them."""
    par = ftocp.mpc_params(6, 2, N, vt)[1]
    par.timeVarying = True
    F, bb = ftocp.build_ineq(par)
    Ps, qs, As, ls, us = [], [], [], [], []
    for b in range(x0.shape[0]):
        A = abc[b][:, 0:36].reshape(N, 6, 6)
        Bm = abc[b][:, 36:48].reshape(N, 6, 2)
        C = abc[b][:, 48:54]
        H, q = ftocp.build_cost(par, uold[b])
        G, E, L = ftocp.build_eq(par, list(A), list(Bm), list(C))
        P, q, Am, l, u = ftocp.osqp_form(H, q, F, bb, G, E @ x0[b] + L)
        Ps.append(P); qs.append(q); As.append(Am); ls.append(l); us.append(u)
    return _stack_problems(Ps, qs, As, ls, us)


# ---------------------------------------------------------------------------------------------- configs[2]
_LMPC_CTX = {}


def _lmpc_one(b):
    """Reference arithmetic of ONE controller step up to the QP hand-off (PC.py:110-121): regression along xLin/uLin,
    safe-set selection around zt, assembly.  Runs in a forked worker; reads the parent's workload through _LMPC_CTX."""
    data, track, N, trToUse = _LMPC_CTX["data"], _LMPC_CTX["track"], _LMPC_CTX["N"], _LMPC_CTX["trToUse"]
    numSS_it, numSS_Points, _, _, Qts, par = ftocp.lmpc_params(track, N)
    par.timeVarying = True
    pm = ltv_model.LocalLTVModel(6, 2, track, trToUse)
    pm.xStored = [lx for lx, _ in data["model_laps"][b]]
    pm.uStored = [lu for _, lu in data["model_laps"][b]]
    pm.lapTime = [lx.shape[0] for lx in pm.xStored]
    lm = ftocp.OracleLMPC(numSS_Points, numSS_it, Qts, par, pm, qp=None)
    lm.SS = [s[0] for s in data["ss_laps"][b]]
    lm.uSS = [s[1] for s in data["ss_laps"][b]]
    lm.Qfun = [s[2] for s in data["ss_laps"][b]]
    lm.LapTime = list(data["lap_times"])
    lm.it, lm.timeStep = 4, int(data["t"][b])
    lm.zt, lm.xLin, lm.uLin = data["zt"][b].copy(), data["xLin"][b].copy(), data["uLin"][b].copy()
    lm.OldInput, lm.xPred = data["OldInput"][b].copy(), data["xPred"][b].copy()
    P, q, A, l, u = lm.assemble(data["x0"][b])
    abc = np.zeros((N, 54))
    for k in range(N):
        abc[k, 0:36] = np.asarray(lm.A[k]).ravel(); abc[k, 36:48] = np.asarray(lm.B[k]).ravel(); abc[k, 48:54] = np.asarray(lm.C[k]).ravel()
    return dict(P=P, q=q, A=A, l=l, u=u, abc=abc, SS_sel=lm.SS_PointSelectedTot, Qfun_sel=lm.Qfun_SelectedTot,
                Succ_SS=lm.Succ_SS_PointSelectedTot, Succ_uSS=lm.Succ_uSS_PointSelectedTot)


def lmpc_oracle_steps(data, track, N=12, trToUse=5, nproc=0):
    """Every instance of a workloads.lmpc_batch() population through the oracle's pre-QP path.  Returns a dict of stacked
    arrays (abc[B,N,54], SS_sel[B,6,M], Qfun_sel[B,M], Succ_SS, Succ_uSS) and the stacked QPs for tight_batch()."""
    B = len(data["model_laps"])
    _LMPC_CTX.update(data=data, track=track, N=N, trToUse=trToUse)
    nproc = nproc or min(host_threads(), 64)
    if nproc > 1 and B >= 4 * nproc:
        with mp.get_context("fork").Pool(nproc) as pool:       # fork: the workload is shared copy-on-write, no pickling of laps
            res = pool.map(_lmpc_one, range(B), chunksize=max(1, B // (8 * nproc)))
    else:
        res = [_lmpc_one(b) for b in range(B)]
    _LMPC_CTX.clear()
    out = {k: np.stack([r[k] for r in res]) for k in ("abc", "SS_sel", "Qfun_sel", "Succ_SS", "Succ_uSS")}
    out["prob"] = _stack_problems([r["P"] for r in res], [r["q"] for r in res], [r["A"] for r in res],
                                  [r["l"] for r in res], [r["u"] for r in res])
    return out
