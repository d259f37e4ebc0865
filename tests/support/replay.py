"""Rebuild oracle controllers from the full-state snapshots stored in tests/golden/reference_golden.npz
(written by tests/golden/make_golden.py while driving the REAL reference).  TEST INFRASTRUCTURE ONLY."""
import numpy as np
from oracle import ftocp, ltv_model, osqp_port

LMPC_KEYS = [(4, 0), (4, 1), (4, 60), (4, 200), (5, 0), (5, 90), (6, 0), (6, 40)]


def lmpc_from_snapshot(g, key, track, qp=osqp_port.tight_qp, N=12):
    k = "lmpc_%d_%d_" % key
    numSS_it, numSS_Points, _, _, Qts, par = ftocp.lmpc_params(track, N)
    par.timeVarying = True
    pm = ltv_model.LocalLTVModel(6, 2, track, 4)
    npm = int(g[k + "pm_nlap"])
    pm.xStored = [g[k + "pmx%d" % j].copy() for j in range(npm)]
    pm.uStored = [g[k + "pmu%d" % j].copy() for j in range(npm)]
    pm.lapTime = [a.shape[0] for a in pm.xStored]
    c = ftocp.OracleLMPC(numSS_Points, numSS_it, Qts, par, pm, qp=qp)
    nl = int(g[k + "nlap"])
    c.SS = [g[k + "SS%d" % j].copy() for j in range(nl)]
    c.uSS = [g[k + "uSS%d" % j].copy() for j in range(nl)]
    c.Qfun = [g[k + "Qfun%d" % j].copy() for j in range(nl)]
    c.LapTime = list(g[k + "LapTime"])
    c.it, c.timeStep = int(g[k + "it"]), int(g[k + "timeStep"])
    c.zt = g[k + "zt"].copy()
    c.xLin, c.uLin = g[k + "xLin"].copy(), g[k + "uLin"].copy()
    c.OldInput = g[k + "OldInput"].copy()
    c.xPred = g[k + "xPred_prev"].copy() if int(g[k + "has_pred"]) else []
    if not int(g[k + "has_pred"]):
        # very first LMPC solve: main.py:103-110 stores ONE PID array object in every safe-set slot and in
        # the regression model, and xLin is a view of it (PC.py:432) -> re-create the aliasing so that the
        # PC.py:394 write lands where it does in the reference.
        shared = c.SS[0]
        for j in range(nl):
            if np.array_equal(c.SS[j], shared):
                c.SS[j] = shared
        for j in range(npm):
            if np.array_equal(pm.xStored[j], shared):
                pm.xStored[j] = shared
        c.xLin = shared[1:N + 2, :]
    return c, g[k + "x0"].copy()
