#!/usr/bin/env python
"""Generate tests/golden/workload_ltv.npz — inputs of BASELINE.json configs[1]
("batch=4096 LTV-MPC QPs N=12, fixed A/B/C matrices"), following SURVEY §8d:

  np.random.seed(0) PID lap (main.py:45-66)  ->  PredictiveModel(n,d,map,1) (main.py:89-90)
  -> (A_k,B_k,C_k)_{k<N} by regressionAndLinearization along the lap at NSTART start indices,
     linearisation points xLin = xPID[i:i+N], uLin = uPID[i:i+N]   (as MPC.__init__, PC.py:88-91)
  -> x0 = xPID[i], OldInput = uPID[i-1].

Produced with the ORACLE restatement (oracle/ltv_model.py), which tests/golden/make_golden.py pins
bit-exactly against the real reference.  bench.py and the tests tile these NSTART model sets to the
batch size and perturb x0 per instance with np.random.default_rng(1) noise.
Horizons: N = 12 (bench), N = 6, 24, 48 (horizon-sweep parity cases, fewer starts) and N = 14 (the reference's own, main.py:43).
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ltv_model            # noqa: E402
from oracle.track import TrackTable     # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    g = np.load(os.path.join(OUT, "reference_golden.npz"))
    xP, uP = g["pid_x"], g["pid_u"]
    trk = TrackTable()
    pm = ltv_model.LocalLTVModel(6, 2, trk, 1)
    pm.addTrajectory(xP, uP)
    out = {}
    for N, nstart in ((12, 256), (6, 32), (24, 32), (48, 32), (14, 32)):
        T = xP.shape[0]
        starts = 1 + (np.arange(nstart) * (T - N - 2)) // nstart
        abc = np.zeros((nstart, N, 54))
        for j, i in enumerate(starts):
            for k in range(N):
                A, B, C = pm.regressionAndLinearization(xP[i + k], uP[i + k])
                abc[j, k, 0:36], abc[j, k, 36:48], abc[j, k, 48:54] = A.ravel(), B.ravel(), C
        out["N%d_abc" % N] = abc
        out["N%d_x0" % N] = xP[starts]
        out["N%d_uold" % N] = uP[starts - 1]
        out["N%d_start" % N] = starts
        print("N", N, "starts", nstart)
    np.savez_compressed(os.path.join(OUT, "workload_ltv.npz"), **out)


if __name__ == "__main__" and "--lmpc" not in sys.argv:
    main()


def make_lmpc_workload():
    """tests/golden/workload_lmpc.npz — inputs of BASELINE.json configs[2] ("batch=4096 full LMPC steps with k-NN LTV
    regression over 5-lap safe set"), SURVEY §8d: the seed-0 PID lap + LMPC laps 4..7 from the oracle closed loop
    (reference arithmetic, OSQP-algorithm solver at 1e-9), and the full controller state before every solve of lap 8."""
    from oracle import ftocp, vehicle, osqp_port
    g = np.load(os.path.join(OUT, "reference_golden.npz"))
    xP, uP, gP = g["pid_x"].copy(), g["pid_u"].copy(), g["pid_glob"].copy()
    trk = TrackTable()
    N = 12
    numSS_it, numSS_Points, _, _, Qts, par = ftocp.lmpc_params(trk, N)
    par.timeVarying = True
    pm = ltv_model.LocalLTVModel(6, 2, trk, 4)
    for _ in range(4):
        pm.addTrajectory(xP, uP)
    lm = ftocp.OracleLMPC(numSS_Points, numSS_it, Qts, par, pm, qp=osqp_port.tight_qp)
    for _ in range(4):
        lm.addTrajectory(xP, uP, gP)
    np.random.seed(11)
    x0 = np.array([0.5, 0, 0, 0, 0, 0.0])
    xs_lap = [x0, x0]
    out = {"pid_x": g["pid_x"], "pid_u": g["pid_u"]}
    for lap in range(4, 9):
        rec = []
        xs, gs, us = [xs_lap[0]], [xs_lap[1]], []
        t = 0
        while True:
            if lap == 8:
                rec.append(dict(x0=np.array(xs[-1]), xLin=np.array(lm.xLin), uLin=np.array(lm.uLin), zt=np.array(lm.zt),
                                OldInput=np.array(lm.OldInput).ravel(), xPred=np.array(lm.xPred)))
            lm.solve(xs[-1])
            us.append(lm.uPred[0].copy())
            lm.addPoint(xs[-1], us[-1])
            xt, gt = vehicle.dyn_model(trk, xs[-1], gs[-1], us[-1])
            xs.append(xt); gs.append(gt)
            t += 1
            if xs[-1][4] > trk.TrackLength or t >= 400:
                break
        xs_lap = [np.array(xs[-1]) - np.array([0, 0, 0, 0, trk.TrackLength, 0]), np.array(gs[-1])]
        xs.pop(); gs.pop()
        xl, ul, gl = np.array(xs), np.array(us), np.array(gs)
        out["lap%d_x" % lap], out["lap%d_u" % lap] = xl, ul
        print("lap", lap, xl.shape[0])
        if lap == 8:
            for k in rec[0]:
                out["lap8_state_" + k] = np.array([r[k] for r in rec])
        lm.addTrajectory(xl, ul, gl)
        pm.addTrajectory(xl, ul)
    np.savez_compressed(os.path.join(OUT, "workload_lmpc.npz"), **out)


if __name__ == "__main__" and "--lmpc" in sys.argv:
    make_lmpc_workload()
