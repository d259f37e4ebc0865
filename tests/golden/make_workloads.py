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
Horizons: N = 12 (bench) and N = 6, 24, 48 (horizon-sweep parity cases, fewer starts).
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
    for N, nstart in ((12, 256), (6, 32), (24, 32), (48, 32)):
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


if __name__ == "__main__":
    main()
