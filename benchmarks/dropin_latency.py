#!/usr/bin/env python
"""Single-controller latency of the drop-in modules (racinglmpc_b200/compat) on the reference's own call sequence
(main.py:99-120, one LMPC lap, N = 12), next to the oracle's restatement of the reference controller (NumPy assembly +
OSQP-algorithm C port with the reference's settings) on one host core.  B = 1 is NOT what the GPU is for (one warp of one SM
works); this measures what a user who only swaps the two modules gets.

    python benchmarks/dropin_latency.py            -> one JSON line"""
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "racinglmpc_b200", "compat"))
import PredictiveControllers as PC          # noqa: E402  (the shim, under the reference's module name)
import PredictiveModel as PM                # noqa: E402
from racinglmpc_b200 import reference_params as rp   # noqa: E402
from oracle import vehicle, ftocp, ltv_model, osqp_port  # noqa: E402  (harness + CPU arm; benchmark only)
from oracle.track import TrackTable         # noqa: E402


class Timed:
    """Wraps a controller: records the wall time of every solve() (the call Simulator.sim makes, SysModel.py:34)."""

    def __init__(self, inner):
        object.__setattr__(self, "_inner", inner)
        object.__setattr__(self, "times", [])

    def solve(self, x):
        t0 = time.perf_counter()
        self._inner.solve(x)
        self.times.append(time.perf_counter() - t0)

    def __getattr__(self, k):
        return getattr(self._inner, k)

    def __setattr__(self, k, v):
        setattr(self._inner, k, v)


def main():
    track = TrackTable()
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    xP, uP, gP = g["pid_x"], g["pid_u"], g["pid_glob"]
    N = 12
    numSS_it, numSS_Points, _, _, Qts, q = rp.lmpc_params(N)
    par = PC.MPCParams(n=6, d=2, N=N, Q=q.Q, R=q.R, dR=q.dR, Fx=q.Fx, bx=np.array([[track.halfWidth], [track.halfWidth]]),
                       Fu=q.Fu, bu=np.array([[0.5], [0.5], [10.0], [10.0]]), slacks=True, Qslack=q.Qslack)
    par.timeVarying = True
    _, _, _, _, oQts, opar = ftocp.lmpc_params(track, N)
    opar.timeVarying = True
    x0 = np.array([0.5, 0, 0, 0, 0, 0.0])

    def lap(pm, make_lm):
        for _ in range(4):                      # main.py:102-104: the model is filled before the LMPC object exists
            pm.addTrajectory(xP.copy(), uP.copy())
        lm = make_lm(pm)
        for _ in range(4):
            lm.addTrajectory(xP.copy(), uP.copy(), gP.copy())
        np.random.seed(7)
        t = Timed(lm)
        t0 = time.perf_counter()
        xl, ul, gl, _ = vehicle.closed_loop(track, [x0, x0], t, multi_lap=False, is_lmpc=True, max_steps=400)
        return xl.shape[0], np.array(t.times), time.perf_counter() - t0

    pm = PM.PredictiveModel(6, 2, track, 4)
    n_gpu, t_gpu, wall_gpu = lap(pm, lambda m: PC.LMPC(numSS_Points, numSS_it, Qts, par, m))
    opm = ltv_model.LocalLTVModel(6, 2, track, 4)
    n_cpu, t_cpu, wall_cpu = lap(opm, lambda m: ftocp.OracleLMPC(numSS_Points, numSS_it, oQts, opar, m, qp=osqp_port.reference_qp))
    ms = lambda a: {"mean": float(a.mean() * 1e3), "median": float(np.median(a) * 1e3), "p99": float(np.quantile(a, 0.99) * 1e3)}
    print(json.dumps({"benchmark": "drop-in single-controller latency, one LMPC lap (main.py:99-120)", "N": N,
                      "gpu_shim": {"lap_steps": int(n_gpu), "solve_ms": ms(t_gpu[3:]), "lap_wall_s": wall_gpu},
                      "cpu_oracle_1core": {"lap_steps": int(n_cpu), "solve_ms": ms(t_cpu[3:]), "lap_wall_s": wall_cpu,
                                           "what": "NumPy restatement of PredictiveControllers/PredictiveModel + OSQP-algorithm C port (eps 1e-3, polish)"}}))


if __name__ == "__main__":
    main()
