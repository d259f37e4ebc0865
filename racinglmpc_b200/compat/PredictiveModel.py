"""Drop-in for src/fnc/controller/PredictiveModel.py (same module and class name, same constructor and
attributes).  Put this directory ahead of the reference's ``fnc/controller`` on ``sys.path`` (e.g.
``PYTHONPATH=.../racinglmpc_b200/compat python main.py``) and the reference's ``main.py`` imports it unchanged.

The object itself only stores laps (like the reference's lists); the k-NN regression
(PredictiveModel.py:48-197) runs on the GPU inside the controller that owns the model
(racinglmpc_b200/csrc/safeset.cuh, knn_ltv_regress_kernel).
"""
import bisect

import numpy as np


class PredictiveModel():
    def __init__(self, n, d, map, trToUse):
        self.map = map
        self.n = n
        self.d = d
        self.xStored = []
        self.uStored = []
        self.MaxNumPoint = 7          # PredictiveModel.py:18
        self.h = 5                    # :19
        self.lamb = 0.0               # :20
        self.dt = 0.1                 # :21
        self.scaling = np.array([[0.1, 0.0, 0.0, 0.0, 0.0],
                                 [0.0, 1.0, 0.0, 0.0, 0.0],
                                 [0.0, 0.0, 1.0, 0.0, 0.0],
                                 [0.0, 0.0, 0.0, 1.0, 0.0],
                                 [0.0, 0.0, 0.0, 0.0, 1.0]])
        self.stateFeatures = [0, 1, 2]
        self.inputFeaturesVx = [1]
        self.inputFeaturesLat = [0]
        self.usedIt = [i for i in range(trToUse)]
        self.lapTime = []
        self._added = []              # (x, u) in call order; controllers replay it into their device store

    def addTrajectory(self, x, u):
        """PredictiveModel.addTrajectory (PredictiveModel.py:35-46): laps stay ordered by length, a new lap goes behind the
        stored laps that are not longer than it -- so usedIt = [0..trToUse-1] always names the fastest laps."""
        T = x.shape[0]
        pos = bisect.bisect_right(self.lapTime, T)
        for store, item in ((self.xStored, x), (self.uStored, u), (self.lapTime, T)):
            store.insert(pos, item)
        self._added.append((x, u))

    def seg_table(self):
        """[s_start, length, curvature] rows consumed by Map.curvature (Track.py:292-310)."""
        return np.ascontiguousarray(np.asarray(self.map.PointAndTangent)[:, 3:6], dtype=float)

    def regressionAndLinearization(self, x, u):
        """PredictiveModel.regressionAndLinearization (PredictiveModel.py:48-139) for ONE point as a host call: runs the device
        regression kernel (csrc/safeset.cuh knn_ltv_regress_kernel) on a private one-instance engine holding the stored
        laps and returns (A 6x6, B 6x2, C 6).  Controllers do not use this -- their regression runs inside solve()."""
        from racinglmpc_b200.controller import BatchedController
        from racinglmpc_b200 import reference_params as rp
        eng = getattr(self, "_probe_engine", None)
        if eng is None:
            eng = self._probe_engine = BatchedController(
                rp.mpc_params(6), 1, self.seg_table(), self.map.TrackLength, trToUse=len(self.usedIt),
                model_kwargs=dict(MaxNumPoint=self.MaxNumPoint, h=float(self.h), lamb=float(self.lamb), dt=float(self.dt),
                                  scaling=tuple(np.diag(self.scaling))))
            self._probe_seen = 0
        while self._probe_seen < len(self._added):
            eng.model_add_trajectory(0, *self._added[self._probe_seen])
            self._probe_seen += 1
        N = eng.N
        eng.set_state(xLin=np.tile(np.asarray(x, float), (N + 1, 1)), uLin=np.tile(np.asarray(u, float), (N, 1)))
        abc, flags = eng.identify()
        if flags[0] & 1:
            raise ArithmeticError("singular local regression (the reference's cvxopt call raises here too)")
        rec = abc[0, 0]
        return rec[0:36].reshape(6, 6).copy(), rec[36:48].reshape(6, 2).copy(), rec[48:54].copy()
