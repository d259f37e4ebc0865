"""Drop-in for src/fnc/controller/PredictiveModel.py (same module and class name, same constructor and
attributes).  Put this directory ahead of the reference's ``fnc/controller`` on ``sys.path`` (e.g.
``PYTHONPATH=.../racinglmpc_b200/compat python main.py``) and the reference's ``main.py`` imports it unchanged.

The object itself only stores laps (like the reference's lists); the k-NN regression
(PredictiveModel.py:48-197) runs on the GPU inside the controller that owns the model
(racinglmpc_b200/csrc/safeset.cuh, knn_ltv_regress_kernel).
"""
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
        # same ordering rule as PredictiveModel.py:35-46 (kept so that .xStored/.uStored read the same)
        if self.lapTime == [] or x.shape[0] >= self.lapTime[-1]:
            self.xStored.append(x)
            self.uStored.append(u)
            self.lapTime.append(x.shape[0])
        else:
            for i in range(0, len(self.xStored)):
                if x.shape[0] < self.lapTime[i]:
                    self.xStored.insert(i, x)
                    self.uStored.insert(i, u)
                    self.lapTime.insert(i, x.shape[0])
                    break
        self._added.append((x, u))

    def seg_table(self):
        """[s_start, length, curvature] rows consumed by Map.curvature (Track.py:292-310)."""
        return np.ascontiguousarray(np.asarray(self.map.PointAndTangent)[:, 3:6], dtype=float)

    def regressionAndLinearization(self, x, u):
        raise NotImplementedError("racinglmpc_b200: the regression runs inside the GPU controller step "
                                  "(MPC.solve / LMPC.solve); it is not available as a host call")
