"""Oracle restatement of the data-driven LTV model (k-NN local regression).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows src/fnc/controller/PredictiveModel.py:
  * lap store sorted by length          PM.py:35-46
  * ``regressionAndLinearization``      PM.py:48-139
  * ``computeIndices`` (k-NN + kernel)  PM.py:180-197
  * ``compute_Q_M`` / ``compute_b``     PM.py:141-168
  * ``LMPC_LocLinReg``                  PM.py:170-178  (cvxopt.qp(Q,b) with no
    constraints == solve(Q, -b); cvxopt is an unpinned pip dependency, README.md:17)
"""
import numpy as np
from numpy import linalg as la

STATE_FEATS = [0, 1, 2]     # PM.py:28
IN_FEATS_VX = [1]           # PM.py:29
IN_FEATS_LAT = [0]          # PM.py:30


class LocalLTVModel:
    def __init__(self, n, d, track, laps_to_use):
        self.map = track
        self.n, self.d = n, d
        self.xStored, self.uStored, self.lapTime = [], [], []
        self.MaxNumPoint = 7                       # PM.py:18
        self.h = 5                                 # PM.py:19
        self.lamb = 0.0                            # PM.py:20
        self.dt = 0.1                              # PM.py:21
        self.scaling = np.diag([0.1, 1.0, 1.0, 1.0, 1.0])   # PM.py:22-26
        self.usedIt = list(range(laps_to_use))     # PM.py:31
        self.last_indices, self.last_K = None, None

    # PM.py:35-46 — keep laps ordered by number of rows, ascending
    def addTrajectory(self, x, u):
        if self.lapTime == [] or x.shape[0] >= self.lapTime[-1]:
            self.xStored.append(x)
            self.uStored.append(u)
            self.lapTime.append(x.shape[0])
            return
        for i in range(len(self.xStored)):
            if x.shape[0] < self.lapTime[i]:
                self.xStored.insert(i, x)
                self.uStored.insert(i, u)
                self.lapTime.insert(i, x.shape[0])
                break

    # PM.py:180-197
    def knn(self, q, lap):
        xs, us = self.xStored[lap], self.uStored[lap]
        data = np.hstack((xs[0:-1, STATE_FEATS], us[0:-1, :]))
        diff = np.dot(data - q[None, :], self.scaling)
        dist = la.norm(diff, 1, axis=1)
        near = np.squeeze(np.where(dist < self.h))
        if near.shape[0] >= self.MaxNumPoint:      # note: 0-d squeeze raises like the reference
            idx = np.argsort(dist)[0:self.MaxNumPoint]
        else:
            idx = near
        K = (1 - (dist[idx] / self.h) ** 2) * 3 / 4
        return idx, K

    def _normal_eq(self, in_feats, idxs, Ks):
        # PM.py:141-155
        rows, wts = [], []
        for c, lap in enumerate(self.usedIt):
            rows.append(np.hstack((self.xStored[lap][np.ix_(idxs[c], STATE_FEATS)],
                                   self.uStored[lap][np.ix_(idxs[c], in_feats)])))
            wts.append(Ks[c])
        X0 = np.vstack(rows)
        Kt = np.concatenate(wts)
        M = np.hstack((X0, np.ones((X0.shape[0], 1))))
        Q0 = np.dot(np.dot(M.T, np.diag(Kt)), M)
        return Q0 + self.lamb * np.eye(Q0.shape[0]), M, Kt

    def _rhs(self, y_index, idxs, M, Kt):
        # PM.py:157-168
        y = np.concatenate([np.atleast_1d(np.squeeze(self.xStored[lap][idxs[c] + 1, y_index]))
                            for c, lap in enumerate(self.usedIt)])
        return -np.dot(np.dot(M.T, np.diag(Kt)), y)

    @staticmethod
    def _fit(Q, b):
        # PM.py:170-178: qp(Q,b) unconstrained  <=>  Q theta = -b
        theta = np.linalg.solve(Q, -b)
        return theta[0:3], theta[3:4], theta[-1]

    # PM.py:48-139
    def regressionAndLinearization(self, x, u):
        A = np.zeros((self.n, self.n))
        B = np.zeros((self.n, self.d))
        C = np.zeros(self.n)
        q = np.hstack((x[STATE_FEATS], u[:]))
        idxs, Ks = [], []
        for lap in self.usedIt:
            i_l, k_l = self.knn(q, lap)
            idxs.append(i_l)
            Ks.append(k_l)
        self.last_indices, self.last_K = idxs, Ks

        Qv, Mv, Kv = self._normal_eq(IN_FEATS_VX, idxs, Ks)
        A[0, STATE_FEATS], B[0, IN_FEATS_VX], C[0] = self._fit(Qv, self._rhs(0, idxs, Mv, Kv))
        Ql, Ml, Kl = self._normal_eq(IN_FEATS_LAT, idxs, Ks)
        A[1, STATE_FEATS], B[1, IN_FEATS_LAT], C[1] = self._fit(Ql, self._rhs(1, idxs, Ml, Kl))
        A[2, STATE_FEATS], B[2, IN_FEATS_LAT], C[2] = self._fit(Ql, self._rhs(2, idxs, Ml, Kl))

        vx, vy, wz, epsi, s, ey = x[0], x[1], x[2], x[3], x[4], x[5]
        dt = self.dt
        cur = self.map.curvature(s)                # PM.py:95-96
        den = 1 - cur * ey
        # epsi row, PM.py:102-110
        A[3, :] = [-dt * np.cos(epsi) / den * cur,
                   dt * np.sin(epsi) / den * cur,
                   dt,
                   1 - dt * (-vx * np.sin(epsi) - vy * np.cos(epsi)) / den * cur,
                   0,
                   dt * (vx * np.cos(epsi) - vy * np.sin(epsi)) / (den ** 2) * cur * (-cur)]
        C[3] = epsi + dt * (wz - (vx * np.cos(epsi) - vy * np.sin(epsi)) / (1 - cur * ey) * cur) - np.dot(A[3, :], x)
        # s row, PM.py:114-122
        A[4, :] = [dt * (np.cos(epsi) / den),
                   -dt * (np.sin(epsi) / den),
                   0,
                   dt * (-vx * np.sin(epsi) - vy * np.cos(epsi)) / den,
                   1,
                   -dt * (vx * np.cos(epsi) - vy * np.sin(epsi)) / (den ** 2) * (-cur)]
        C[4] = s + dt * ((vx * np.cos(epsi) - vy * np.sin(epsi)) / (1 - cur * ey)) - np.dot(A[4, :], x)
        # ey row, PM.py:127-135
        A[5, :] = [dt * np.sin(epsi), dt * np.cos(epsi), 0,
                   dt * (vx * np.cos(epsi) - vy * np.sin(epsi)), 0, 1]
        C[5] = ey + dt * (vx * np.sin(epsi) + vy * np.cos(epsi)) - np.dot(A[5, :], x)
        return A, B, C
