"""Oracle restatement of the closed-loop harness used to GENERATE inputs.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  These are the *callers* of
the hot path, restated so that the seeded PID lap
(`np.random.seed(0)`, main.py:45-66) and LMPC laps can be regenerated on a box
where /root/reference does not exist.

Follows:
  * ``Simulator.sim``       src/fnc/simulator/SysModel.py:22-54
  * ``Simulator.dynModel``  src/fnc/simulator/SysModel.py:56-147
  * ``PID.solve``           src/fnc/Utilities.py:42-68
  * ``Regression``          src/fnc/Utilities.py:5-28
The global NumPy RNG is consumed in the same order as the reference
(2 draws in PID.solve, then 3 in dynModel per step).
"""
import numpy as np

# vehicle constants, SysModel.py:61-70
_M, _LF, _LR, _IZ = 1.98, 0.125, 0.125, 0.024
_DF = 0.8 * _M * 9.81 / 2.0
_CF, _BF = 1.25, 1.0
_DR = 0.8 * _M * 9.81 / 2.0
_CR, _BR = 1.25, 1.0


def dyn_model(track, x, x_glob, u, dt=0.1, rng=None):
    """One control period: 100 explicit-Euler sub-steps of 1 ms + clipped noise.

    SysModel.py:56-147.  ``rng`` None => global np.random (reference behaviour).
    """
    randn = np.random.randn if rng is None else rng.standard_normal
    h = 0.001
    delta, a = u[0], u[1]
    psi, X, Y = x_glob[3], x_glob[4], x_glob[5]
    vx, vy, wz, epsi, s, ey = x[0], x[1], x[2], x[3], x[4], x[5]
    glob = np.zeros(6)
    cur_next = np.zeros(6)
    i = 0
    while (i + 1) * h <= dt:
        af = delta - np.arctan2(vy + _LF * wz, vx)
        ar = -np.arctan2(vy - _LF * wz, vx)
        Fyf = _DF * np.sin(_CF * np.arctan(_BF * af))
        Fyr = _DR * np.sin(_CR * np.arctan(_BR * ar))
        nvx = vx + h * (a - 1 / _M * Fyf * np.sin(delta) + wz * vy)
        nvy = vy + h * (1 / _M * (Fyf * np.cos(delta) + Fyr) - wz * vx)
        nwz = wz + h * (1 / _IZ * (_LF * Fyf * np.cos(delta) - _LR * Fyr))
        glob[0], glob[1], glob[2] = nvx, nvy, nwz
        glob[3] = psi + h * wz
        glob[4] = X + h * (vx * np.cos(psi) - vy * np.sin(psi))
        glob[5] = Y + h * (vx * np.sin(psi) + vy * np.cos(psi))
        cur = track.curvature(s)
        cur_next[0], cur_next[1], cur_next[2] = nvx, nvy, nwz
        cur_next[3] = epsi + h * (wz - (vx * np.cos(epsi) - vy * np.sin(epsi)) / (1 - cur * ey) * cur)
        cur_next[4] = s + h * ((vx * np.cos(epsi) - vy * np.sin(epsi)) / (1 - cur * ey))
        cur_next[5] = ey + h * (vx * np.sin(epsi) + vy * np.cos(epsi))
        psi, X, Y = glob[3], glob[4], glob[5]
        vx, vy, wz, epsi, s, ey = cur_next
        i += 1
    n_vx = np.max([-0.05, np.min([randn() * 0.01, 0.05])])
    n_vy = np.max([-0.05, np.min([randn() * 0.01, 0.05])])
    n_wz = np.max([-0.05, np.min([randn() * 0.005, 0.05])])
    cur_next[0] += 0.01 * n_vx
    cur_next[1] += 0.01 * n_vy
    cur_next[2] += 0.01 * n_wz
    return cur_next.copy(), glob.copy()


class PIDFollower:
    """Utilities.py:42-68 — noisy path follower at constant target speed."""

    def __init__(self, vt, rng=None):
        self.vt = vt
        self.uPred = np.zeros([1, 2])
        self._randn = np.random.randn if rng is None else rng.standard_normal

    def solve(self, x0):
        r = self._randn
        self.uPred[0, 0] = -0.6 * x0[5] - 0.9 * x0[3] + np.max([-0.9, np.min([r() * 0.25, 0.9])])
        self.uPred[0, 1] = 1.5 * (self.vt - x0[0]) + np.max([-0.2, np.min([r() * 0.10, 0.2])])


def closed_loop(track, x_start, controller, *, multi_lap=True, is_lmpc=False,
                max_time=100, dt=0.1, rng=None, max_steps=None):
    """SysModel.py:22-54.  Returns x_cl[T,6], u_cl[T,2], x_glob[T,6], xF."""
    xs = [np.asarray(x_start[0], dtype=float)]
    gs = [np.asarray(x_start[1], dtype=float)]
    us = []
    i = 0
    done = False
    limit = int(max_time / dt) if max_steps is None else max_steps
    while i < limit and not done:
        controller.solve(xs[-1])
        us.append(controller.uPred[0, :].copy())
        if is_lmpc:
            controller.addPoint(xs[-1], us[-1])
        xt, gt = dyn_model(track, xs[-1], gs[-1], us[-1], dt=dt, rng=rng)
        xs.append(xt)
        gs.append(gt)
        if (not multi_lap) and xs[-1][4] > track.TrackLength:
            done = True
        i += 1
    xF = [np.array(xs[-1]) - np.array([0, 0, 0, 0, track.TrackLength, 0]), np.array(gs[-1])]
    xs.pop()
    gs.pop()
    return np.array(xs), np.array(us), np.array(gs), xF


def ridge_sysid(x, u, lamb):
    """Utilities.py:5-28 — LTI least squares [A B] with ridge ``lamb``."""
    Y = x[2:x.shape[0], :]
    X = np.hstack((x[1:(x.shape[0] - 1), :], u[1:(x.shape[0] - 1), :]))
    Qi = np.linalg.inv(np.dot(X.T, X) + lamb * np.eye(X.shape[1]))
    W = np.dot(Qi, np.dot(X.T, Y))
    A = W.T[:, 0:6]
    B = W.T[:, 6:8]
    err = np.dot(X, W) - Y
    return A, B, np.vstack((np.max(err, axis=0), np.min(err, axis=0)))
