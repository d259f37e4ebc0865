"""Oracle restatement of the MPC / LMPC finite-time optimal control problem.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows src/fnc/controller/PredictiveControllers.py (``PC.py``):
  * parameter bag                    PC.py:24-51
  * ``MPC.__init__`` / ``solve``     PC.py:63-137
  * ``buildIneqConstr``              PC.py:166-198
  * ``buildEqConstr``                PC.py:200-226
  * ``buildCost``                    PC.py:228-257
  * OSQP call semantics              PC.py:259-283  (rows = [F; G], l = [-inf; beq], u = [b; beq])
  * LMPC additions                   PC.py:293-416
  * safe-set bookkeeping             PC.py:418-514
Quirks that are reproduced on purpose (SURVEY §8a notes): Qslack = [quad, lin]
(PC.py:249-250); ``xLin[4,-1]`` write at the lap wrap (PC.py:394) which, on the
very first LMPC solve, lands in the lap array that ``xLin`` is a view of
(PC.py:432); ``self.xPred == []`` (PC.py:502) read as "no prediction yet";
float ``numPoints`` (PC.py:403,492-495).

Decision vector: z = [x_0..x_N | u_0..u_{N-1} | s (stage major) | lambda | xi].
"""
from dataclasses import dataclass, field
import numpy as np
from numpy import linalg as la


@dataclass
class FTOCPParams:
    """Same field names as the reference ``MPCParams`` (PC.py:24-51)."""
    n: int = None
    d: int = None
    N: int = None
    A: object = None
    B: object = None
    Q: object = None
    R: object = None
    Qf: object = None
    dR: object = None
    Qslack: object = None
    Fx: object = None
    bx: object = None
    Fu: object = None
    bu: object = None
    xRef: object = None
    slacks: bool = True
    timeVarying: bool = False

    def __post_init__(self):
        if self.Qf is None:
            self.Qf = np.zeros((self.n, self.n))
        if self.dR is None:
            self.dR = np.zeros(self.d)
        if self.xRef is None:
            self.xRef = np.zeros(self.n)


def mpc_params(n, d, N, vt):
    """initControllerParameters.py:4-26 (values only)."""
    Fx = np.array([[0., 0., 0., 0., 0., 1.], [0., 0., 0., 0., 0., -1.]])
    bx = np.array([[2.], [2.]])
    Fu = np.kron(np.eye(2), np.array([1, -1])).T
    bu = np.array([[0.5], [0.5], [10.0], [10.0]])
    Q = np.diag([1.0, 1.0, 1, 1, 0.0, 100.0])
    R = np.diag([1.0, 10.0])
    xRef = np.array([vt, 0, 0, 0, 0, 0])
    Qslack = 1 * np.array([0, 50])
    mk = lambda: FTOCPParams(n=n, d=d, N=N, Q=Q, R=R, Fx=Fx, bx=bx, Fu=Fu, bu=bu,
                             xRef=xRef, slacks=True, Qslack=Qslack)
    return mk(), mk()


def lmpc_params(track, N):
    """initControllerParameters.py:28-58 (values only)."""
    Fx = np.array([[0., 0., 0., 0., 0., 1.], [0., 0., 0., 0., 0., -1.]])
    bx = np.array([[track.halfWidth], [track.halfWidth]])
    Fu = np.kron(np.eye(2), np.array([1, -1])).T
    bu = np.array([[0.5], [0.5], [10.0], [10.0]])
    numSS_it = 4
    numSS_Points = 12 * numSS_it
    Laps = 40 + numSS_it
    TimeLMPC = 400
    QterminalSlack = 500 * np.diag([1, 1, 1, 1, 1, 1])
    Qslack = 1 * np.array([5, 25])
    Q = 0 * np.diag([0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
    R = 0 * np.diag([1.0, 1.0])
    dR = 5 * np.array([1.0, 10.0])
    p = FTOCPParams(n=6, d=2, N=N, Q=Q, R=R, dR=dR, Fx=Fx, bx=bx, Fu=Fu, bu=bu,
                    slacks=True, Qslack=Qslack)
    return numSS_it, numSS_Points, Laps, TimeLMPC, QterminalSlack, p


# ----------------------------------------------------------------------------
# matrix assembly (dense, like the reference; sizes are tiny)
# ----------------------------------------------------------------------------
def build_ineq(p):
    """PC.py:166-198.  F z <= b over z = [x | u | s]."""
    n, d, N = p.n, p.d, p.N
    ncx, ncu = p.Fx.shape[0], p.Fu.shape[0]
    nz = n * (N + 1) + d * N + (ncx * N if p.slacks else 0)
    rows = ncx * N + ncu * N + (ncx * N if p.slacks else 0)
    F = np.zeros((rows, nz))
    for k in range(N):                           # x_N is left unconstrained (PC.py:171)
        F[k * ncx:(k + 1) * ncx, k * n:(k + 1) * n] = p.Fx
        r0 = ncx * N + k * ncu
        c0 = n * (N + 1) + k * d
        F[r0:r0 + ncu, c0:c0 + d] = p.Fu
    b = np.hstack((np.tile(np.squeeze(p.bx), N), np.tile(np.squeeze(p.bu), N)))
    if p.slacks:
        s0 = n * (N + 1) + d * N
        F[0:ncx * N, s0:s0 + ncx * N] = -np.eye(ncx * N)                 # Fx x - s <= bx
        F[ncx * N + ncu * N:, s0:s0 + ncx * N] = -np.eye(ncx * N)        # -s <= 0
        b = np.hstack((b, np.zeros(ncx * N)))
    return F, b


def build_eq(p, A, B, C):
    """PC.py:200-226.  G z = E x(t) + L."""
    n, d, N = p.n, p.d, p.N
    Gx = np.eye(n * (N + 1))
    Gu = np.zeros((n * (N + 1), d * N))
    E = np.zeros((n * (N + 1), n))
    E[0:n] = np.eye(n)
    L = np.zeros(n * (N + 1))
    for i in range(N):
        r = slice(n + i * n, n + i * n + n)
        if p.timeVarying:
            Gx[r, i * n:i * n + n] = -A[i]
            Gu[r, i * d:i * d + d] = -B[i]
            L[r] = C[i]
        else:
            Gx[r, i * n:i * n + n] = -A
            Gu[r, i * d:i * d + d] = -B
    G = np.hstack((Gx, Gu, np.zeros((Gx.shape[0], p.Fx.shape[0] * N)))) if p.slacks else np.hstack((Gx, Gu))
    return G, E, L


def build_cost(p, old_input):
    """PC.py:228-257.  0.5 z'Hz + q'z (H already carries the factor 2)."""
    n, d, N = p.n, p.d, p.N
    dR = np.asarray(p.dR, dtype=float)
    Hx = np.kron(np.eye(N), p.Q)
    Hu = np.kron(np.eye(N), p.R + 2 * np.diag(dR))
    for i in range(d):                           # last input appears once in the rate cost
        Hu[i - d, i - d] -= dR[i]
    off = -np.tile(dR, N - 1)
    np.fill_diagonal(Hu[d:], off)
    np.fill_diagonal(Hu[:, d:], off)
    Hxu = np.zeros((n * (N + 1) + d * N,) * 2)
    Hxu[0:n * N, 0:n * N] = Hx
    Hxu[n * N:n * (N + 1), n * N:n * (N + 1)] = p.Qf
    Hxu[n * (N + 1):, n * (N + 1):] = Hu
    q = -2 * np.dot(np.append(np.tile(p.xRef, N + 1), np.zeros(p.R.shape[0] * N)), Hxu)
    q[n * (N + 1):n * (N + 1) + d] = -2 * np.dot(old_input, np.diag(dR))
    if p.slacks:
        ns = p.Fx.shape[0] * N
        H = np.zeros((Hxu.shape[0] + ns,) * 2)
        H[0:Hxu.shape[0], 0:Hxu.shape[0]] = Hxu
        H[Hxu.shape[0]:, Hxu.shape[0]:] = p.Qslack[0] * np.eye(ns)
        q = np.append(q, p.Qslack[1] * np.ones(ns))
    else:
        H = Hxu
    return 2 * H, q


def add_safe_set(F, b, G, E, L, H, q, n, N, SS_sel, Qfun_sel, Qts):
    """PC.py:340-362: append (lambda, xi) columns and the terminal rows."""
    m = SS_sel.shape[1]
    nz = G.shape[1]
    F2 = np.zeros((F.shape[0] + m, nz + m + n))
    F2[0:F.shape[0], 0:nz] = F
    F2[F.shape[0]:, nz:nz + m] = -np.eye(m)
    b2 = np.append(b, np.zeros(m))
    G2 = np.zeros((G.shape[0] + n + 1, nz + m + n))
    G2[0:G.shape[0], 0:nz] = G
    r = G.shape[0]
    G2[r:r + n, N * n:(N + 1) * n] = np.eye(n)
    G2[r:r + n, nz:nz + m] = -SS_sel
    G2[r:r + n, nz + m:] = np.eye(n)
    G2[r + n, nz:nz + m] = 1.0
    E2 = np.vstack((E, np.zeros((n + 1, n))))
    L2 = np.append(np.append(L, np.zeros(n)), 1)
    H2 = np.zeros((nz + m + n,) * 2)
    H2[0:nz, 0:nz] = H
    H2[nz + m:, nz + m:] = 2 * Qts
    q2 = np.append(np.append(q, Qfun_sel), np.zeros(n))
    return F2, b2, G2, E2, L2, H2, q2


def osqp_form(H, q, F, b, G, beq):
    """PC.py:270-273: A = [F; G], l = [-inf; beq], u = [b; beq]."""
    A = np.vstack((F, G))
    l = np.hstack((-np.inf * np.ones(len(b)), beq))
    u = np.hstack((b, beq))
    return H, q, A, l, u


def rollout_cost(x, track_length):
    """PC.py:447-464 — backward count of steps to the finish line."""
    T = x.shape[0]
    cost = 10000 * np.ones(T)
    for i in range(T):
        j = T - 1 - i
        if i == 0:
            cost[j] = 0
        elif x[j, 4] < track_length:
            cost[j] = cost[j + 1] + 1
        else:
            cost[j] = 0
    return cost


# ----------------------------------------------------------------------------
# controller state machines
# ----------------------------------------------------------------------------
class OracleMPC:
    """PC.py:56-283 with a pluggable QP back-end ``qp(H,q,A,l,u) -> (z, info)``."""

    def __init__(self, params, model=None, qp=None):
        p = self.p = params
        self.N, self.n, self.d = p.N, p.n, p.d
        self.model = model
        self.qp = qp
        self.A, self.B, self.C = p.A, p.B, None
        if p.timeVarying:
            self.xLin = model.xStored[-1][0:self.N + 1, :]
            self.uLin = model.uStored[-1][0:self.N, :]
            self.identify()
        self.OldInput = np.zeros((1, 2))
        self.F, self.b = build_ineq(p)
        self.H, self.q = build_cost(p, self.OldInput)
        self.G, self.E, self.L = build_eq(p, self.A, self.B, self.C)
        self.xPred = []
        self.uPred = None
        self.timeStep = 0
        self.feasible = 1
        self.last_qp = None
        self.last_info = None

    def identify(self):
        # PC.py:140-145
        self.A, self.B, self.C = [], [], []
        for i in range(self.N):
            Ai, Bi, Ci = self.model.regressionAndLinearization(self.xLin[i], self.uLin[i])
            self.A.append(Ai)
            self.B.append(Bi)
            self.C.append(Ci)

    # hooks overridden by the LMPC
    def terminal_components(self, x0):
        return self.H, self.q, self.F, self.b, self.G, self.E, self.L

    def unpack(self, z):
        n, d, N = self.n, self.d, self.N
        self.xPred = z[0:n * (N + 1)].reshape(N + 1, n).copy()
        self.uPred = z[n * (N + 1):n * (N + 1) + d * N].reshape(N, d).copy()

    def feasible_state_input(self):
        self.zt = self.xPred[-1, :]
        self.zt_u = self.uPred[-1, :]

    def assemble(self, x0):
        """Everything solve() does before the QP call; returns OSQP-form data."""
        if self.p.timeVarying:
            self.identify()
            self.H, self.q = build_cost(self.p, self.OldInput)
            self.G, self.E, self.L = build_eq(self.p, self.A, self.B, self.C)
        H, q, F, b, G, E, L = self.terminal_components(x0)
        return osqp_form(H, q, F, b, G, np.add(np.dot(E, x0), L))

    def solve(self, x0):
        # PC.py:110-137
        P, q, A, l, u = self.assemble(x0)
        self.last_qp = (P, q, A, l, u)
        z, info = self.qp(P, q, A, l, u)
        self.last_info = info
        self.feasible = 1 if (info is None or info.get("status", 1) == 1) else 0
        self.Solution = z
        self.unpack(z)
        self.feasible_state_input()
        if self.p.timeVarying:
            self.xLin = np.vstack((self.xPred[1:, :], self.zt))
            self.uLin = np.vstack((self.uPred[1:, :], self.zt_u))
        self.OldInput = self.uPred[0, :]
        self.timeStep += 1


class OracleLMPC(OracleMPC):
    """PC.py:286-514."""

    def __init__(self, numSS_Points, numSS_it, QterminalSlack, params, model, qp=None):
        super().__init__(params, model, qp)
        self.numSS_Points, self.numSS_it = numSS_Points, numSS_it
        self.QterminalSlack = QterminalSlack
        self.OldInput = np.zeros((1, 2))
        self.xPred = []
        self.LapTime, self.SS, self.uSS, self.Qfun, self.SS_glob = [], [], [], [], []
        self.xStoredPredTraj, self.xStoredPredTraj_it = [], []
        self.uStoredPredTraj, self.uStoredPredTraj_it = [], []
        self.SSStoredPredTraj, self.SSStoredPredTraj_it = [], []
        self.zt = np.array([0.0, 0.0, 0.0, 0.0, 10.0, 0.0])
        self.it = 0
        self.F, self.b = build_ineq(params)
        self.H, self.q = build_cost(params, self.OldInput)

    # PC.py:478-514
    def select_points(self, lap, zt, num_points):
        x, u = self.SS[lap], self.uSS[lap]
        dist = la.norm(x - zt[None, :], 1, axis=1)
        i_min = int(np.argmin(dist))
        if i_min - num_points / 2 >= 0:
            idx = range(-int(num_points / 2) + i_min, int(num_points / 2) + i_min + 1)
        else:
            idx = range(i_min, i_min + int(num_points))
        idx = list(idx)
        pts, upts = x[idx, :].T, u[idx, :].T
        qsel = self.Qfun[lap][idx]
        L = self.model.map.TrackLength
        no_pred = isinstance(self.xPred, list) and len(self.xPred) == 0
        if no_pred:
            pass
        elif np.all((self.xPred[:, 4] > L) == False):  # noqa: E712  (kept as in PC.py:504)
            pass
        elif lap < self.it - 1:
            qsel = self.Qfun[lap][idx] + self.Qfun[lap][0]
        else:
            pred_curr = self.N - sum(self.xPred[:, 4] > L)
            qsel = self.Qfun[lap][idx] + self.timeStep + pred_curr
        self.last_sel_index.append((lap, i_min, idx[0]))
        return pts, upts, qsel

    # PC.py:386-416
    def terminal_components(self, x0):
        n, d = self.n, self.d
        L = self.model.map.TrackLength
        if self.zt[4] - x0[4] > L / 2:
            self.zt[4] = np.max([self.zt[4] - L, 0])
            self.xLin[4, -1] = self.xLin[4, -1] - L           # PC.py:394 quirk (row 4, last column)
        order = np.argsort(np.array(self.LapTime))
        SSs, Sx, Su, Qs = np.empty((n, 0)), np.empty((n, 0)), np.empty((d, 0)), np.empty((0))
        self.last_sel_index = []
        for jj in order[0:self.numSS_it]:
            pts, upts, qsel = self.select_points(jj, self.zt, self.numSS_Points / self.numSS_it + 1)
            Sx = np.append(Sx, pts[:, 1:], axis=1)
            Su = np.append(Su, upts[:, 1:], axis=1)
            SSs = np.append(SSs, pts[:, 0:-1], axis=1)
            Qs = np.append(Qs, qsel[0:-1], axis=0)
        self.Succ_SS_PointSelectedTot, self.Succ_uSS_PointSelectedTot = Sx, Su
        self.SS_PointSelectedTot, self.Qfun_SelectedTot = SSs, Qs
        F2, b2, G2, E2, L2, H2, q2 = add_safe_set(self.F, self.b, self.G, self.E, self.L, self.H, self.q,
                                                  n, self.N, SSs, Qs, self.QterminalSlack)
        return H2, q2, F2, b2, G2, E2, L2

    # PC.py:364-379
    def unpack(self, z):
        n, d, N = self.n, self.d, self.N
        i_u = n * (N + 1)
        i_s = i_u + d * N
        i_l = i_s + self.p.Fx.shape[0] * N
        i_t = i_l + self.SS_PointSelectedTot.shape[1]
        self.xPred = z[0:i_u].reshape(N + 1, n).copy()
        self.uPred = z[i_u:i_s].reshape(N, d).copy()
        self.slack = z[i_s:i_l]
        self.lambd = z[i_l:i_t]
        self.slackTerminal = z[i_t:]
        self.xStoredPredTraj_it.append(self.xPred)
        self.uStoredPredTraj_it.append(self.uPred)
        self.SSStoredPredTraj_it.append(self.SS_PointSelectedTot.T)

    # PC.py:382-384
    def feasible_state_input(self):
        self.zt = np.dot(self.Succ_SS_PointSelectedTot, self.lambd)
        self.zt_u = np.dot(self.Succ_uSS_PointSelectedTot, self.lambd)

    # PC.py:418-445
    def addTrajectory(self, x, u, x_glob):
        self.LapTime.append(x.shape[0])
        self.SS.append(x)
        self.SS_glob.append(x_glob)
        self.uSS.append(u)
        self.Qfun.append(rollout_cost(x, self.model.map.TrackLength))
        if self.it == 0:
            self.xLin = self.SS[self.it][1:self.N + 2, :]      # a VIEW of the lap (PC.py:432)
            self.uLin = self.uSS[self.it][1:self.N + 1, :]
        self.xStoredPredTraj.append(self.xStoredPredTraj_it)
        self.xStoredPredTraj_it = []
        self.uStoredPredTraj.append(self.uStoredPredTraj_it)
        self.uStoredPredTraj_it = []
        self.SSStoredPredTraj.append(self.SSStoredPredTraj_it)
        self.SSStoredPredTraj_it = []
        self.it += 1
        self.timeStep = 0

    # PC.py:466-476
    def addPoint(self, x, u):
        L = self.model.map.TrackLength
        j = self.it - 1
        self.SS[j] = np.append(self.SS[j], np.array([x + np.array([0, 0, 0, 0, L, 0])]), axis=0)
        self.uSS[j] = np.append(self.uSS[j], np.array([u]), axis=0)
        self.Qfun[j] = np.append(self.Qfun[j], self.Qfun[j][-1] - 1)
