"""GPU tier: drop-in boundary.  The shim modules racinglmpc_b200/compat/{PredictiveControllers,PredictiveModel}.py
are imported BY THE REFERENCE'S MODULE NAMES and driven through the reference's own call sequence
(src/main.py:39-120 with N = 12): PID lap -> LTI sys-id -> MPC lap -> LTV-MPC lap -> LMPC laps, with the
harness (Simulator.sim, PID, Regression, Map) taken from the oracle's bit-exact restatement because
/root/reference does not exist on the GPU box.  Each closed loop is compared with the same loop driven by the
oracle controllers (reference arithmetic + OSQP-algorithm solver at 1e-9)."""
import os
import sys
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")


@pytest.fixture(scope="module")
def shim():
    sys.path.insert(0, os.path.join(ROOT, "racinglmpc_b200", "compat"))
    import PredictiveControllers as PC
    import PredictiveModel as PM
    yield PC, PM
    sys.path.remove(os.path.join(ROOT, "racinglmpc_b200", "compat"))


def _init_mpc_params(PC, n, d, N, vt):
    # values of initMPCParams (initControllerParameters.py:4-26), built with the shim's MPCParams
    from racinglmpc_b200 import reference_params as rp
    q = rp.mpc_params(N, vt)
    mk = lambda: PC.MPCParams(n=n, d=d, N=N, Q=q.Q, R=q.R, Fx=q.Fx, bx=np.array([[2.], [2.]]), Fu=q.Fu,
                              bu=np.array([[0.5], [0.5], [10.0], [10.0]]), xRef=q.xRef, slacks=True, Qslack=q.Qslack)
    return mk(), mk()


def _init_lmpc_params(PC, track, N):
    from racinglmpc_b200 import reference_params as rp
    numSS_it, numSS_Points, Laps, TimeLMPC, Qts, q = rp.lmpc_params(N)
    p = PC.MPCParams(n=6, d=2, N=N, Q=q.Q, R=q.R, dR=q.dR, Fx=q.Fx, bx=np.array([[track.halfWidth], [track.halfWidth]]),
                     Fu=q.Fu, bu=np.array([[0.5], [0.5], [10.0], [10.0]]), slacks=True, Qslack=q.Qslack)
    return numSS_it, numSS_Points, Laps, TimeLMPC, Qts, p


def test_mpcparams_is_frozen_like_the_reference(shim):
    PC, _ = shim
    p = PC.MPCParams(n=6, d=2, N=12, Q=np.eye(6), R=np.eye(2))
    p.A = np.eye(6)                 # assignable field (main.py:76)
    p.timeVarying = True            # main.py:91
    with pytest.raises(TypeError):
        p.notAField = 1
    assert p.Qf.shape == (6, 6) and p.dR.shape == (2,) and p.xRef.shape == (6,)


def test_main_py_sequence_against_oracle(shim, gold, track):
    _need_gpu()
    PC, PM = shim
    from oracle import vehicle, ftocp, ltv_model, osqp_port
    N, n, d = 12, 6, 2
    x0 = np.array([0.5, 0, 0, 0, 0, 0.0])
    xS = [x0, x0]
    xPID, uPID, gPID = gold["pid_x"], gold["pid_u"], gold["pid_glob"]          # main.py:65-66 (seed 0)

    def both(make_gpu, make_oracle, lmpc=False, steps=None, xstart=xS):
        out = []
        for make in (make_gpu, make_oracle):
            np.random.seed(123)
            ctrl = make()
            out.append((ctrl,) + vehicle.closed_loop(track, xstart, ctrl, multi_lap=not lmpc, is_lmpc=lmpc, max_steps=steps))
        return out

    # ---- LTI MPC (main.py:72-80)
    A, B, _ = vehicle.ridge_sysid(xPID, uPID, 1e-7)
    mpcParam, ltvParam = _init_mpc_params(PC, n, d, N, 0.8)
    mpcParam.A, mpcParam.B = A, B
    omp, oltv = ftocp.mpc_params(n, d, N, 0.8)
    omp.A, omp.B = A, B
    (g, gx, gu, _, _), (o, ox, ou, _, _) = both(lambda: PC.MPC(mpcParam), lambda: ftocp.OracleMPC(omp, qp=osqp_port.tight_qp), steps=60)
    assert np.max(np.abs(gx - ox)) < 1e-6 and np.max(np.abs(gu - ou)) < 1e-6

    # ---- LTV MPC (main.py:86-94)
    def mk_gpu_ltv():
        pm = PM.PredictiveModel(n, d, track, 1)
        pm.addTrajectory(xPID, uPID)
        ltvParam.timeVarying = True
        return PC.MPC(ltvParam, pm)

    def mk_or_ltv():
        pm = ltv_model.LocalLTVModel(n, d, track, 1)
        pm.addTrajectory(xPID, uPID)
        oltv.timeVarying = True
        return ftocp.OracleMPC(oltv, pm, qp=osqp_port.tight_qp)
    (g, gx, gu, _, _), (o, ox, ou, _, _) = both(mk_gpu_ltv, mk_or_ltv, steps=60)
    assert np.max(np.abs(gx - ox)) < 1e-6 and np.max(np.abs(gu - ou)) < 1e-6
    assert np.max(np.abs(g.xLin - o.xLin)) < 1e-6

    # ---- LMPC, two laps (main.py:99-120)
    numSS_it, numSS_Points, Laps, _, Qts, lmpcPar = _init_lmpc_params(PC, track, N)
    _, _, _, _, oQts, olmpcPar = ftocp.lmpc_params(track, N)
    lmpcPar.timeVarying = True
    olmpcPar.timeVarying = True

    def run(make_pm, make_lmpc, copies):
        np.random.seed(7)
        xp, up, gp = copies
        pm = make_pm()
        for _ in range(4):
            pm.addTrajectory(xp, up)
        lm = make_lmpc(pm)
        for _ in range(4):
            lm.addTrajectory(xp, up, gp)
        xs_lap = xS
        laps = []
        for it in range(numSS_it, numSS_it + 2):
            xl, ul, gl, xs_lap = vehicle.closed_loop(track, xs_lap, lm, multi_lap=False, is_lmpc=True, max_steps=400)
            lm.addTrajectory(xl, ul, gl)
            pm.addTrajectory(xl, ul)
            laps.append((xl, ul))
        return lm, laps
    glm, glaps = run(lambda: PM.PredictiveModel(n, d, track, 4),
                     lambda pm: PC.LMPC(numSS_Points, numSS_it, Qts, lmpcPar, pm), (xPID.copy(), uPID.copy(), gPID.copy()))
    olm, olaps = run(lambda: ltv_model.LocalLTVModel(n, d, track, 4),
                     lambda pm: ftocp.OracleLMPC(numSS_Points, numSS_it, oQts, olmpcPar, pm, qp=osqp_port.tight_qp),
                     (xPID.copy(), uPID.copy(), gPID.copy()))
    for (gxl, gul), (oxl, oul) in zip(glaps, olaps):
        assert gxl.shape == oxl.shape                                   # same lap time
        assert np.max(np.abs(gxl - oxl)) < 1e-5 and np.max(np.abs(gul - oul)) < 1e-5
    # attributes main.py / plot.py read
    assert glm.it == olm.it == 6
    assert [q[0] for q in glm.Qfun] == [q[0] for q in olm.Qfun]          # main.py:120,127
    assert len(glm.SS) == 6 and glm.SS[4].shape == olm.SS[4].shape       # laps grown by addPoint
    assert np.array_equal(glm.LapTime, olm.LapTime)
    assert len(glm.xStoredPredTraj[4]) == glaps[0][0].shape[0] and len(glm.xStoredPredTraj[5]) == glaps[1][0].shape[0]
    assert glm.SSStoredPredTraj[5][0].shape == (numSS_Points, 6)
    # the device copy of a grown lap equals the host mirror
    xdev, udev, qdev = glm._engine.get_lap(0, 4)
    assert np.array_equal(xdev, glm.SS[4]) and np.array_equal(qdev, glm.Qfun[4])
