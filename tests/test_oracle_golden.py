"""CPU tier: the oracle restatement against golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  Bit-exact where the arithmetic is a restatement."""
import numpy as np
import pytest
from oracle import vehicle, ftocp, ltv_model, osqp_port, kkt
from oracle.track import TrackTable
import replay


def test_track_table_and_curvature(gold, track):
    assert np.array_equal(track.PointAndTangent, gold["track_table"])
    assert track.TrackLength == float(gold["track_length"])
    for s, ref in zip(gold["curv_s"], gold["curv_val"]):
        if np.isnan(ref):
            with pytest.raises(Exception):
                track.curvature(s)
        else:
            assert track.curvature(s) == ref


def test_pid_lap_regenerates_bit_exact(gold, track):
    np.random.seed(0)
    x0 = np.array([0.5, 0, 0, 0, 0, 0])
    xP, uP, gP, _ = vehicle.closed_loop(track, [x0, x0], vehicle.PIDFollower(0.8), max_steps=200)
    assert np.array_equal(xP, gold["pid_x"][:200])
    assert np.array_equal(uP, gold["pid_u"][:200])
    assert np.array_equal(gP, gold["pid_glob"][:200])


def test_lti_sysid_and_rollout_cost(gold, track):
    A, B, _ = vehicle.ridge_sysid(gold["pid_x"], gold["pid_u"], 1e-7)
    assert np.array_equal(A, gold["lti_A"]) and np.array_equal(B, gold["lti_B"])
    assert np.array_equal(ftocp.rollout_cost(gold["pid_x"], track.TrackLength), gold["pid_Qfun"])


def test_lti_mpc_assembly(gold):
    mp, _ = ftocp.mpc_params(6, 2, 12, 0.8)
    mp.A, mp.B = gold["lti_A"], gold["lti_B"]
    for t in (0, 7):
        k = "lti_t%d_" % t
        c = ftocp.OracleMPC(mp, qp=None)
        c.OldInput = gold[k + "old"]
        c.H, c.q = ftocp.build_cost(mp, c.OldInput)       # what MPC.__init__/solve leave in place
        P, q, A, l, u = c.assemble(gold[k + "x0"])
        for mine, nm in zip((P, q, A, l, u), "PqAlu"):
            ref = gold[k + "qp_" + nm]
            fin = np.isfinite(ref)
            assert np.array_equal(fin, np.isfinite(mine))
            if nm == "q" and t == 7:
                # the reference rebuilds q only for LTV problems (PC.py:116-119): for the LTI MPC the
                # OldInput term stays at its constructor value (zeros) -> compare with that quirk
                c0 = ftocp.OracleMPC(mp, qp=None)
                assert np.array_equal(c0.q, ref)
            else:
                assert np.array_equal(mine[fin], ref[fin]), nm


def test_knn_indices_and_weights(gold, track):
    pm = ltv_model.LocalLTVModel(6, 2, track, 1)
    pm.addTrajectory(gold["pid_x"], gold["pid_u"])
    for p, idx, K in zip(gold["knn_probe"], gold["knn_idx"], gold["knn_K"]):
        i, k = pm.knn(p, 0)
        assert np.array_equal(i, idx) and np.array_equal(k, K)


@pytest.mark.parametrize("t", [0, 1, 20])
def test_ltv_regression_and_assembly(gold, track, t):
    k = "ltv_t%d_" % t
    pm = ltv_model.LocalLTVModel(6, 2, track, 1)
    pm.addTrajectory(gold["pid_x"], gold["pid_u"])
    _, ltv = ftocp.mpc_params(6, 2, 12, 0.8)
    ltv.timeVarying = True
    c = ftocp.OracleMPC(ltv, pm, qp=None)
    c.xLin, c.uLin, c.OldInput = gold[k + "xLin"], gold[k + "uLin"], gold[k + "old"]
    P, q, A, l, u = c.assemble(gold[k + "x0"])
    assert np.array_equal(np.array(c.A), gold[k + "A"])
    assert np.array_equal(np.array(c.B), gold[k + "B"])
    assert np.array_equal(np.array(c.C), gold[k + "C"])
    for mine, nm in zip((P, q, A, l, u), "PqAlu"):
        ref = gold[k + "qp_" + nm]
        fin = np.isfinite(ref)
        assert np.array_equal(mine[fin], ref[fin]), nm


@pytest.mark.parametrize("key", replay.LMPC_KEYS)
def test_lmpc_step_replay(gold, track, key):
    """Safe-set selection, Q-function shifts, LTV regression and the assembled QP of one LMPC.solve."""
    k = "lmpc_%d_%d_" % key
    c, x0 = replay.lmpc_from_snapshot(gold, key, track)
    P, q, A, l, u = c.assemble(x0)
    assert np.array_equal(c.SS_PointSelectedTot, gold[k + "SS_sel"])
    assert np.array_equal(c.Qfun_SelectedTot, gold[k + "Qfun_sel"])
    assert np.array_equal(c.Succ_SS_PointSelectedTot, gold[k + "Succ_SS"])
    assert np.array_equal(c.Succ_uSS_PointSelectedTot, gold[k + "Succ_uSS"])
    assert np.array_equal(np.array(c.A), gold[k + "A"])
    assert np.array_equal(np.array(c.C), gold[k + "C"])
    for mine, nm in zip((P, q, A, l, u), "PqAlu"):
        ref = gold[k + "qp_" + nm]
        fin = np.isfinite(ref)
        assert np.array_equal(mine[fin], ref[fin]), nm
    # solve and compare with what the reference loop obtained with the same back-end
    z, info = osqp_port.tight_qp(P, q, A, l, u)
    c.unpack(z)
    c.feasible_state_input()
    assert np.max(np.abs(c.xPred - gold[k + "xPred"])) < 1e-9
    assert np.max(np.abs(c.uPred - gold[k + "uPred"])) < 1e-9
    assert np.max(np.abs(c.zt - gold[k + "zt_out"])) < 1e-7


def test_track_global_position_against_reference_values(track):
    """oracle/track.py::global_position (Track.py:135-189) against values computed by the real reference
    (tests/golden/make_track_global.py): bit for bit."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "track_global.npz"))
    assert np.array_equal(g["table"], track.PointAndTangent)
    xy = np.array([track.global_position(float(a), float(b)) for a, b in zip(g["s"], g["ey"])])
    assert np.array_equal(xy, g["xy"])
