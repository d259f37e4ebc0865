"""CPU tier: the CUDA solver core (racinglmpc_b200/csrc/ftocp_pdip.cuh) compiled as a 1-lane host
emulation, against the reference-pinned golden solutions and the KKT checker.  This checks the
kernel's arithmetic without a GPU; the warp-parallel execution itself is covered by the -m gpu tests."""
import os
import numpy as np
import pytest
import hostcore as hc
from oracle import ftocp, kkt
from racinglmpc_b200 import workloads, reference_params as rp
import replay


def test_lti_and_ltv_against_golden(gold):
    mp, _ = ftocp.mpc_params(6, 2, 12, 0.8)
    c = hc.make_const(mp)
    for t in (0, 7):
        k = "lti_t%d_" % t
        old = gold[k + "old"] if t == 0 else np.zeros(2)     # LTI MPC never refreshes q (PC.py:116-119)
        sol = hc.solve(c, 12, hc.pack_abc(gold["lti_A"], gold["lti_B"], None, 12), gold[k + "x0"], old)
        assert sol["status"] == 1 and sol["iters"] <= 20
        assert np.max(np.abs(sol["x"] - gold[k + "xPred"])) < 1e-8
        assert np.max(np.abs(sol["u"] - gold[k + "uPred"])) < 1e-8
    for t in (0, 1, 20):
        k = "ltv_t%d_" % t
        sol = hc.solve(c, 12, hc.pack_abc(gold[k + "A"], gold[k + "B"], gold[k + "C"], 12), gold[k + "x0"], gold[k + "old"])
        assert sol["status"] == 1
        assert np.max(np.abs(sol["x"] - gold[k + "xPred"])) < 1e-8
        assert np.max(np.abs(sol["u"] - gold[k + "uPred"])) < 1e-8


@pytest.mark.parametrize("key", replay.LMPC_KEYS)
def test_lmpc_against_golden(gold, track, key):
    k = "lmpc_%d_%d_" % key
    _, _, _, _, Qts, lp = ftocp.lmpc_params(track, 12)
    c = hc.make_const(lp, Qts)
    sol = hc.solve(c, 12, hc.pack_abc(gold[k + "A"], gold[k + "B"], gold[k + "C"], 12), gold[k + "x0"],
                   gold[k + "OldInput"], gold[k + "SS_sel"], gold[k + "Qfun_sel"])
    assert sol["status"] == 1 and sol["iters"] <= 25
    assert max(sol["r_prim"], sol["r_dual"], sol["gap"]) <= 1.000001e-8
    assert np.max(np.abs(sol["x"] - gold[k + "xPred"])) < 1e-6
    assert np.max(np.abs(sol["u"] - gold[k + "uPred"])) < 1e-6
    # solver-independent check against the matrices the REFERENCE assembled
    P, q, A, l, u = [gold[k + "qp_" + ch] for ch in "PqAlu"]
    z = np.concatenate([sol["x"].ravel(), sol["u"].ravel(), sol["s"].ravel(), sol["lam"],
                        gold[k + "SS_sel"] @ sol["lam"] - sol["x"][-1]])
    y = kkt.dual_from_primal(P, q, A, l, u, z, tol=1e-6)
    r = kkt.residuals(P, q, A, l, u, z, y)
    assert r["r_prim"] < 1e-6 and r["r_dual"] < 1e-6, r


@pytest.mark.parametrize("N", [6, 12, 24, 48])
def test_horizon_sweep_workload(N):
    """configs[4] shapes: LTV-MPC QPs at N in {6,12,24,48}; KKT-checked against an oracle assembly."""
    from oracle import osqp_port
    x0, uold, abc = workloads.ltv_mpc_batch(6, N=N)
    _, ltv = ftocp.mpc_params(6, 2, N, 0.8)
    ltv.timeVarying = True
    c = hc.make_const(ltv)
    for b in range(6):
        sol = hc.solve(c, N, abc[b], x0[b], uold[b])
        assert sol["status"] == 1 and sol["iters"] <= 25, (N, b, sol["iters"])
        A = abc[b][:, 0:36].reshape(N, 6, 6); B = abc[b][:, 36:48].reshape(N, 6, 2); C = abc[b][:, 48:54]
        H, q = ftocp.build_cost(ltv, uold[b]); F, bb = ftocp.build_ineq(ltv); G, E, L = ftocp.build_eq(ltv, list(A), list(B), list(C))
        P, q, Am, l, u = ftocp.osqp_form(H, q, F, bb, G, E @ x0[b] + L)
        z, info = osqp_port.tight_qp(P, q, Am, l, u)
        n = 6 * (N + 1)
        assert np.max(np.abs(sol["x"].ravel() - z[:n])) < 1e-6
        assert np.max(np.abs(sol["u"].ravel() - z[n:n + 2 * N])) < 1e-6


def test_recentring_step_unsticks_boundary_riding_instances(track):
    """Closed-loop LMPC QPs dumped from a device rollout (tools/scratch/dbg_rollout.py) on which the solver used to ride the
    central-path neighbourhood boundary with tiny steps until max_iter (6 instances) or for 30+ iterations (2): with the
    recentring pass (RECENTRE_AFTER) each is solved to 1e-9 in at most 20 iterations, and the NumPy model of the kernel
    (oracle/pdip_model.py) agrees on the solution."""
    from oracle import pdip_model as pm
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "stalled_lmpc_qps.npz"))
    _, _, _, _, Qts, lp = ftocp.lmpc_params(track, 12)
    c = hc.make_const(lp, Qts)
    for i in range(8):
        abc, x0, uold, SS, Qf = (d["q%d_%s" % (i, k)] for k in ("abc", "x0", "uold", "SS", "Qf"))
        sol = hc.solve(c, 12, abc, x0, uold, SS, Qf)
        assert sol["status"] == 1 and sol["iters"] <= 20, (i, sol["status"], sol["iters"])
        assert max(sol["r_prim"], sol["r_dual"]) <= 1.000001e-9 and sol["gap"] <= 1.000001e-11
        A = abc[:, 0:36].reshape(12, 6, 6); B = abc[:, 36:48].reshape(12, 6, 2); C = abc[:, 48:54]
        ref = pm.solve(pm.from_params(lp, A, B, C, x0, uold, SS, Qf, Qts), eps=1e-9, eps_gap=1e-11, max_iter=40)
        assert ref["status"] == 1
        assert np.max(np.abs(sol["x"] - ref["x"])) < 1e-6 and np.max(np.abs(sol["u"] - ref["u"])) < 1e-6
