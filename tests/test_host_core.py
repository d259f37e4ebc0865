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


def test_properties_on_perturbed_problems(gold):
    """Solver-independent properties on randomly perturbed LTV-MPC problems (hypothesis): the returned point satisfies the KKT
    conditions of the QP the oracle assembles with the reference's code path (1e-6), hard input bounds hold, lane slacks are
    non-negative and equal max(0, violation) where they matter, and a solve is a pure function of its inputs."""
    from hypothesis import given, settings, strategies as st
    N = 12
    _, ltv = ftocp.mpc_params(6, 2, N, 0.8)
    ltv.timeVarying = True
    c = hc.make_const(ltv)
    k = "ltv_t20_"
    A0, B0, C0 = gold[k + "A"], gold[k + "B"], gold[k + "C"]
    F, bb = ftocp.build_ineq(ltv)

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(dx=st.lists(st.floats(-1, 1), min_size=6, max_size=6), du=st.lists(st.floats(-1, 1), min_size=2, max_size=2),
           scale=st.floats(0.9, 1.1), ey=st.floats(-2.5, 2.5))
    def run(dx, du, scale, ey):
        x0 = gold[k + "x0"] + np.array(dx) * np.array([0.2, 0.05, 0.2, 0.05, 0.0, 0.1])
        x0[5] = ey                                           # up to 0.5 outside the 2.0 lane half-width: soft constraint active
        uold = np.clip(gold[k + "old"].ravel() + 0.2 * np.array(du), [-0.45, -9.0], [0.45, 9.0])
        A, B, C = A0 * scale, B0, C0
        sol = hc.solve(c, N, hc.pack_abc(A, B, C, N), x0, uold)
        assert sol["status"] == 1 and sol["iters"] <= 30
        again = hc.solve(c, N, hc.pack_abc(A, B, C, N), x0, uold)
        assert np.array_equal(sol["x"], again["x"]) and np.array_equal(sol["u"], again["u"])
        assert np.all(np.abs(sol["u"][:, 0]) <= 0.5 + 1e-9) and np.all(np.abs(sol["u"][:, 1]) <= 10.0 + 1e-9)
        assert np.all(sol["s"] >= -1e-12)
        viol = np.maximum(np.abs(sol["x"][:N, 5]) - 2.0, 0.0)
        assert np.max(np.abs(sol["s"].reshape(N, 2).sum(axis=1) - viol)) < 1e-6       # slack = lane violation, nothing more
        H, q = ftocp.build_cost(ltv, uold); G, E, L = ftocp.build_eq(ltv, list(A), list(B), list(C))
        P, q, Am, l, u = ftocp.osqp_form(H, q, F, bb, G, E @ x0 + L)
        z = np.concatenate([sol["x"].ravel(), sol["u"].ravel(), sol["s"].ravel()])
        y = kkt.dual_from_primal(P, q, Am, l, u, z, tol=1e-6)
        r = kkt.residuals(P, q, Am, l, u, z, y)
        assert r["r_prim"] < 1e-6 and r["r_dual"] < 1e-6, r
    run()


def test_lmpc_properties_on_perturbed_problems(gold, track):
    """LMPC-type QPs (simplex terminal block) around a golden step: lambda lies on the simplex, the terminal slack is what the
    terminal equality says, and the point satisfies the KKT conditions of the oracle-assembled QP (reference code path)."""
    from hypothesis import given, settings, strategies as st
    N = 12
    _, _, _, _, Qts, lp = ftocp.lmpc_params(track, N)
    lp.timeVarying = True
    c = hc.make_const(lp, Qts)
    k = "lmpc_5_90_"
    A, B, C, SS, Qf = gold[k + "A"], gold[k + "B"], gold[k + "C"], gold[k + "SS_sel"], gold[k + "Qfun_sel"]
    F, bb = ftocp.build_ineq(lp)

    @settings(max_examples=15, deadline=None, derandomize=True)
    @given(dx=st.lists(st.floats(-1, 1), min_size=6, max_size=6), dq=st.lists(st.floats(0, 3), min_size=4, max_size=4))
    def run(dx, dq):
        x0 = gold[k + "x0"] + np.array(dx) * np.array([0.1, 0.03, 0.1, 0.03, 0.05, 0.05])
        uold = gold[k + "OldInput"].ravel()
        Q = Qf + np.repeat(np.array(dq), 12)                    # shift the cost-to-go of each selected lap
        sol = hc.solve(c, N, hc.pack_abc(A, B, C, N), x0, uold, SS, Q)
        assert sol["status"] == 1 and sol["iters"] <= 30
        lam = sol["lam"]
        assert lam.min() >= -1e-12 and abs(lam.sum() - 1.0) < 1e-8
        xi = SS @ lam - sol["x"][-1]
        G, E, L = ftocp.build_eq(lp, list(A), list(B), list(C)); H, q = ftocp.build_cost(lp, uold[None, :])
        F2, b2, G2, E2, L2, H2, q2 = ftocp.add_safe_set(F, bb, G, E, L, H, q, 6, N, SS, Q, Qts)
        P, qq, Am, l, u = ftocp.osqp_form(H2, q2, F2, b2, G2, E2 @ x0 + L2)
        z = np.concatenate([sol["x"].ravel(), sol["u"].ravel(), sol["s"].ravel(), lam, xi])
        y = kkt.dual_from_primal(P, qq, Am, l, u, z, tol=1e-6)
        r = kkt.residuals(P, qq, Am, l, u, z, y)
        assert r["r_prim"] < 1e-6 and r["r_dual"] < 1e-6, r
    run()


@pytest.mark.parametrize("N,cap_mean,cap_max", [(6, 6.4, 11), (12, 7.3, 16), (24, 8.0, 17), (48, 9.6, 20)])
def test_iteration_budget_on_the_baseline_workloads(N, cap_mean, cap_max):
    """The starting point / step rule constants (LMPC_TUNE_* in ftocp_pdip.cuh) were chosen by sweeping them through this
    emulation on the BASELINE LTV-MPC workloads; this pins the result: every QP solved to the default tolerance, mean and
    maximum interior-point iterations within the swept optimum's margin (round-1 constants: 7.84 / 8.60 / 8.96 / 10.96 mean)."""
    B = 256
    x0, uold, abc = workloads.ltv_mpc_batch(B, N=N)
    c = hc.make_const(rp.mpc_params(N))
    its = []
    for b in range(B):
        sol = hc.solve(c, N, abc[b], x0[b], uold[b])
        assert sol["status"] == 1 and max(sol["r_prim"], sol["r_dual"]) <= 1.000001e-9 and sol["gap"] <= 1.000001e-11, (N, b)
        its.append(sol["iters"])
    assert np.mean(its) <= cap_mean and max(its) <= cap_max, (N, np.mean(its), max(its))


@pytest.mark.parametrize("N,B", [(12, 16), (48, 3)])
def test_numpy_spec_follows_the_kernel_with_its_constants(N, B):
    """oracle/pdip_model.py is the executable specification of the kernel: with the kernel's constants
    (pdip_model.kernel_constants) it takes the same number of iterations and lands on the same point as the compiled core."""
    from oracle import pdip_model as pm
    x0, uold, abc = workloads.ltv_mpc_batch(B, N=N)
    p = rp.mpc_params(N)
    c = hc.make_const(p)
    for b in range(B):
        A = abc[b][:, 0:36].reshape(N, 6, 6); Bm = abc[b][:, 36:48].reshape(N, 6, 2); C = abc[b][:, 48:54]
        ref = pm.solve(pm.from_params(p, A, Bm, C, x0[b], uold[b]), eps=1e-9, eps_gap=1e-11, max_iter=40, **pm.kernel_constants(N, False))
        sol = hc.solve(c, N, abc[b], x0[b], uold[b])
        assert ref["status"] == 1 and sol["status"] == 1 and ref["iters"] == sol["iters"], (b, ref["iters"], sol["iters"])
        assert np.max(np.abs(sol["x"] - ref["x"])) < 1e-10 and np.max(np.abs(sol["u"] - ref["u"])) < 1e-10
