"""CPU tier: the QP-solver oracle is pinned by (a) the solver-independent KKT checker and
(b) agreement between two independent algorithms (OSQP-algorithm ADMM port vs PDIP model)."""
import numpy as np
import pytest
from oracle import osqp_port, kkt, pdip_model as pm, ftocp
import replay


def _qp(gold, k):
    return [gold[k + "qp_" + c] for c in "PqAlu"]


@pytest.mark.parametrize("k", ["lti_t0_", "ltv_t20_", "lmpc_4_60_", "lmpc_5_0_"])
def test_osqp_port_reference_settings_kkt(gold, k):
    """Defaults + polish (what PC.py:275 asks for): solved status and small KKT residuals."""
    P, q, A, l, u = _qp(gold, k)
    z, info, y = osqp_port.solve(P, q, A, l, u)
    assert info["status"] == 1
    r = kkt.residuals(P, q, A, l, u, z, y)
    # OSQP's own stopping rule is eps_abs = eps_rel = 1e-3 (relative to the iterate norms)
    assert r["r_prim"] < 5e-2 and r["r_dual"] < 2.0   # unpolished LMPC iterates stop at 1e-3 RELATIVE to |q| ~ 1e2..1e3


@pytest.mark.parametrize("k", ["lti_t0_", "lti_t7_", "ltv_t0_", "ltv_t20_"] + ["lmpc_%d_%d_" % kk for kk in replay.LMPC_KEYS])
def test_tight_oracle_kkt_1e6(gold, k):
    P, q, A, l, u = _qp(gold, k)
    z, info, y = osqp_port.solve(P, q, A, l, u, eps_abs=1e-9, eps_rel=1e-9, max_iter=400000, polish_strict=1)
    r = kkt.residuals(P, q, A, l, u, z, y)
    assert max(r.values()) < 1e-6, r


@pytest.mark.parametrize("key", replay.LMPC_KEYS)
def test_pdip_model_matches_osqp_oracle(gold, track, key):
    k = "lmpc_%d_%d_" % key
    _, _, _, _, Qts, lp = ftocp.lmpc_params(track, 12)
    qp = pm.from_params(lp, gold[k + "A"], gold[k + "B"], gold[k + "C"], gold[k + "x0"], gold[k + "OldInput"],
                        gold[k + "SS_sel"], gold[k + "Qfun_sel"], Qts)
    sol = pm.solve(qp, eps=1e-10, max_iter=40)
    assert sol["status"] == 1
    assert np.max(np.abs(sol["x"] - gold[k + "xPred"])) < 1e-7
    assert np.max(np.abs(sol["u"] - gold[k + "uPred"])) < 1e-7
    P, q, A, l, u = _qp(gold, k)
    z = pm.pack(qp, sol)
    zo = np.concatenate([gold[k + "xPred"].ravel(), gold[k + "uPred"].ravel()])
    assert abs(kkt.objective(P, q, z) - info_obj(P, q, A, l, u)) < 1e-6


def info_obj(P, q, A, l, u):
    z, _ = osqp_port.tight_qp(P, q, A, l, u)
    return kkt.objective(P, q, z)


def test_batch_api_matches_single(gold):
    P, q, A, l, u = _qp(gold, "ltv_t0_")
    patP, patA = osqp_port.csc_pattern(np.ones_like(P, bool), np.ones_like(A, bool))
    Px = np.stack([osqp_port.gather_values(P, *patP)] * 3)
    Ax = np.stack([osqp_port.gather_values(A, *patA)] * 3)
    qq, ll, uu = np.stack([q] * 3), np.stack([l] * 3), np.stack([u] * 3)
    x, infos, _ = osqp_port.solve_batch(patP, patA, Px, qq, Ax, ll, uu, nthreads=2)
    z, _, _ = osqp_port.solve(P, q, A, l, u)
    assert all(i["status"] == 1 for i in infos)
    assert np.max(np.abs(x - z[None, :])) < 1e-8
