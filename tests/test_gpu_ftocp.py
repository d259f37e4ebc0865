"""GPU tier (-m gpu): the CUDA path through the C ABI against the oracle.

Tolerances (fp64 end to end): primal parity with the reference-pinned golden solutions and with the
OSQP-algorithm oracle driven to 1e-9: |dz| <= 1e-6; solver-reported residuals (unscaled inf-norm)
<= 1e-8; host-side KKT check against the matrices the REFERENCE assembled: <= 1e-6."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from racinglmpc_b200 import BatchedFTOCP, pack_abc, workloads, reference_params as rp   # noqa: E402
from oracle import ftocp, osqp_port, kkt                                                 # noqa: E402
import replay                                                                            # noqa: E402


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")


def test_golden_lti_ltv(gold):
    _need_gpu()
    keys = [("lti_t0_", None), ("ltv_t0_", "ltv"), ("ltv_t1_", "ltv"), ("ltv_t20_", "ltv")]
    B = len(keys)
    x0 = np.zeros((B, 6)); uold = np.zeros((B, 2)); abc = np.zeros((B, 12, 54))
    for i, (k, kind) in enumerate(keys):
        x0[i] = gold[k + "x0"]
        uold[i] = gold[k + "old"]
        if kind is None:
            abc[i] = pack_abc(np.tile(gold["lti_A"], (12, 1, 1)), np.tile(gold["lti_B"], (12, 1, 1)))
        else:
            abc[i] = pack_abc(gold[k + "A"], gold[k + "B"], gold[k + "C"])
    s = BatchedFTOCP(rp.mpc_params(12), batch=B)
    o = s.solve(x0, uold, abc)
    assert np.all(o["status"] == 1), o["status"]
    assert o["resid"].max() <= 1.000001e-8
    for i, (k, _) in enumerate(keys):
        assert np.max(np.abs(o["xPred"][i] - gold[k + "xPred"])) < 1e-6
        assert np.max(np.abs(o["uPred"][i] - gold[k + "uPred"])) < 1e-6
    assert s.kernel_launches == 1
    s.close()


def test_golden_lmpc_steps(gold, track):
    _need_gpu()
    keys = replay.LMPC_KEYS
    B = len(keys)
    numSS_it, numSS_Points, _, _, Qts, par = rp.lmpc_params(12)
    x0 = np.stack([gold["lmpc_%d_%d_x0" % k] for k in keys])
    uold = np.stack([gold["lmpc_%d_%d_OldInput" % k].ravel() for k in keys])
    abc = np.stack([pack_abc(gold["lmpc_%d_%d_A" % k], gold["lmpc_%d_%d_B" % k], gold["lmpc_%d_%d_C" % k]) for k in keys])
    SS = np.stack([gold["lmpc_%d_%d_SS_sel" % k] for k in keys])
    Qf = np.stack([gold["lmpc_%d_%d_Qfun_sel" % k] for k in keys])
    SuS = np.stack([gold["lmpc_%d_%d_Succ_SS" % k] for k in keys])
    SuU = np.stack([gold["lmpc_%d_%d_Succ_uSS" % k] for k in keys])
    s = BatchedFTOCP(par, batch=B, numSS_Points=numSS_Points, numSS_it=numSS_it, QterminalSlack=Qts)
    o = s.solve(x0, uold, abc, SS, Qf, SuS, SuU)
    assert np.all(o["status"] == 1), (o["status"], o["iters"])
    assert o["resid"].max() <= 1.000001e-9
    for i, k in enumerate(keys):
        kk = "lmpc_%d_%d_" % k
        assert np.max(np.abs(o["xPred"][i] - gold[kk + "xPred"])) < 1e-6
        assert np.max(np.abs(o["uPred"][i] - gold[kk + "uPred"])) < 1e-6
        assert np.max(np.abs(o["zt"][i] - gold[kk + "zt_out"])) < 1e-5
        assert np.max(np.abs(o["zt_u"][i] - gold[kk + "ztu_out"])) < 1e-5
        # KKT against the reference-assembled QP (solver independent)
        P, q, A, l, u = [gold[kk + "qp_" + ch] for ch in "PqAlu"]
        z = np.concatenate([o["xPred"][i].ravel(), o["uPred"][i].ravel(), o["slack"][i], o["lambd"][i], o["slackTerminal"][i]])
        y = kkt.dual_from_primal(P, q, A, l, u, z, tol=1e-6)
        r = kkt.residuals(P, q, A, l, u, z, y)
        assert r["r_prim"] < 1e-6 and r["r_dual"] < 1e-6, (k, r)
        assert abs(kkt.objective(P, q, z) - kkt.objective(P, q, np.concatenate([gold[kk + "xPred"].ravel(), gold[kk + "uPred"].ravel(), z[102:]]))) < 1e-3
    s.close()


def _oracle_solution(par, N, abc_b, x0_b, uold_b):
    A = abc_b[:, 0:36].reshape(N, 6, 6); B = abc_b[:, 36:48].reshape(N, 6, 2); C = abc_b[:, 48:54]
    par.timeVarying = True
    H, q = ftocp.build_cost(par, uold_b)
    F, bb = ftocp.build_ineq(par)
    G, E, L = ftocp.build_eq(par, list(A), list(B), list(C))
    P, q, Am, l, u = ftocp.osqp_form(H, q, F, bb, G, E @ x0_b + L)
    z, info = osqp_port.tight_qp(P, q, Am, l, u)
    return z, (P, q, Am, l, u)


@pytest.mark.parametrize("N", [6, 12, 24, 48])
def test_ltv_batch_vs_oracle(N):
    """configs[1]/[4]: batched LTV-MPC QPs; a sample is compared with the oracle, all are status-checked."""
    _need_gpu()
    B = 512 if N == 12 else 64
    x0, uold, abc = workloads.ltv_mpc_batch(B, N=N)
    par = rp.mpc_params(N)
    s = BatchedFTOCP(par, batch=B)
    o = s.solve(x0, uold, abc)
    assert np.all(o["status"] == 1), np.unique(o["status"], return_counts=True)
    assert o["resid"].max() <= 1.000001e-8
    assert o["iters"].max() <= 30
    opar = ftocp.mpc_params(6, 2, N, 0.8)[1]
    for b in range(0, B, max(B // 16, 1)):
        z, _ = _oracle_solution(opar, N, abc[b], x0[b], uold[b])
        n = 6 * (N + 1)
        assert np.max(np.abs(o["xPred"][b].ravel() - z[:n])) < 1e-6, b
        assert np.max(np.abs(o["uPred"][b].ravel() - z[n:n + 2 * N])) < 1e-6, b
    s.close()


def test_full_batch_properties():
    """configs[1] at full size (B=4096): determinism, batch-permutation invariance, dynamics feasibility."""
    _need_gpu()
    B, N = 4096, 12
    x0, uold, abc = workloads.ltv_mpc_batch(B, N=N)
    s = BatchedFTOCP(rp.mpc_params(N), batch=B)
    o1 = {k: v.copy() for k, v in s.solve(x0, uold, abc).items()}
    o2 = s.solve(x0, uold, abc)
    assert np.all(o1["status"] == 1)
    for k in ("xPred", "uPred", "iters"):
        assert np.array_equal(o1[k], o2[k]), k          # bit-identical re-run
    perm = np.random.default_rng(0).permutation(B)
    o3 = s.solve(x0[perm], uold[perm], abc[perm])
    assert np.array_equal(o3["uPred"], o1["uPred"][perm])  # an instance's result does not depend on its slot
    # dynamics hold: x_{k+1} = A x_k + B u_k + C
    A = abc[:, :, 0:36].reshape(B, N, 6, 6); Bm = abc[:, :, 36:48].reshape(B, N, 6, 2); C = abc[:, :, 48:54]
    pred = np.einsum("bkij,bkj->bki", A, o1["xPred"][:, :-1]) + np.einsum("bkij,bkj->bki", Bm, o1["uPred"]) + C
    assert np.max(np.abs(pred - o1["xPred"][:, 1:])) < 1e-9
    assert np.max(np.abs(o1["xPred"][:, 0] - x0)) == 0.0
    # input box respected
    assert np.all(np.abs(o1["uPred"][:, :, 0]) <= 0.5 + 1e-9) and np.all(np.abs(o1["uPred"][:, :, 1]) <= 10 + 1e-9)
    s.close()


def test_device_pointer_entry_matches_host_entry():
    _need_gpu()
    B, N = 256, 12
    x0, uold, abc = workloads.ltv_mpc_batch(B, N=N)
    s = BatchedFTOCP(rp.mpc_params(N), batch=B)
    oh = s.solve(x0, uold, abc)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    dx0, du, dabc = t(x0), t(uold), t(abc)
    xP = torch.zeros(B, N + 1, 6, dtype=torch.float64, device=dev)
    uP = torch.zeros(B, N, 2, dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
    rs = torch.zeros(B, 3, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    s.solve_dev(dx0, du, dabc, N * 54, 54, xP, uP, st, it, rs)
    s.sync()
    assert np.array_equal(uP.cpu().numpy(), oh["uPred"])
    assert np.array_equal(xP.cpu().numpy(), oh["xPred"])
    s.close()


def test_async_host_entry_slots_match_the_synchronous_call():
    """lmpc_solve_mpc_host_async on buffer sets 0 .. 3 (up to four batches in flight) returns exactly what the synchronous
    entry point returns for the same inputs, also when a slot is reused; host arrays from the library's pinned allocator
    (lmpc_host_alloc) and from torch's."""
    _need_gpu()
    from racinglmpc_b200 import _native as nat
    B, N = 256, 12
    x0, uold, abc = workloads.ltv_mpc_batch(B, N=N)
    s = BatchedFTOCP(rp.mpc_params(N), batch=B)
    ref = s.solve(x0, uold, abc)
    tpin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    lpin = lambda a: nat.pinned_like(np.ascontiguousarray(a))
    try:
        nat.pinned_empty(16)
        have_alloc = True
    except nat.NativeError:            # a host that refuses to register memory (locked-memory limit): the entry points still
        have_alloc, lpin = False, tpin  # take any pinned or pageable array; only the allocator's own checks are skipped
    perm = [np.arange(B), np.arange(B)[::-1], np.roll(np.arange(B), 7), np.roll(np.arange(B), 100)[::-1]]
    pins = [tpin, lpin, lpin, tpin]
    ins = [(pins[i](x0[perm[i]]), pins[i](uold[perm[i]]), pins[i](abc[perm[i]])) for i in range(4)]
    outs = [{k: pins[i](v) for k, v in s.alloc_outputs(False).items()} for i in range(4)]
    assert all(np.array_equal(ins[1][j], (x0, uold, abc)[j][perm[1]]) for j in range(3))       # the pinned copies hold the data
    for rep in range(3):
        for slot in range(4):
            if rep:
                s.wait(slot)
            s.solve_async(slot, *ins[slot], outs[slot])
    for slot in range(4):
        s.wait(slot)
    for i in range(4):
        for k in ("xPred", "uPred", "slack", "status", "iters", "resid"):
            assert np.array_equal(outs[i][k], ref[k][perm[i]]), (i, k)
    with pytest.raises(ValueError):
        s.solve_async(0, x0[::2], uold, abc, outs[0])          # non-contiguous view: refused, not silently copied
    with pytest.raises(nat.NativeError):
        s.solve_async(4, *ins[0], outs[0])                     # only four buffer sets
    s.close()
    if not have_alloc:
        pytest.skip("lmpc_host_alloc unavailable on this host (memory registration refused); slots verified with torch-pinned arrays")
    # the allocator's blocks: zero-filled, writable, freed with the last view
    a = nat.pinned_empty((3, 5), np.float64)
    assert a.shape == (3, 5) and not a.any() and a.flags.c_contiguous and a.flags.writeable
    a[...] = 1.5
    v = a[1]
    del a
    assert v.sum() == 7.5
    assert len(nat.page_nodes(v, 1)) == 1
