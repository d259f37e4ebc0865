"""GPU tier: presentation support.  Map.getGlobalPosition on the device against values produced by the real
reference (tests/golden/track_global.npz, tests/golden/make_track_global.py), and the rollout trace against what the host
sees step by step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from racinglmpc_b200 import export, reference_params as rp                  # noqa: E402
from racinglmpc_b200.controller import BatchedController                     # noqa: E402
import os                                                                    # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")


def test_global_position_matches_reference():
    _need_gpu()
    g = np.load(os.path.join(ROOT, "tests", "golden", "track_global.npz"))
    xy, ok = export.global_position(g["table"], float(g["track_length"]), g["s"], g["ey"])
    assert np.all(ok == 1)
    assert np.abs(xy - g["xy"]).max() < 1e-12, np.abs(xy - g["xy"]).max()        # sin/cos may differ in the last ulp
    _, ok2 = export.global_position(g["table"], float(g["track_length"]), np.array([-0.5, np.inf]), np.zeros(2))
    assert list(ok2) == [0, 0]                                                    # the reference raises there


def test_rollout_trace_records_what_the_host_sees(gold, track):
    _need_gpu()
    N, B = 12, 4
    numSS_it, numSS_Points, _, _, Qts, par = rp.lmpc_params(N)
    xP, uP = gold["pid_x"].copy(), gold["pid_u"].copy()
    c = BatchedController(par, B, track.seg_table(), track.TrackLength, trToUse=4, numSS_Points=numSS_Points, numSS_it=numSS_it,
                          QterminalSlack=Qts, Tmax=1536, ss_cap=7, model_cap=5)
    for b in range(B):
        for _ in range(4):
            c.model_add_trajectory(b, xP, uP)
        for _ in range(4):
            c.add_trajectory(b, xP, uP)
    c.set_state(xLin=np.tile(xP[1:N + 2], (B, 1, 1)), uLin=np.tile(uP[1:N + 1], (B, 1, 1)), zt=np.tile(np.array([0.0, 0, 0, 0, 10.0, 0]), (B, 1)),
                OldInput=np.zeros((B, 2)), timeStep=np.zeros(B, np.int32), has_pred=np.zeros(B, np.int32))
    c.enable_rollout(Tcl=512)
    c.enable_device_books()
    x0 = np.tile(np.array([0.5, 0, 0, 0, 0, 0.0]), (B, 1))
    c.rollout_set_state(x0, x0)
    tr = export.RolloutTrace(c, [1, 3], cap_steps=400)
    zs = np.random.default_rng(8).standard_normal((260, B, 3))
    seen = []
    for k in range(260):
        st = c.rollout_state()
        c.rollout_step(z=zs[k])
        xp = c.read_buffer("xPred", 0, (B, N + 1, 6)); ss = c.read_buffer("SS_sel", 0, (B, 6, numSS_Points)); up = c.read_buffer("uPred", 0, (B, N, 2))
        seen.append((st["x"].copy(), st["xglob"].copy(), up[:, 0].copy(), xp, ss))
        c.rollout_commit_laps_dev()
    for j, inst in enumerate((1, 3)):
        t = tr.get(j)
        assert t["x"].shape[0] == 260
        for k in (0, 1, 57, 259):
            assert np.array_equal(t["x"][k], seen[k][0][inst]) and np.array_equal(t["x_glob"][k], seen[k][1][inst])
            assert np.array_equal(t["u"][k], seen[k][2][inst]) and np.array_equal(t["xPred"][k], seen[k][3][inst])
            assert np.array_equal(t["SS_sel"][k], seen[k][4][inst])
        assert t["lap"][0] == 0 and t["lap"][-1] >= 1 and np.all(np.diff(t["lap"]) >= 0)
        v = tr.plot_view(j, first_lap_number=4)
        T0 = int(c.books()["lap_hist"][inst, 0])
        assert v.LapTime[0] == T0 and v.SS[0].shape == (T0, 6) and v.SS_glob[0].shape == (T0, 6) and v.it == 4 + len(v.SS)
        assert v.xStoredPredTraj[0][5].shape == (N + 1, 6) and v.SSStoredPredTraj[0][5].shape == (numSS_Points, 6)
        # global track position of the recorded states agrees with the vehicle's own global integration to plotting accuracy
        g = np.load(os.path.join(ROOT, "tests", "golden", "track_global.npz"))
        xy, ok = export.global_position(g["table"], float(g["track_length"]), v.SS[0][:, 4], v.SS[0][:, 5])
        assert np.all(ok == 1) and np.abs(xy - v.SS_glob[0][:, 4:6]).max() < 0.05
    c.close()
