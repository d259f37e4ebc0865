"""CPU tier: the host-side lap bookkeeping of racinglmpc_b200.controller.BatchedController — which stored laps are "the numSS_it
fastest" (PC.py:395,402), which are usedIt (PM.py:31,35-46), what LMPC.addPoint extends (lap it-1, PC.py:466-476) and how laps
received through the pooled-safe-set exchange are numbered — against a recording stub in place of liblmpc_b200 (no GPU, no
compute: every native call returns 0 and is logged).  The device side of the same calls is covered by the -m gpu tests."""
import ctypes as C
import numpy as np
import pytest

from racinglmpc_b200 import _native as nat, reference_params as rp
from racinglmpc_b200.controller import BatchedController


class _Stub:
    """Stands in for the ctypes library object: any lmpc_* function returns 0 and is logged by name."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if not name.startswith("lmpc_"):
            raise AttributeError(name)

        def f(*args):
            self.calls.append(name)
            return 0
        return f


@pytest.fixture
def ctrl(monkeypatch):
    stub = _Stub()
    monkeypatch.setattr(nat, "lib", lambda: stub)
    numSS_it, numSS_Points, _, _, Qts, par = rp.lmpc_params(12)
    seg = np.array([[0.0, 19.3, 0.0]])
    c = BatchedController(par, 3, seg, 19.3, trToUse=4, numSS_Points=numSS_Points, numSS_it=numSS_it, QterminalSlack=Qts,
                          Tmax=512, ss_cap=7, model_cap=5)
    c._h = C.c_void_p(1)
    yield c, stub
    c._h = None


def _lap(T):
    return np.zeros((T, 6)), np.zeros((T, 2))


def test_selection_is_the_stable_argsort_of_lap_times_and_addpoint_target_is_the_latest_lap(ctrl):
    c, stub = ctrl
    for b in range(3):
        for T in (300, 250, 250, 400):                      # lap numbers 0..3; two equal lap times
            c.model_add_trajectory(b, *_lap(T))
            c.add_trajectory(b, *_lap(T))
    c._flush()
    assert c.it == [4, 4, 4]
    for b in range(3):
        slots = [c.ss_book[b].slot_of[j] for j in (1, 2, 0, 3)]          # argsort([300,250,250,400], stable) = 1,2,0,3
        assert list(c._sel[b]) == slots
        assert list(c._isp[b]) == [0, 0, 0, 1]                            # only lap it-1 = 3 is "the previous lap" (PC.py:506)
        assert c._prev[b] == c.ss_book[b].slot_of[3]
        # usedIt = the trToUse shortest laps in PredictiveModel's insertion order (equal lengths: later lap after earlier one)
        assert [ln for (_, ln) in c.model_laps[b][:4]] == [1, 2, 0, 3]
    assert "lmpc_ss_set_selection" in stub.calls and "lmpc_model_set_used" in stub.calls


def test_incremental_push_rebuilds_only_the_changed_instances(ctrl):
    c, stub = ctrl
    for b in range(3):
        for T in (300, 280, 260, 240):
            c.model_add_trajectory(b, *_lap(T)); c.add_trajectory(b, *_lap(T))
    c._flush()
    before = c._sel.copy()
    done = np.array([0, 1, 0], np.int32); n = np.array([0, 200, 0], np.int32)
    fin = c.rollout_finish_laps(done, n)                                  # instance 1 finishes a 200-step lap
    assert list(fin) == [1] and c._sel_rows == {1} and c._used_rows == {1}
    c._flush()
    assert c._sel_rows == set() and c._used_rows == set()
    assert np.array_equal(c._sel[[0, 2]], before[[0, 2]])
    assert c._sel[1][0] == c.ss_book[1].slot_of[4] and c._isp[1][0] == 1  # the new lap is the fastest and is lap it-1
    assert c.it == [4, 5, 4] and c.LapTime[1] == [300, 280, 260, 240, 200]
    assert c.model_laps[1][0] == (200, 4)


def test_foreign_lap_is_numbered_before_the_own_latest_lap(ctrl):
    c, stub = ctrl
    for b in range(3):
        for T in (1000, 1000, 1000, 1000, 230):                           # four seed laps + one own LMPC lap (lap 4 = it-1)
            c.model_add_trajectory(b, *_lap(T)); c.add_trajectory(b, *_lap(T))
    c._flush()
    own_slot = [c.ss_book[b].slot_of[4] for b in range(3)]
    rows = np.zeros((6, 288, 9)); lens = np.zeros(6, np.int32)            # stand-ins for the gathered device tensors
    took = c.import_laps(np.array([-1, 3, 5]), np.array([0, 210, 5000]), 288, rows, lens)
    assert list(took) == [1]                                              # 0: nothing offered, 2: 5000 would never be selected
    assert c.it == [5, 6, 5]
    assert c.LapTime[1] == [1000, 1000, 1000, 1000, 210, 230]             # foreign lap took number 4, the own lap moved to 5
    assert c.ss_book[1].slot_of[5] == own_slot[1] and c.own_lap_number(1, 4) == 5
    c._flush()
    assert c._prev[1] == own_slot[1]                                       # addPoint keeps extending the OWN latest lap
    assert c._sel[1][0] == c.ss_book[1].slot_of[4] and c._isp[1][0] == 0   # fastest = the foreign lap, an "older" lap
    assert c._sel[1][1] == own_slot[1] and c._isp[1][1] == 1
    assert (210, 5) in c.model_laps[1][:1] or c.model_laps[1][0][0] == 210  # and it leads usedIt of the regression model
    # the next own lap lands after both
    c.rollout_finish_laps(np.array([0, 1, 0], np.int32), np.array([0, 205, 0], np.int32))
    assert c.LapTime[1] == [1000, 1000, 1000, 1000, 210, 230, 205] and c.own_laps[1][-2:] == [5, 6]


def test_pool_eviction_keeps_the_fastest_laps_and_the_latest(ctrl):
    c, stub = ctrl
    times = [1000, 1000, 1000, 1000, 240, 230, 220, 210, 200]             # ss_cap = 7 < 9 laps
    for T in times:
        c.add_trajectory(0, *_lap(T))
    kept = sorted(c.ss_book[0].slot_of)
    assert len(kept) <= 7 and all(l in kept for l in (5, 6, 7, 8))         # the four fastest + the latest are stored
    assert 8 in kept and len(set(c.ss_book[0].slot_of.values())) == len(kept)
