import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from racinglmpc_b200 import workloads, reference_params as rp
from racinglmpc_b200.controller import BatchedController
B, N = 4096, 12
numSS_it, numSS_Points, _, _, Qts, par = rp.lmpc_params(N)
c = BatchedController(par, B, workloads.track_seg_table(), rp.TRACK_LENGTH, trToUse=4, numSS_Points=numSS_Points, numSS_it=numSS_it, QterminalSlack=Qts, Tmax=1280, ss_cap=7, model_cap=5)
c.enable_rollout(Tcl=1024)
x0 = np.tile(np.array([0.5, 0, 0, 0, 0, 0.0]), (B, 1)); c.rollout_set_state(x0, x0)
for _ in range(1000): c.rollout_pid_step(0.8, seed=4321)
_, n = c.rollout_done(); c.rollout_seed_from_record(n, copies=4); c.rollout_set_state(x0, x0)
first = {}; hist = np.zeros(64, int); xs_hist = {}
prev_x = c.rollout_state()["x"]
for k in range(80):
    stt = c.get_state()
    c.rollout_step(seed=1234)
    r = c.step_results(); st = c.rollout_state()
    hist += np.bincount(np.minimum(r["iters"], 63), minlength=64)
    for b in np.nonzero(r["flags"])[0]:
        if int(b) not in first:
            first[int(b)] = (k, int(r["flags"][b]), int(r["status"][b]), int(r["iters"][b]), prev_x[b].round(4).tolist())
            print('inst', int(b), 'step', k, 'flags', int(r['flags'][b]), 'x', prev_x[b].round(4).tolist()); print(' xLin s', stt['xLin'][b][:, 4].round(4).tolist()); print(' xLin ey', stt['xLin'][b][:, 5].round(3).tolist()); print(' xLin vx', stt['xLin'][b][:, 0].round(3).tolist()); print(' uLin', stt['uLin'][b].round(3).tolist())
    prev_x = st["x"]
    done, nn = c.rollout_done()
    if done.any(): c.rollout_finish_laps(done, nn)
print("iters hist", {i: int(v) for i, v in enumerate(hist) if v})
for b, v in sorted(first.items(), key=lambda kv: kv[1][0]): print(b, v)
pid = c.get_lap(next(iter(first)) if first else 0, 0)[0]
print("a flagged instance's PID lap: vx min/max", pid[:, 0].min(), pid[:, 0].max(), "ey min/max", pid[:, 5].min(), pid[:, 5].max())
