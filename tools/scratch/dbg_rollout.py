import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from racinglmpc_b200 import workloads, reference_params as rp
from racinglmpc_b200.controller import BatchedController
B, N = 4096, 12
g = np.load('/root/repo/tests/golden/reference_golden.npz'); xP, uP = g["pid_x"], g["pid_u"]
numSS_it, numSS_Points, _, _, Qts, par = rp.lmpc_params(N)
c = BatchedController(par, B, workloads.track_seg_table(), rp.TRACK_LENGTH, trToUse=4, numSS_Points=numSS_Points, numSS_it=numSS_it, QterminalSlack=Qts, Tmax=1280, ss_cap=7, model_cap=5)
for b in range(B):
    for _ in range(4): c.model_add_trajectory(b, xP, uP)
    for _ in range(4): c.add_trajectory(b, xP, uP)
c.set_state(xLin=np.tile(xP[1:N + 2], (B, 1, 1)), uLin=np.tile(uP[1:N + 1], (B, 1, 1)), zt=np.tile(np.array([0.0, 0, 0, 0, 10.0, 0]), (B, 1)), OldInput=np.zeros((B, 2)), timeStep=np.zeros(B, np.int32), has_pred=np.zeros(B, np.int32))
c.enable_rollout(Tcl=512)
x0 = np.tile(np.array([0.5, 0, 0, 0, 0, 0.0]), (B, 1)); c.rollout_set_state(x0, x0)
bad = []; dump = []; hist = np.zeros(64, int)
for k in range(300):
    xs = c.rollout_state()["x"]; uo = c.get_state()["OldInput"]
    c.rollout_step(seed=1234)
    r = c.step_results()
    hist += np.bincount(np.minimum(r["iters"], 63), minlength=64)
    for b in np.nonzero(r["status"] != 1)[0]:
        bad.append((k, int(b), int(r["status"][b]), int(r["iters"][b]), r["resid"][b].tolist()))
        dump.append(dict(k=k, b=int(b), x0=xs[b].copy(), uold=uo[b].copy(), abc=c.read_buffer("abc", b, (12, 54)), SS=c.read_buffer("SS_sel", b, (6, 48)), Qf=c.read_buffer("Qfun_sel", b, (48,))))
    for b in np.nonzero((r["status"] == 1) & (r["iters"] >= 22))[0]:
        dump.append(dict(k=k, b=int(b), x0=xs[b].copy(), uold=uo[b].copy(), abc=c.read_buffer("abc", b, (12, 54)), SS=c.read_buffer("SS_sel", b, (6, 48)), Qf=c.read_buffer("Qfun_sel", b, (48,)), slow=int(r["iters"][b])))
    done, n = c.rollout_done()
    if done.any(): c.rollout_finish_laps(done, n)
print('iters hist', {i: int(v) for i, v in enumerate(hist) if v})
import pickle; pickle.dump(dump, open('/root/repo/gpurun_out/unsolved_qps.pkl','wb'))
print(len(bad), 'unsolved'); 
for x in bad[:60]: print(x)
