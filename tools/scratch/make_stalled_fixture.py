#!/usr/bin/env python
"""gpurun_out/unsolved_qps.pkl (written on a B200 by tools/scratch/dbg_rollout.py at commit 14d7288, i.e. BEFORE the recentring
pass existed: every closed-loop LMPC QP that hit max_iter or needed >= 22 iterations, with its inputs read back from the device)
-> tests/golden/stalled_lmpc_qps.npz: six instances that hit max_iter and two that took >= 30 iterations."""
import pickle
import numpy as np
dump = pickle.load(open("gpurun_out/unsolved_qps.pkl", "rb"))
fails = [d for d in dump if "slow" not in d]
sel = fails[:6] + [d for d in dump if d.get("slow", 0) >= 30][:2]
out = {}
for i, d in enumerate(sel):
    for k in ("x0", "uold", "abc", "SS", "Qf"):
        out["q%d_%s" % (i, k)] = d[k]
    out["q%d_step" % i] = np.array([d["k"], d["b"]])
np.savez_compressed("tests/golden/stalled_lmpc_qps.npz", **out)
