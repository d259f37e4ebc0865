"""End-to-end leg of bench.py in one process, chunk policies interleaved (LMPC_B200_CHUNKS is read at every enqueue):
which policy the library picks per batch, and solves/s for auto / 1 / 2 / 4 chunks with 3 buffer sets in flight."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from racinglmpc_b200 import BatchedFTOCP, workloads, reference_params as rp, _native as nat

B, N, NS = 4096, 12, 3
x0, uold, abc = workloads.ltv_mpc_batch(B, N=N)
s = BatchedFTOCP(rp.mpc_params(N), batch=B)
pin = lambda a: nat.pinned_like(np.ascontiguousarray(a))
h = [pin(x0), pin(uold), pin(abc)]
outs = [{k: pin(v) for k, v in s.alloc_outputs(False).items()} for _ in range(NS)]
L = nat.lib()


def run(steps, log=False):
    seen = []
    for sl in range(NS):
        s.wait(sl)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        sl = i % NS
        if i >= NS:
            s.wait(sl)
        s.solve_async(sl, *h, outs[sl])
        if log:
            seen.append(int(L.lmpc_host_chunks(s._h, sl)))
    for sl in range(NS):
        s.wait(sl)
    dt = time.perf_counter() - t0
    return B * steps / dt, seen


for mode in ("auto", "1"):
    if mode == "auto":
        os.environ.pop("LMPC_B200_CHUNKS", None)
    else:
        os.environ["LMPC_B200_CHUNKS"] = mode
    run(30)
res = {}
for rep in range(4):
    for mode in ("auto", "1", "2", "4"):
        if mode == "auto":
            os.environ.pop("LMPC_B200_CHUNKS", None)
        else:
            os.environ["LMPC_B200_CHUNKS"] = mode
        v20, seen = run(20, log=(rep == 0))
        v200, _ = run(200)
        res.setdefault(mode, []).append((round(v20), round(v200)))
        if rep == 0:
            print(mode, "chunks per batch:", seen)
for k, v in res.items():
    print(k, v)
