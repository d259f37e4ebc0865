import os, sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from racinglmpc_b200 import BatchedFTOCP, workloads, reference_params as rp
B = 4096
x0, uold, abc = workloads.ltv_mpc_batch(B, N=12)
solver = BatchedFTOCP(rp.mpc_params(12), batch=B)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h_x0, h_u, h_abc = pin(x0), pin(uold), pin(abc)
out = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in solver.alloc_outputs(False).items()}
for ch in (1, 2, 4, 8, 4, 1, 8):
    os.environ["LMPC_B200_CHUNKS"] = str(ch)
    for _ in range(3): solver.solve(h_x0, h_u, h_abc, out=out)
    ts = []
    for rep in range(5):
        t0 = time.perf_counter()
        for _ in range(20): solver.solve(h_x0, h_u, h_abc, out=out)
        ts.append((time.perf_counter() - t0) / 20)
    print("chunks", ch, "ms/solve min %.3f med %.3f max %.3f" % (min(ts) * 1e3, sorted(ts)[2] * 1e3, max(ts) * 1e3), "-> %.2f M/s" % (B / sorted(ts)[2] / 1e6))
