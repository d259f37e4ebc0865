import os, sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from racinglmpc_b200 import BatchedFTOCP, workloads, reference_params as rp
B = 4096
x0, uold, abc = workloads.ltv_mpc_batch(B, N=12)
solver = BatchedFTOCP(rp.mpc_params(12), batch=B)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h_x0, h_u, h_abc = pin(x0), pin(uold), pin(abc)
out = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in solver.alloc_outputs(False).items()}
for ch in (4, 8):
    os.environ["LMPC_B200_CHUNKS"] = str(ch)
    for _ in range(5): solver.solve(h_x0, h_u, h_abc, out=out)
    os.environ["LMPC_B200_TRACE"] = "1"
    t0 = time.perf_counter(); solver.solve(h_x0, h_u, h_abc, out=out); print("wall ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
    solver.solve(h_x0, h_u, h_abc, out=out)
    del os.environ["LMPC_B200_TRACE"]
