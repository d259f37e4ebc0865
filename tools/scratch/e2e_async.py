import os, sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from racinglmpc_b200 import BatchedFTOCP, workloads, reference_params as rp
B = 4096
x0, uold, abc = workloads.ltv_mpc_batch(B, N=12)
solver = BatchedFTOCP(rp.mpc_params(12), batch=B)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h_x0, h_u, h_abc = pin(x0), pin(uold), pin(abc)
outs = [{k: pin(v) for k, v in solver.alloc_outputs(False).items()} for _ in range(2)]
for i in range(4):
    solver.solve_async(i & 1, h_x0, h_u, h_abc, outs[i & 1]); solver.wait(i & 1)
for steps in (20, 20, 20, 20, 20, 100, 100, 400):
    t0 = time.perf_counter()
    for i in range(steps):
        slot = i & 1
        if i >= 2: solver.wait(slot)
        solver.solve_async(slot, h_x0, h_u, h_abc, outs[slot])
    solver.wait(0); solver.wait(1)
    dt = time.perf_counter() - t0
    print("steps %d: %.3f ms/step -> %.2f M/s" % (steps, dt / steps * 1e3, B * steps / dt / 1e6))
# synchronous for comparison
for steps in (20, 100):
    t0 = time.perf_counter()
    for i in range(steps): solver.solve(h_x0, h_u, h_abc, out=outs[0])
    dt = time.perf_counter() - t0
    print("sync steps %d: %.3f ms/step -> %.2f M/s" % (steps, dt / steps * 1e3, B * steps / dt / 1e6))
