#!/usr/bin/env python
"""tools/diag_unsolved.py — solve the configs[1] / configs[2] workloads and print every instance that is not reported solved
(status, iterations, residuals) plus iteration statistics.  Run on the GPU box."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racinglmpc_b200 import BatchedFTOCP, workloads, reference_params as rp
B, N = 4096, 12
x0, uold, abc = workloads.ltv_mpc_batch(B, N=N)
s = BatchedFTOCP(rp.mpc_params(N), batch=B)
o = s.solve(x0, uold, abc)
bad = np.nonzero(o["status"] != 1)[0]
print("configs[1]: unsolved", len(bad), "late", s.late_accepts, "iters mean %.3f max %d" % (o["iters"].mean(), o["iters"].max()))
for b in bad[:10]:
    print("  inst", b, "status", o["status"][b], "iters", o["iters"][b], "resid", o["resid"][b])
print("  iteration histogram", np.bincount(o["iters"]))
