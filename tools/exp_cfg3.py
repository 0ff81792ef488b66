#!/usr/bin/env python
"""BASELINE configs[2]: acrobot batch=65536 T=500 fp32, finite-difference fx/fu, closed-form cost derivatives, one GPU.
Solve to termination through the C ABI (host buffers in, costs out); prints time, trips and iterations/s."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ilqr_b200 import abi
from ilqr_b200.solver import BatchILQR, make_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
x0, u0 = make_inputs(bench.SEED, B, T, 4, 1)
s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, dtype=abi.F32, cost_deriv=abi.COST_ANALYTIC)
best = 1e30
for rep in range(2):
    s.set_initial(x0.astype(np.float32), u0.astype(np.float32)); s.sync()
    t0 = time.perf_counter(); s.solve(); s.sync(); best = min(best, time.perf_counter() - t0)
it = s.get("iters"); c = s.get("cost"); st = s.get("status")
print("configs[2] f32 B=%d T=%d: solve %.1f ms, trips %d (mean %.1f), %.3f M iterations/s; finite costs %.4f; status counts %s" % (
    B, T, best * 1e3, it.sum(), it.mean(), it.sum() / best / 1e6, np.isfinite(c).mean(), np.bincount(st).tolist()))
