#!/usr/bin/env python
"""Experiment: solve time of one batch vs the trip count of the two-per-warp first launch (ILQR_B200_HYBRID_TRIPS)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ilqr_b200 import abi
from ilqr_b200.solver import BatchILQR, make_inputs
Bs = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4096").split(",")]
Ks = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,20,30,40,50").split(",")]
x0a, u0a = make_inputs(bench.SEED, max(Bs), 200, 4, 1)
for B in Bs:
    for K in Ks:
        os.environ["ILQR_B200_HYBRID_TRIPS"] = str(K)
        s = BatchILQR(abi.MODEL_ACROBOT, T=200, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
        best = 1e30
        for _ in range(3):
            s.set_initial(x0a[:B], u0a[:B]); s.sync()
            t0 = time.perf_counter(); s.solve(); s.sync(); best = min(best, time.perf_counter() - t0)
        it = s.get("iters"); c = s.get("cost")
        print("B=%d first-launch trips %3d: solve %.3f ms  %.3f Mtrips/s  (trips %d, cost checksum %.10e)" % (B, K, best * 1e3, it.sum() / best / 1e6, it.sum(), c.sum()), flush=True)
        s.close()
