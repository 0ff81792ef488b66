#!/usr/bin/env python
"""Experiment driver: solve time vs batch size and lane-group width (ILQR_B200_LANES is read at ilqr_create).

    python tools/exp_scaling.py [cfg] [B,B,...] [lanes,lanes,...] [n_iters]
Prints one line per (B, lanes): best-of-3 solve time (CUDA-stream sync wall clock), trips, trips/s.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ilqr_b200 import abi  # noqa: E402
from ilqr_b200.solver import BatchILQR, make_inputs  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
Bs = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,592,4096").split(",")]
lanes = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "32,16").split(",")]
n_iters = int(sys.argv[4]) if len(sys.argv) > 4 else -1
Bmax = max(Bs)
x0a, u0a = make_inputs(bench.SEED, Bmax, cfg["T"], 4, 1)
kw = dict(u_min=[-cfg["limits"]], u_max=[cfg["limits"]]) if cfg["limits"] else {}
for B in Bs:
    for ln in lanes:
        os.environ["ILQR_B200_LANES"] = str(ln)
        s = BatchILQR(abi.MODEL_ACROBOT, T=cfg["T"], B=B, dt=0.02,
                      cost_deriv=abi.COST_ANALYTIC if cfg["cost_deriv"] == "analytic" else abi.COST_FD, **kw)
        best = 1e30
        for _ in range(3):
            s.set_initial(x0a[:B], u0a[:B])
            s.sync()
            t0 = time.perf_counter()
            if n_iters < 0:
                s.solve()
            else:
                s.iterate(n_iters)
            s.sync()
            best = min(best, time.perf_counter() - t0)
        it = s.get("iters")
        print("B=%d lanes=%d  solve %.3f ms  trips %d (mean %.1f max %d)  %.3f Mtrips/s  %.1f us/trip-of-longest  cost checksum %.12e" % (
            B, ln, best * 1e3, it.sum(), it.mean(), it.max(), it.sum() / best / 1e6, best * 1e6 / it.max(), s.get("cost").sum()), flush=True)
        s.close()
