#!/usr/bin/env python
"""Small driver for ncu: two back-to-back solves of one workload (the second is the one to capture).

    ncu ... -k regex:ilqr_warp_kernel --launch-skip 3 --launch-count 1 python tools/profile_solve.py [cfg] [B]
Launch order: init, iterate (warm-up solve), init, iterate (capture this one).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ilqr_b200 import abi  # noqa: E402
from ilqr_b200.solver import BatchILQR, make_inputs  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["B"]
n_iters = int(sys.argv[3]) if len(sys.argv) > 3 else -1
x0, u0 = make_inputs(bench.SEED, B, cfg["T"], 4, 1)
kw = dict(u_min=[-cfg["limits"]], u_max=[cfg["limits"]]) if cfg["limits"] else {}
s = BatchILQR(abi.MODEL_ACROBOT, T=cfg["T"], B=B, dt=0.02, dtype=abi.F32 if cfg.get("dtype") == "f32" else abi.F64,
              cost_deriv=abi.COST_ANALYTIC if cfg["cost_deriv"] == "analytic" else abi.COST_FD, flags=cfg.get("flags", 0), **kw)
for _ in range(2):
    s.set_initial(x0, u0)
    if n_iters < 0:
        s.solve()
    else:
        s.iterate(n_iters)
    s.sync()
print("trips", int(s.get("iters").sum()), "accepted", int(s.get("n_accept").sum()), "rejected", int(s.get("n_reject").sum()))
