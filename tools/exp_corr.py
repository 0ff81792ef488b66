import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import bench
from ilqr_b200 import abi
from ilqr_b200.solver import BatchILQR, make_inputs
B, T = 4096, 200
x0, u0 = make_inputs(bench.SEED, B, T, 4, 1)
s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
c0 = s.init_traj(x0, u0).copy()
s.iterate(1); c1 = s.get("cost").copy(); g1 = s.get("gnorm").copy()
s.iterate(4); c5 = s.get("cost").copy()
s.solve(); it = s.get("iters"); cf = s.get("cost"); st = s.get("status")
print("trips: mean %.1f  pct>=60: %.3f  pct>=80: %.3f pct==100(+): %.3f" % (it.mean(), (it >= 60).mean(), (it >= 80).mean(), (it >= 100).mean()))
print("hist", np.histogram(it, bins=[0, 20, 30, 40, 50, 60, 70, 80, 90, 100, 200])[0])
for name, v in [("init cost", c0), ("cost after 1", c1), ("cost after 5", c5), ("|x0|", np.abs(x0).sum(1)), ("x0[0]", x0[:, 0]), ("x0[1]", x0[:, 1]), ("x0[2]", x0[:, 2]), ("x0[3]", x0[:, 3]), ("final cost", cf), ("gnorm1", g1), ("dcost01", c0 - c1), ("c5/c0", c5 / c0)]:
    r = np.corrcoef(v, it)[0, 1]
    # rank correlation
    rr = np.corrcoef(np.argsort(np.argsort(v)), np.argsort(np.argsort(it)))[0, 1]
    print("%-14s pearson %+.3f  spearman %+.3f" % (name, r, rr))
print("status counts", np.bincount(st))
