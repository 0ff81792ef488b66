#!/usr/bin/env python
"""A small workload for compute-sanitizer (memcheck / racecheck / synccheck): both models, both cost-derivative modes,
both lane decompositions, a few loop trips each.

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqr_b200 import abi  # noqa: E402
from ilqr_b200.solver import BatchILQR, make_inputs  # noqa: E402

for lanes in ("32", "16"):
    os.environ["ILQR_B200_LANES"] = lanes
    for model, T, m in ((abi.MODEL_ACROBOT, 37, 1), (abi.MODEL_DOUBLE_INTEGRATOR, 19, 2)):
        for cd in (abi.COST_ANALYTIC, abi.COST_FD):
            B = 9
            x0, u0 = make_inputs(7, B, T, 4, m)
            kw = dict(goal=[1.0, 0.5, 0.0, 0.0]) if model == abi.MODEL_DOUBLE_INTEGRATOR else {}
            s = BatchILQR(model, T=T, B=B, dt=0.02, cost_deriv=cd, **kw)
            s.set_initial(x0, u0)
            s.iterate(3)
            s.warm_start(x0 + 0.01)
            s.iterate(2)
            c = s.get("cost")
            print("lanes", lanes, "model", model, "cost_deriv", cd, "cost[0] %.6g" % c[0], flush=True)
            s.close()
print("done")
