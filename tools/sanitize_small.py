#!/usr/bin/env python
"""A small workload for compute-sanitizer (memcheck / racecheck / synccheck): both models, both cost-derivative modes,
every engine path a default build can take — the batch-lockstep phase kernels with each of their three trip heads
(8 lanes per trajectory, one thread per trajectory, one warp per trajectory) and the persistent 32-lane warp kernel —
a few loop trips each.  `--lanes16` adds the optional two-per-warp instantiation (ILQR_B200_LANES=16).

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqr_b200 import abi  # noqa: E402
from ilqr_b200.solver import BatchILQR, make_inputs  # noqa: E402

PATHS = [("phase/rows", dict(ILQR_B200_ROWS_MAX="1000000", ILQR_B200_WARP_PRE_MAX="0", ILQR_B200_HANDOVER="0")),
         ("phase/thread", dict(ILQR_B200_ROWS_MAX="0", ILQR_B200_WARP_PRE_MAX="0", ILQR_B200_HANDOVER="0")),
         ("phase/warp-head", dict(ILQR_B200_ROWS_MAX="0", ILQR_B200_WARP_PRE_MAX="1000000", ILQR_B200_HANDOVER="0")),
         ("phase/reroll", dict(ILQR_B200_ROWS_MAX="0", ILQR_B200_REROLL_MIN="0", ILQR_B200_HANDOVER="0")),
         ("phase/staged", dict(ILQR_B200_STAGE_MIN="0", ILQR_B200_STAGE_K="2", ILQR_B200_HANDOVER="0")),
         ("phase/ordered", dict(ILQR_B200_ORDERED_MIN="0", ILQR_B200_HANDOVER="0")),
         ("phase+handover", dict(ILQR_B200_HANDOVER="5", ILQR_B200_CHECK_EVERY="1")),
         ("warp32", dict(ILQR_B200_ENGINE="warp"))]
if "--lanes16" in sys.argv:
    PATHS.append(("warp16", dict(ILQR_B200_LANES="16")))
KEYS = sorted({k for _, e in PATHS for k in e})
for name, env in PATHS:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
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
            s.resume()
            s.iterate(4)
            c = s.get("cost")
            print("path", name, "model", model, "cost_deriv", cd, "cost[0] %.6g" % c[0], flush=True)
            s.close()
print("done")
