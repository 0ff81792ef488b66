#!/usr/bin/env python
"""f32 solves (BASELINE configs[2] arithmetic) against the f64 oracle on the same instances: the numbers behind the
gates of tests/test_gpu_parity.py::test_f32_parity_vs_f64_oracle.   python tools/exp_f32_parity.py [B]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("ILQR_B200_HANDOVER", "0")
from ilqr_b200 import abi  # noqa: E402
from ilqr_b200.solver import BatchILQR, make_inputs  # noqa: E402
import oracleport as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64


def err(a, b, atol=1e-6):
    a2, b2 = np.asarray(a, float).reshape(len(a), -1), np.asarray(b, float).reshape(len(b), -1)
    return np.maximum(np.abs(a2 - b2).max(1) - atol, 0) / np.maximum(np.abs(b2).max(1), 1e-300)


for T in (200, 500):
    x0, u0 = make_inputs(12345, B, T, 4, 1)
    x0, u0 = x0.astype(np.float32).astype(np.float64), u0.astype(np.float32).astype(np.float64)
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, dtype=abi.F32, cost_deriv=abi.COST_ANALYTIC)
    c0 = s.init_traj(x0, u0)
    os_ = []
    for b in range(B):
        o = O.OracleSolver(abi.MODEL_ACROBOT, 0.02, cost_deriv=abi.COST_ANALYTIC)
        o.init(x0[b], u0[b])
        os_.append(o)
    e0 = err(c0, np.array([o.cost for o in os_]), 0)
    print("T", T, "init cost rel: med %.2e max %.2e frac<2e-4 %.3f" % (np.median(e0), e0.max(), (e0 < 2e-4).mean()))
    done = 0
    for n in (1, 5, 20):
        s.iterate(n - done)
        for o in os_:
            o.iterate(n - done)
        done = n
        same = s.get("alpha_index") == np.array([o.count("alpha_index") for o in os_])
        eK = err(s.get("K"), np.stack([o.get("K") for o in os_]))
        ek = err(s.get("k"), np.stack([o.get("k") for o in os_]))
        ec = err(s.get("cost"), np.array([o.cost for o in os_]), 0)
        print("  trip %2d same_alpha %.3f | same-branch instances: K med %.2e max %.2e  k med %.2e max %.2e  cost med %.2e max %.2e"
              " | all: K med %.2e frac<5e-2 %.3f cost frac<1e-2 %.3f" % (
                  n, same.mean(), np.median(eK[same]), eK[same].max(), np.median(ek[same]), ek[same].max(), np.median(ec[same]),
                  ec[same].max(), np.median(eK), (eK < 5e-2).mean(), (ec < 1e-2).mean()))
    s.solve()
    for o in os_:
        o.iterate(200)
    ec = err(s.get("cost"), np.array([o.cost for o in os_]), 0)
    st = s.get("status")
    so = np.array([o.count("status") for o in os_])
    print("  termination: cost frac<1e-3 %.3f frac<1e-2 %.3f | MAXITER f32 %.3f f64 %.3f | mean trips f32 %.1f f64 %.1f" % (
        (ec < 1e-3).mean(), (ec < 1e-2).mean(), (st == 4).mean(), (so == 4).mean(), s.get("iters").mean(),
        np.mean([o.count("loop_trips") for o in os_])))
