"""CPU, dev container only: the oracle against the UNMODIFIED reference itself, live
(oracle/_ref/libref_oracle.so, built by `make -C oracle ref` from /root/reference).  Skipped where
that library is absent; the committed golden vectors (test_oracle_port.py) cover the same ground."""
import numpy as np
import pytest

from ilqr_b200 import abi

import oracleport as O
import refharness as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref_oracle.so not built")


def test_synth_matches_std_mt19937_64():
    import bench
    x0, u0 = bench.synth_inputs_cpu(4, 7, 999)
    ref = R.std_uniform(999, 4 * (4 + 7))
    got = np.concatenate([np.concatenate([x0[b], u0[b].ravel() / 0.5]) for b in range(4)])
    assert np.array_equal(got[11:], ref[11:])  # instance 0 is the canonical zero instance


def test_replica_loop_equals_native_generate_trajectory():
    """The probe's replica of the loop body (oracle/ref_harness.cpp) must land where the reference's own
    generate_trajectory() lands: that is what makes per-iteration golden traces trustworthy."""
    rng = np.random.default_rng(5)
    x0, u0 = rng.uniform(-1, 1, 4), 0.5 * rng.uniform(-1, 1, (120, 1))
    a, b = R.RefSolver(R.ACROBOT, 0.02), R.RefSolver(R.ACROBOT, 0.02)
    a.solve_native(x0, u0)
    b.init(x0, u0)
    b.iterate(1000)
    for f in ("xs", "us", "K", "k"):
        assert np.array_equal(a.get(f), b.get(f)), f
    assert a.cost == b.cost


@pytest.mark.parametrize("model,kw,T,dt", [(abi.MODEL_ACROBOT, {}, 150, 0.02),
                                           (abi.MODEL_ACROBOT, dict(u_min=[-1.5], u_max=[1.5]), 100, 0.02),
                                           (abi.MODEL_DOUBLE_INTEGRATOR, dict(goal=[1.0, 0.5, 0.0, 0.0]), 80, 0.03)])
def test_oracle_tracks_reference_on_fresh_instances(model, kw, T, dt):
    rng = np.random.default_rng(11)
    n, m = abi.MODEL_DIMS[model]
    for _ in range(3):
        x0, u0 = rng.uniform(-1, 1, n), 0.3 * rng.uniform(-1, 1, (T, m))
        r, o = R.RefSolver(model, dt, **kw), O.OracleSolver(model, dt, **kw)
        assert abs(r.init(x0, u0) - o.init(x0, u0)) <= 1e-12 * abs(r.cost)
        r.backward_once(1.0)
        o.backward_once(1.0)
        for f in ("fx", "fu", "cx", "cu"):
            assert np.allclose(r.get(f), o.get(f), rtol=1e-9, atol=1e-9), f
        for f in ("K", "k", "dV"):
            assert np.abs(r.get(f) - o.get(f)).max() <= 1e-7 * np.abs(r.get(f)).max(), f
        r.init(x0, u0)
        o.init(x0, u0)
        r.iterate(6)
        o.iterate(6)
        assert abs(r.cost - o.cost) <= 1e-7 * abs(r.cost)
        assert r.count("alpha_index") == o.count("alpha_index")
        assert np.abs(r.get("K") - o.get("K")).max() <= 1e-6 * np.abs(r.get("K")).max()


def test_replica_warm_start_and_resume_equal_native():
    """generate_trajectory(x_0) (src/ilqr_core.cpp:65-76) and a repeated generate_trajectory() (:78-102): the probe's
    replica entry points land exactly where the reference's own calls land, lambda / dlambda carried over."""
    rng = np.random.default_rng(8)
    x0, u0 = rng.uniform(-1, 1, 4), 0.5 * rng.uniform(-1, 1, (110, 1))
    x1 = x0 + 0.02
    a, b = R.RefSolver(R.ACROBOT, 0.02), R.RefSolver(R.ACROBOT, 0.02)
    a.solve_native(x0, u0)
    a.warm_native(x1)
    b.init(x0, u0)
    b.iterate(1000)
    b.warm_start(x1)
    b.iterate(1000)
    for f in ("xs", "us", "K", "k"):
        assert np.array_equal(a.get(f), b.get(f)), f
    assert a.cost == b.cost and a.scalar("lam") == b.scalar("lam") and a.scalar("dlam") == b.scalar("dlam")
    a.resume_native()
    b.resume()
    b.iterate(1000)
    for f in ("xs", "us", "K", "k"):
        assert np.array_equal(a.get(f), b.get(f)), f
    assert a.cost == b.cost


def test_oracle_warm_start_tracks_reference():
    rng = np.random.default_rng(9)
    for trial in range(3):
        x0, u0 = rng.uniform(-1, 1, 4), 0.5 * rng.uniform(-1, 1, (90, 1))
        x1 = x0 + rng.uniform(-0.03, 0.03, 4)
        r, o = R.RefSolver(R.ACROBOT, 0.02), O.OracleSolver(abi.MODEL_ACROBOT, 0.02)
        r.init(x0, u0), o.init(x0, u0)
        r.iterate(7), o.iterate(7)
        cr, co = r.warm_start(x1), o.warm_start(x1)
        assert abs(cr - co) <= 1e-9 * abs(cr)
        assert np.abs(r.get("xs") - o.get("xs")).max() <= 1e-9 * np.abs(r.get("xs")).max()
        r.iterate(4), o.iterate(4)
        assert abs(r.cost - o.cost) <= 1e-7 * abs(r.cost)
        assert r.scalar("lam") == o.scalar("lam")
        assert np.abs(r.get("K") - o.get("K")).max() <= 1e-6 * np.abs(r.get("K")).max()
