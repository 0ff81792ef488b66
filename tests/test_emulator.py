"""CPU: the KERNEL SOURCE (ilqr_b200/csrc/ilqr_core.cuh + boxqp.cuh + models.cuh) compiled with g++
and run lane by lane (tests/emu), against the oracle.

  * with the platform libm for sin/cos (-DILQR_TRIG_LIBM) the kernel source must reproduce the
    oracle BIT FOR BIT — every array, every scalar, every counter, at every loop trip — for both
    models, both cost-derivative modes and the control-limited case: the lane decomposition and
    the arithmetic order of the warp code are exactly the reference's;
  * with the deterministic sincos the CUDA build uses (trig.cuh) the double integrator stays
    bit-exact and the acrobot moves by the sin/cos noise floor, which this file measures.
The GPU suite then checks the real kernels bit for bit against this same CPU build.
"""
import numpy as np
import pytest

from ilqr_b200 import abi
from ilqr_b200.solver import ILQRError  # noqa: F401  (import check only; no CUDA call)

import emuport as E
import oracleport as O

ARR = ("xs", "us", "K", "k")
SCAL = ("lam", "dlam", "gnorm", "dcost", "expected", "alpha", "new_cost")
CNT = ("iter", "loop_trips", "status", "alpha_index", "accepts", "rejects", "rollouts", "backwards", "derivs", "diverge")


def lockstep(model, x0, u0, dt, libm, max_trips=101, lanes=32, **kw):
    o, e = O.OracleSolver(model, dt, **kw), E.EmuSolver(model, dt, libm=libm, lanes=lanes, **kw)
    assert o.init(x0, u0) == e.init(x0, u0)
    assert np.array_equal(o.get("xs"), e.get("xs"))
    assert o.backward_once(1.0) == e.backward_once(1.0)
    for f in ("K", "k", "dV"):
        assert np.array_equal(o.get(f), e.get(f)), f
    assert np.array_equal(o.get("Vx")[0], e.get("Vx0")) and np.array_equal(o.get("Vxx")[0], e.get("Vxx0"))
    assert o.scalar("gnorm") == e.scalar("gnorm")
    assert o.rollout_once(0.5012) == e.rollout_once(0.5012)
    o.init(x0, u0)
    e.init(x0, u0)
    for trip in range(max_trips):
        o.iterate(1)
        e.iterate(1)
        for f in ARR:
            assert np.array_equal(o.get(f), e.get(f)), (trip, f)
        for f in SCAL:
            assert o.scalar(f) == e.scalar(f), (trip, f)
        for f in CNT:
            assert o.count(f) == e.count(f), (trip, f)
        assert o.cost == e.cost
        if o.count("status") != 0:
            break
    return o.count("loop_trips")


@pytest.mark.parametrize("case,cd,kw", [
    ("acrobot_T200_b0", abi.COST_FD, {}), ("acrobot_T200_b1", abi.COST_FD, {}), ("acrobot_T200_b2", abi.COST_ANALYTIC, {}),
    ("acrobot_T200_b3", abi.COST_ANALYTIC, {}), ("acrobot_lim15_T200_b1", abi.COST_FD, dict(u_min=[-1.5], u_max=[1.5])),
    ("acrobot_lim15_T200_b2", abi.COST_ANALYTIC, dict(u_min=[-1.5], u_max=[1.5])), ("acrobot_cli_T499", abi.COST_FD, {}),
])
def test_acrobot_kernel_source_equals_oracle_bit_for_bit(golden_solver, case, cd, kw):
    g = golden_solver
    trips = lockstep(abi.MODEL_ACROBOT, g[case + "/x0"], g[case + "/u0"], float(g[case + "/dt"]), True, cost_deriv=cd, **kw)
    assert trips >= 10


@pytest.mark.parametrize("libm", [True, False])
@pytest.mark.parametrize("case,cd", [("integrator_cli_T99", abi.COST_FD), ("integrator_rand_T60_b0", abi.COST_FD),
                                     ("integrator_rand_T60_b1", abi.COST_ANALYTIC), ("integrator_rand_T60_b2", abi.COST_FD)])
def test_double_integrator_kernel_source_equals_oracle_bit_for_bit(golden_solver, case, cd, libm):
    g = golden_solver
    lockstep(abi.MODEL_DOUBLE_INTEGRATOR, g[case + "/x0"], g[case + "/u0"], float(g[case + "/dt"]), libm,
             goal=list(g[case + "/goal"]), cost_deriv=cd)


@pytest.mark.parametrize("case,model,cd,kw", [
    ("acrobot_T200_b1", abi.MODEL_ACROBOT, abi.COST_FD, {}), ("acrobot_T200_b4", abi.MODEL_ACROBOT, abi.COST_ANALYTIC, {}),
    ("acrobot_lim15_T200_b0", abi.MODEL_ACROBOT, abi.COST_FD, dict(u_min=[-1.5], u_max=[1.5])),
    ("integrator_rand_T60_b1", abi.MODEL_DOUBLE_INTEGRATOR, abi.COST_FD, None)])
def test_sixteen_lane_decomposition_equals_oracle_bit_for_bit(golden_solver, case, model, cd, kw):
    """the lane decomposition used when a warp carries two trajectories (16 lanes each)"""
    g = golden_solver
    if kw is None:
        kw = dict(goal=list(g[case + "/goal"]))
    lockstep(model, g[case + "/x0"], g[case + "/u0"], float(g[case + "/dt"]), True, lanes=16, cost_deriv=cd, **kw)


@pytest.mark.parametrize("case,model,cd,kw", [
    ("acrobot_T200_b0", abi.MODEL_ACROBOT, abi.COST_FD, {}), ("acrobot_T200_b1", abi.MODEL_ACROBOT, abi.COST_ANALYTIC, {}),
    ("acrobot_T200_b2", abi.MODEL_ACROBOT, abi.COST_ANALYTIC, {}), ("acrobot_T200_b5", abi.MODEL_ACROBOT, abi.COST_FD, {}),
    ("acrobot_lim15_T200_b0", abi.MODEL_ACROBOT, abi.COST_FD, dict(u_min=[-1.5], u_max=[1.5])),
    ("acrobot_lim15_T200_b2", abi.MODEL_ACROBOT, abi.COST_ANALYTIC, dict(u_min=[-1.5], u_max=[1.5])),
    ("acrobot_cli_T499", abi.MODEL_ACROBOT, abi.COST_FD, {}),
    ("integrator_cli_T99", abi.MODEL_DOUBLE_INTEGRATOR, abi.COST_FD, None),
    ("integrator_rand_T60_b0", abi.MODEL_DOUBLE_INTEGRATOR, abi.COST_ANALYTIC, None),
    ("integrator_rand_T60_b2", abi.MODEL_DOUBLE_INTEGRATOR, abi.COST_FD, None)])
def test_phase_engine_source_equals_oracle_bit_for_bit(golden_solver, case, model, cd, kw):
    """the batch-lockstep engine (ilqr_b200/csrc/ilqr_phases.cuh: one thread per sweep task / trajectory / candidate):
    the functions its kernels call, run task by task on the CPU, reproduce the oracle bit for bit at every trip.  The
    emulation cycles through the three line-search modes of the kernels from trip to trip (tests/emu/ilqr_emu.cpp):
    every candidate kept, cost-only rollouts + re-roll of the accepted one, and the staged search (first four kept
    compactly, the rest cost-only where none of those passed)."""
    g = golden_solver
    if kw is None:
        kw = dict(goal=list(g[case + "/goal"]))
    lockstep(model, g[case + "/x0"], g[case + "/u0"], float(g[case + "/dt"]), True, lanes=1, cost_deriv=cd, **kw)


def test_sixteen_lane_decomposition_batch_of_64():
    """the two-per-warp decomposition on a batch: 64 acrobot instances in the reference's FD mode and 64 double-integrator
    instances (m = 2) against the oracle bit for bit over their first trips, and in float against the 32-lane
    decomposition of the same source"""
    import bench
    x0, u0 = bench.synth_inputs_cpu(64, 120, 4321)
    for b in range(64):
        o = O.OracleSolver(abi.MODEL_ACROBOT, 0.02)
        e = E.EmuSolver(abi.MODEL_ACROBOT, 0.02, libm=True, lanes=16)
        assert o.init(x0[b], u0[b]) == e.init(x0[b], u0[b])
        o.iterate(6), e.iterate(6)
        for f in ARR:
            assert np.array_equal(o.get(f), e.get(f)), (b, f)
        assert o.cost == e.cost and o.scalar("lam") == e.scalar("lam") and o.count("alpha_index") == e.count("alpha_index")
    rng = np.random.default_rng(6)
    for b in range(64):
        xd, ud = rng.uniform(-1, 1, 4), 0.5 * rng.uniform(-1, 1, (40, 2))
        kw = dict(goal=[1.0, 1.0, 0.0, 0.0], cost_deriv=abi.COST_FD if b % 2 else abi.COST_ANALYTIC)
        o = O.OracleSolver(abi.MODEL_DOUBLE_INTEGRATOR, 0.05, **kw)
        e = E.EmuSolver(abi.MODEL_DOUBLE_INTEGRATOR, 0.05, libm=True, lanes=16, **kw)
        assert o.init(xd, ud) == e.init(xd, ud)
        o.iterate(50), e.iterate(50)
        for f in ARR:
            assert np.array_equal(o.get(f), e.get(f)), (b, f)
        assert o.count("status") == e.count("status") and o.count("loop_trips") == e.count("loop_trips")
    for b in range(16):
        a = E.EmuSolver(abi.MODEL_ACROBOT, 0.02, cost_deriv=abi.COST_ANALYTIC, dtype=abi.F32, lanes=16)
        c = E.EmuSolver(abi.MODEL_ACROBOT, 0.02, cost_deriv=abi.COST_ANALYTIC, dtype=abi.F32, lanes=32)
        d = E.EmuSolver(abi.MODEL_ACROBOT, 0.02, cost_deriv=abi.COST_ANALYTIC, dtype=abi.F32, lanes=1)
        for s_ in (a, c, d):
            s_.init(x0[b], u0[b])
            s_.iterate(8)
        for f in ARR:
            assert np.array_equal(a.get(f), c.get(f)) and np.array_equal(a.get(f), d.get(f)), (b, f)


@pytest.mark.parametrize("lanes", [32, 1])
@pytest.mark.parametrize("case,model,flags,kw", [
    ("acrobot_lim15_T200_b0", abi.MODEL_ACROBOT, abi.FLAG_CLAMP_ROLLOUT, dict(u_min=[-1.5], u_max=[1.5])),
    ("acrobot_T200_b1", abi.MODEL_ACROBOT, abi.FLAG_ANALYTIC_DYN, {}),
    ("acrobot_lim15_T200_b2", abi.MODEL_ACROBOT, abi.FLAG_ANALYTIC_DYN | abi.FLAG_CLAMP_ROLLOUT, dict(u_min=[-1.5], u_max=[1.5])),
    ("integrator_rand_T60_b1", abi.MODEL_DOUBLE_INTEGRATOR, abi.FLAG_ANALYTIC_DYN | abi.FLAG_CLAMP_ROLLOUT, None)])
def test_opt_in_modes_kernel_source_equals_oracle(golden_solver, case, model, flags, kw, lanes):
    """SURVEY §8 (f4), both OFF by default: rollouts that clamp the applied control (the reference's commented-out
    "right way", src/ilqr_core.cpp:322-329) and closed-form dynamics Jacobians (notes.md:15,45).  The kernel source (warp
    phases and phase-engine tasks) against the oracle's implementation of the same two modes, bit for bit."""
    g = golden_solver
    if kw is None:
        kw = dict(goal=list(g[case + "/goal"]))
    lockstep(model, g[case + "/x0"], g[case + "/u0"], float(g[case + "/dt"]), True, lanes=lanes, flags=flags, **kw)


def test_opt_in_modes_do_what_they_say(golden_solver):
    g = golden_solver
    case = "acrobot_lim15_T200_b0"
    kw = dict(u_min=[-1.5], u_max=[1.5])
    a = O.OracleSolver(abi.MODEL_ACROBOT, 0.02, **kw)
    b = O.OracleSolver(abi.MODEL_ACROBOT, 0.02, flags=abi.FLAG_CLAMP_ROLLOUT, **kw)
    for o in (a, b):
        o.init(g[case + "/x0"], g[case + "/u0"])
        o.iterate(200)
    assert np.abs(a.get("us")).max() > 1.5 + 1e-6 and np.abs(b.get("us")).max() <= 1.5   # default = the reference's unclamped rollouts
    # (it does not converge better — "the wrong way, but the only way that works right now", src/ilqr_core.cpp:322 — which
    # is why it stays an option; its cost is the cost of controls inside the limits)
    assert np.isfinite(b.cost) and b.count("status") != 0
    # closed-form Jacobians agree with the central differences to their O(eps^2) truncation error
    c = O.OracleSolver(abi.MODEL_ACROBOT, 0.02)
    d = O.OracleSolver(abi.MODEL_ACROBOT, 0.02, flags=abi.FLAG_ANALYTIC_DYN)
    x0, u0 = g["acrobot_T200_b3/x0"], g["acrobot_T200_b3/u0"]
    for o in (c, d):
        o.init(x0, u0)
        o.backward_once(1.0)
    assert np.abs(c.get("fx") - d.get("fx")).max() < 2e-6 and np.abs(c.get("fu") - d.get("fu")).max() < 2e-6
    assert np.abs(c.get("fx") - d.get("fx")).max() > 1e-12                                 # and are not the same numbers


def test_warm_start_kernel_source_equals_oracle():
    rng = np.random.default_rng(3)
    x0, u0 = rng.uniform(-1, 1, 4), 0.5 * rng.uniform(-1, 1, (90, 1))
    o, e = O.OracleSolver(abi.MODEL_ACROBOT, 0.02), E.EmuSolver(abi.MODEL_ACROBOT, 0.02, libm=True)
    o.init(x0, u0), e.init(x0, u0)
    o.iterate(4), e.iterate(4)
    assert o.warm_start(x0 + 0.02) == e.warm_start(x0 + 0.02)
    o.iterate(3), e.iterate(3)
    for f in ARR:
        assert np.array_equal(o.get(f), e.get(f)), f


def test_boxqp_paths_against_reference_golden(golden_leaf):
    """Both boxQP code paths of the kernels (scalar m == 1, general m) against the reference's results."""
    g = golden_leaf
    for i in range(len(g["qp_m"])):
        m = int(g["qp_m"][i])
        args = (g["qp_Q"][i][:m, :m], g["qp_c"][i][:m], g["qp_x0"][i][:m], g["qp_lo"][i][:m], g["qp_hi"][i][:m])
        ores, ox, ovf, oR = O.boxqp(*args)
        for generic in ([False, True] if m == 1 else [True]):
            res, x, vf, R = E.boxqp(*args, generic=generic)
            assert res == ores == g["qp_result"][i]
            assert np.array_equal(x, ox) and (vf == ovf).all()
            assert np.allclose(x, g["qp_x"][i][:m], rtol=1e-12, atol=1e-14)
            if res != 6:
                assert np.array_equal(R, oR)


def test_scalar_boxqp_straight_line_equals_the_loop():
    """The m == 1 boxQP of the kernels is straight-line code (early exit when the warm start is clamped, two
    speculated Newton iterations, Armijo test without the division); it must return what the restated loop of
    src/boxqp.cpp returns — result code, x bit for bit, free flag — on every kind of problem: interior optimum,
    steps cut short by a bound (back-tracking), starts on / near a bound, tiny gradients, flat and non-convex Q,
    huge and tiny scales, steps that land within the 1e-4 clamp tolerance."""
    rng = np.random.default_rng(2026)
    n = 0
    seen = set()

    def check(Q, c, x0, lo, hi):
        nonlocal n
        args = (np.array([[Q]]), np.array([c]), np.array([x0]), np.array([lo]), np.array([hi]))
        ores, ox, ovf, oR = O.boxqp(*args)
        gres, gx, gvf, gR = E.boxqp(*args, generic=True)
        res, x, vf, R = E.boxqp(*args, generic=False)
        assert res == ores == gres, (Q, c, x0, lo, hi, res, ores, gres)
        assert x.tobytes() == ox.tobytes() == gx.tobytes(), (Q, c, x0, lo, hi, x, ox)
        assert (vf == ovf).all() and (vf == gvf).all(), (Q, c, x0, lo, hi)
        if res != 6:
            assert np.array_equal(R, oR)
        seen.add(int(res))
        n += 1

    for _ in range(4000):
        scale = 10.0 ** rng.uniform(-6, 6)
        Q = scale * 10.0 ** rng.uniform(-2, 2)
        c = scale * rng.normal() * 10.0 ** rng.uniform(-3, 3)
        lo, hi = sorted(rng.normal(size=2) * 10.0 ** rng.uniform(-2, 1))
        if lo == hi:
            hi = lo + 1.0
        kind = rng.integers(0, 6)
        if kind == 0:
            x0 = rng.uniform(lo, hi)
        elif kind == 1:
            x0 = lo if rng.random() < 0.5 else hi          # on a bound (the warm start of a clamped control)
        elif kind == 2:
            x0 = (lo if rng.random() < 0.5 else hi) + rng.uniform(-2e-4, 2e-4)  # within / just outside the clamp tolerance
        elif kind == 3:
            x0 = rng.normal() * 100                         # far outside: clamped first
        elif kind == 4:
            x0 = -c / Q + rng.normal() * 1e-9               # already at the optimum: gradient-small exits
        else:
            x0 = rng.uniform(lo, hi)
            c = -Q * (hi + (hi - lo) * 10.0 ** rng.uniform(-3, 3))  # optimum far beyond the upper bound: back-tracking
        check(Q, c, x0, lo, hi)
    # the Armijo threshold itself: steps cut to exactly the fraction theta of the Newton step, theta - theta^2 / 2 ~ 0.1
    for theta in np.concatenate([np.linspace(0.10, 0.112, 400), [0.1055728090000841]]):
        Q, x0, lo = 2.0, 0.0, -1.0
        for full in (1.0, 3.0, 1e3):
            c = -Q * full
            check(Q, c, x0, lo, theta * full)
    # degenerate curvature (Eigen's LLT leaves a non-positive pivot) and zero gradient
    for Q in (0.0, -1.0, 1e-300, 1e300):
        for c in (0.0, 1.0, -1.0):
            for x0 in (-1.0, 0.0, 0.5):
                check(Q, c, x0, -1.0, 1.0)
    assert n > 5000 and {2, 4, 5, 6} <= seen | {2, 4}, seen   # all exits of the shipped path were exercised
    assert {5, 6} <= seen


def test_trig_noise_floor():
    """How far the deterministic sincos moves an acrobot solve away from the libm oracle: this is the
    floor under every GPU-vs-oracle tolerance in test_gpu_parity.py."""
    import bench
    B, T = 24, 200
    x0, u0 = bench.synth_inputs_cpu(B, T, 12345)
    errK = {1: [], 5: [], 20: []}
    errc = []
    for b in range(B):
        o = O.OracleSolver(abi.MODEL_ACROBOT, 0.02, cost_deriv=abi.COST_ANALYTIC)
        e = E.EmuSolver(abi.MODEL_ACROBOT, 0.02, cost_deriv=abi.COST_ANALYTIC)
        o.init(x0[b], u0[b]), e.init(x0[b], u0[b])
        done = 0
        for n in (1, 5, 20):
            o.iterate(n - done), e.iterate(n - done)
            done = n
            errK[n].append(np.abs(o.get("K") - e.get("K")).max() / np.abs(o.get("K")).max())
        o.iterate(200), e.iterate(200)
        errc.append(abs(o.cost - e.cost) / abs(o.cost))
    assert max(errK[1]) < 1e-8 and max(errK[5]) < 1e-7
    assert np.mean(np.array(errK[20]) < 1e-6) >= 0.85
    assert np.mean(np.array(errc) < 1e-6) >= 0.9
