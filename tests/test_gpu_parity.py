"""GPU parity tests: the CUDA path, called through the C ABI (ilqr_b200.solver -> libilqr_b200.so),
against the CPU oracle (oracle/ilqr_oracle.c) on the same seeded inputs and against the golden
vectors produced by the unmodified reference (tests/golden/).

Tolerances.  BASELINE.json's north_star asks for 1e-6 relative on K, k and terminal cost.  The
f64 kernels reproduce the oracle's arithmetic order without FMA, so
  * the double integrator (no transcendental functions) must match the oracle BIT FOR BIT;
  * the acrobot differs only through libdevice-vs-libm sin/cos (<= 2 ulp).  That noise is amplified
    by the central differences (1 / 2eps = 500x) and by the Riccati recursion over 200 unstable
    steps: injecting +-1 ulp into the ORACLE's own sin/cos moves K by ~1e-11 relative after one
    trip, ~1e-10 after five, and after twenty trips about one instance in twenty has taken a
    different line-search branch (tests/test_emulator.py::test_trig_noise_floor measures this on
    the CPU).  So: every instance must agree to 1e-6 (relative to the array's own scale, with a
    1e-9 floor) for a single pass and for the 1- and 5-trip checkpoints; at 20 trips and at
    termination the bulk must agree to 1e-6 and a small fraction of bifurcated instances is
    tolerated (SURVEY.md §7 "Parity methodology").
"""
import os
import subprocess

import numpy as np
import pytest

from ilqr_b200 import abi
from ilqr_b200.solver import BatchILQR, make_inputs

import emuport as E
import oracleport as O

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-6, 1e-9


def inst_err(a, b, atol=ATOL):
    """per-instance error: max |a - b| over the instance's array, relative to max |b| of that array."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.ndim == 0:
        a, b = a[None], b[None]
    a2, b2 = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    err = np.abs(a2 - b2).max(axis=1)
    scale = np.maximum(np.abs(a2).max(axis=1), np.abs(b2).max(axis=1))
    return np.maximum(err - atol, 0.0) / np.maximum(scale, 1e-300)


def close(a, b, rtol=RTOL, atol=ATOL, frac=1.0):
    """at least `frac` of the instances (leading axis) agree to rtol relative to their own scale."""
    e = inst_err(a, b, atol)
    ok = e <= rtol
    assert ok.mean() >= frac, "only %d/%d instances within %.1e (worst %.3e at instance %d)" % (
        ok.sum(), ok.size, rtol, e.max(), int(e.argmax()))


def oracle_batch(model, x0, u0, dt, n_iters, what, **kw):
    """Run the oracle instance by instance; what(o) -> dict of arrays; returns dict of stacked arrays."""
    outs = []
    for b in range(x0.shape[0]):
        o = O.OracleSolver(model, dt, **kw)
        o.init(x0[b], u0[b])
        if n_iters:
            o.iterate(n_iters)
        outs.append(what(o))
    return {k: np.stack([d[k] for d in outs]) for k in outs[0]}


def snap(o):
    return dict(xs=o.get("xs"), us=o.get("us"), K=o.get("K"), k=o.get("k"), cost=np.float64(o.cost),
                lam=np.float64(o.scalar("lam")), trips=np.int64(o.count("loop_trips")),
                status=np.int64(o.count("status")), alpha_index=np.int64(o.count("alpha_index")))


def gpu_snap(s):
    return dict(xs=s.get("xs"), us=s.get("us"), K=s.get("K"), k=s.get("k"), cost=s.get("cost"), lam=s.get("lambda"),
                trips=s.get("iters"), status=s.get("status"), alpha_index=s.get("alpha_index"))


# ---------------------------------------------------------------------------------------------
# acrobot, default limits (BASELINE configs 1/2), FD and analytic cost derivatives
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cost_deriv", [abi.COST_FD, abi.COST_ANALYTIC])
def test_acrobot_single_phases(cost_deriv):
    B, T = 8, 200
    x0, u0 = make_inputs(12345, B, T, 4, 1)
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=cost_deriv)
    c0 = s.init_traj(x0, u0)
    ref = []
    for b in range(B):
        o = O.OracleSolver(abi.MODEL_ACROBOT, 0.02, cost_deriv=cost_deriv)
        ci = o.init(x0[b], u0[b])
        xs_i = o.get("xs")
        o.backward_once(1.0)
        d = dict(c0=np.float64(ci), xs0=xs_i, K=o.get("K"), k=o.get("k"), dV=o.get("dV"), Vx0=o.get("Vx")[0],
                 Vxx0=o.get("Vxx")[0], gnorm=np.float64(o.scalar("gnorm")), diverge=np.int64(o.count("diverge")))
        d["rc"] = np.float64(o.rollout_once(0.5012))
        d["xs1"], d["us1"] = o.get("xs"), o.get("us")
        ref.append(d)
    ref = {k: np.stack([d[k] for d in ref]) for k in ref[0]}
    close(c0, ref["c0"], 1e-12, 0)
    close(s.get("xs"), ref["xs0"], 1e-11, 1e-12)
    s.backward_once(1.0)
    assert (s.get("diverge") == ref["diverge"]).all()
    for f in ("K", "k", "dV", "Vx0", "Vxx0", "gnorm"):
        close(s.get(f), ref[f], 1e-7, 1e-10)
    s.rollout_once(0.5012)
    close(s.get("cost"), ref["rc"], 1e-7, 0)
    close(s.get("xs"), ref["xs1"], 1e-7, 1e-9)
    close(s.get("us"), ref["us1"], 1e-7, 1e-9)


@pytest.mark.parametrize("cost_deriv", [abi.COST_FD, abi.COST_ANALYTIC])
def test_acrobot_checkpoints(cost_deriv):
    """K, k, xs, us, cost after N = 1, 5, 20 loop trips (SURVEY.md §7 parity methodology (ii))."""
    B, T = 96, 200
    x0, u0 = make_inputs(12345, B, T, 4, 1)
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=cost_deriv)
    s.set_initial(x0, u0)
    done = 0
    # the reference's own FD cost Hessian carries ~5e-10 of rounding noise (SURVEY.md §7 "FD noise budget"),
    # which a 1-ulp change in x_T re-rolls: the FD-cost mode reaches the 1e-6 line a little earlier
    # gates = measured rates minus a margin (profiles/r2_attribution.md: K at 20 trips agrees for 97.5 % of 2048 instances in
    # closed-form mode and 96.5 % in FD mode, and the same fractions appear when the ORACLE's sin/cos is perturbed by 1 ulp)
    fracs = ((1, 1.0), (5, 0.95), (20, 0.9)) if cost_deriv == abi.COST_FD else ((1, 1.0), (5, 1.0), (20, 0.93))
    for n, frac in fracs:
        s.iterate(n - done)
        done = n
        ref = oracle_batch(abi.MODEL_ACROBOT, x0, u0, 0.02, n, snap, cost_deriv=cost_deriv)
        g = gpu_snap(s)
        assert np.mean(g["trips"] == ref["trips"]) >= frac
        assert np.mean(g["alpha_index"] == ref["alpha_index"]) >= frac
        for f in ("cost", "lam", "xs", "us", "K", "k"):
            close(g[f], ref[f], frac=frac)


def test_acrobot_termination():
    """Terminal cost at termination (parity methodology (iii)) on 512 instances, gated at the measured rate minus a
    margin: 98.1 % of 2048 instances end within 1e-6 of the oracle and 99.6 % within 1e-3, and perturbing the ORACLE's
    own sin/cos by one ulp gives the same two numbers (profiles/r2_attribution.md).  The last trips of a solve accept
    or reject on the SIGN of a cost change that is pure rounding noise (src/ilqr_core.cpp:206), so the trip count may
    legitimately differ by a few between two correct implementations; the terminal cost may not."""
    B, T = 512, 200
    x0, u0 = make_inputs(12345, B, T, 4, 1)
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
    s.generate_trajectory(x0, u0)
    r = O.solve_range(abi.make_desc(model=abi.MODEL_ACROBOT, T=T, dt=0.02, cost_deriv=abi.COST_ANALYTIC), x0, u0, 0, B)
    g = gpu_snap(s)
    assert (g["status"] != abi.RUNNING).all()
    close(g["cost"], r["cost"], frac=0.97)
    close(g["cost"], r["cost"], rtol=1e-3, frac=0.99)
    assert np.mean(g["trips"] == r["iters"]) >= 0.9
    # the reference's own derivative mode (finite-difference costs), a smaller sample with the full state compared
    B = 96
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02)
    s.generate_trajectory(x0[:B], u0[:B])
    ref = oracle_batch(abi.MODEL_ACROBOT, x0[:B], u0[:B], 0.02, 101, snap)
    g = gpu_snap(s)
    close(g["cost"], ref["cost"], frac=0.94)
    close(g["cost"], ref["cost"], rtol=1e-3, frac=0.97)
    assert np.mean(g["trips"] == ref["trips"]) >= 0.85
    close(g["xs"], ref["xs"], rtol=1e-5, frac=0.92)


def test_acrobot_golden_reference(golden_solver):
    """Against vectors written by the UNMODIFIED reference (tests/golden/make_golden.py)."""
    g = golden_solver
    cases = ["acrobot_T200_b%d" % b for b in range(6)]
    x0 = np.stack([g[c + "/x0"] for c in cases])
    u0 = np.stack([g[c + "/u0"] for c in cases])
    s = BatchILQR(abi.MODEL_ACROBOT, T=200, B=len(cases), dt=0.02)
    close(s.init_traj(x0, u0), [g[c + "/init_cost"] for c in cases], 1e-12, 0)
    s.backward_once(1.0)
    for f, name in (("K", "bw_K"), ("k", "bw_k"), ("dV", "bw_dV"), ("Vx0", "bw_Vx0"), ("Vxx0", "bw_Vxx0")):
        close(s.get(f), np.stack([g[c + "/" + name] for c in cases]), 1e-7, 1e-9)
    s.set_initial(x0, u0)
    done = 0
    # all six at 1 and 5 trips.  After 5 trips instance b3 (an ill-conditioned one: it ends on lambda > lambdaMax at cost 860)
    # sits at 1.2e-6 in K and 5e-6 in k, the other five below 1e-7: the gate there is 1e-5 on every instance and 1e-6 on five of six.
    # At 20 trips b2 is on a line-search branch point (tests/test_oracle_port.py finds the same for the oracle).
    for n, rtol, frac in ((1, 1e-6, 1.0), (5, 1e-5, 1.0), (5, 1e-6, 0.83), (20, 1e-6, 0.66)):
        s.iterate(n - done)
        done = n
        for f in ("K", "k", "xs", "us"):
            close(s.get(f), np.stack([g["%s/it%d_%s" % (c, n, f)] for c in cases]), rtol=rtol, frac=frac)
        close(s.get("cost"), [g["%s/it%d_cost" % (c, n)] for c in cases], rtol=rtol, frac=frac)
    s.solve()
    close(s.get("cost"), [g[c + "/final_cost"] for c in cases], frac=0.66)


def test_acrobot_cli_T499(golden_solver):
    """BASELINE config 1: the reference CLI's instance (src/run_ilqr.cpp:39-54), 100 iterations."""
    g = golden_solver
    c = "acrobot_cli_T499"
    s = BatchILQR(abi.MODEL_ACROBOT, T=499, B=1, dt=0.02)
    s.generate_trajectory(g[c + "/x0"][None], g[c + "/u0"][None])
    assert s.get("status")[0] == abi.EXIT_MAXITER and s.get("iters")[0] == 100
    close(s.get("cost"), g[c + "/final_cost"][None])
    close(s.get("K"), g[c + "/final_K"][None], 1e-5)
    close(s.get("k"), g[c + "/final_k"][None], 1e-5, 1e-7)
    close(s.get("xs"), g[c + "/final_xs"][None], 1e-5)


@pytest.mark.parametrize("cost_deriv,limits", [(abi.COST_FD, None), (abi.COST_ANALYTIC, None), (abi.COST_FD, 1.5)])
def test_acrobot_bit_exact_vs_kernel_source_on_cpu(cost_deriv, limits):
    """The GPU must execute the kernel source's arithmetic EXACTLY: tests/emu compiles the same
    header (ilqr_core.cuh, with the same deterministic sincos of trig.cuh) with g++ and runs the warp
    phases lane by lane; every number the GPU returns must equal it bit for bit, at every
    checkpoint up to termination.  (That CPU build is in turn bit-identical to the oracle when it
    is given the oracle's libm sin/cos: tests/test_emulator.py.)"""
    B, T = 10, 200
    x0, u0 = make_inputs(12345, B, T, 4, 1)
    kw = dict(u_min=[-limits], u_max=[limits]) if limits else {}
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=cost_deriv, **kw)
    s.set_initial(x0, u0)
    emus = []
    for b in range(B):
        e = E.EmuSolver(abi.MODEL_ACROBOT, 0.02, cost_deriv=cost_deriv, **kw)
        e.init(x0[b], u0[b])
        emus.append(e)
    assert (s.get("cost") == np.array([e.cost for e in emus])).all()
    done = 0
    for n in (1, 5, 20, 101):
        s.iterate(n - done)
        for e in emus:
            e.iterate(n - done)
        done = n
        g = gpu_snap(s)
        ref = {k: np.stack([snap(e)[k] for e in emus]) for k in g}
        for f in g:
            assert (g[f] == ref[f]).all(), (n, f)
    assert (s.get("status") != abi.RUNNING).all()


# ---------------------------------------------------------------------------------------------
# the two engines: batch-lockstep phase kernels (default) and the persistent warp-per-trajectory kernel
# ---------------------------------------------------------------------------------------------
ALL_FIELDS = ("xs", "us", "K", "k", "cost", "lambda", "dlambda", "dV", "gnorm", "Vx0", "Vxx0", "iters", "status",
              "alpha_index", "n_accept", "n_reject", "n_backward", "diverge")


@pytest.mark.parametrize("model,cd,dtype,T,B,kw", [
    (abi.MODEL_ACROBOT, abi.COST_ANALYTIC, abi.F64, 200, 300, {}),
    (abi.MODEL_ACROBOT, abi.COST_FD, abi.F64, 200, 70, {}),
    (abi.MODEL_ACROBOT, abi.COST_FD, abi.F64, 200, 70, dict(u_min=[-1.5], u_max=[1.5])),
    (abi.MODEL_ACROBOT, abi.COST_ANALYTIC, abi.F32, 500, 70, {}),
    (abi.MODEL_ACROBOT, abi.COST_ANALYTIC, abi.F64, 37, 33, {}),
    (abi.MODEL_DOUBLE_INTEGRATOR, abi.COST_FD, abi.F64, 60, 40, dict(goal=[1.0, 1.0, 0.0, 0.0])),
    (abi.MODEL_DOUBLE_INTEGRATOR, abi.COST_ANALYTIC, abi.F64, 99, 40, dict(goal=[0.5, -0.5, 0.0, 0.0])),
    (abi.MODEL_DOUBLE_INTEGRATOR, abi.COST_ANALYTIC, abi.F32, 60, 40, dict(goal=[1.0, 1.0, 0.0, 0.0])),
])
@pytest.mark.parametrize("head", ["rows", "thread", "warp"])
def test_phase_engine_equals_warp_engine_bit_for_bit(model, cd, dtype, T, B, kw, head, monkeypatch):
    """Both engines run the same arithmetic entry for entry (ilqr_phases.cuh uses Core's helpers), so every array,
    scalar and counter must agree BIT FOR BIT at every checkpoint up to termination, in every mode and dtype."""
    # the head of a trip (sweep + backward) has three lane decompositions, picked by the size of the active set
    # (ilqr_phase_launch.cuh); each is forced here in turn: 8 lanes per trajectory / 1 thread / 32 lanes
    monkeypatch.setenv("ILQR_B200_ROWS_MAX", "0" if head == "thread" else "1000000")
    monkeypatch.setenv("ILQR_B200_WARP_PRE_MAX", "1000000" if head == "warp" else "0")
    monkeypatch.setenv("ILQR_B200_HANDOVER", "0")  # lockstep rounds to the end (by default a batch this small goes straight
    #                                                to the persistent kernel, ilqr_phase_launch.cuh: kPhaseHandover)
    n, m = abi.MODEL_DIMS[model]
    x0, u0 = make_inputs(2024, B, T, n, m)
    dt = 0.02 if model == abi.MODEL_ACROBOT else 0.05
    ph = BatchILQR(model, T=T, B=B, dt=dt, cost_deriv=cd, dtype=dtype, **kw)
    wp = BatchILQR(model, T=T, B=B, dt=dt, cost_deriv=cd, dtype=dtype, flags=abi.FLAG_ENGINE_WARP, **kw)
    ph.set_initial(x0, u0)
    wp.set_initial(x0, u0)
    done = 0
    for n_it in (1, 4, 15, 101):
        ph.iterate(n_it - done)
        wp.iterate(n_it - done)
        done = n_it
        for f in ALL_FIELDS:
            a, b = ph.get(f), wp.get(f)
            assert np.array_equal(a, b), (n_it, f, int((a != b).reshape(B, -1).any(axis=1).sum()))
    assert (ph.get("status") != abi.RUNNING).all()
    # the phase engine really ran its own kernels: four launches per trip, not one per iterate call
    assert ph.launch_count > wp.launch_count + 8


@pytest.mark.parametrize("model,cd,dtype,T,kw", [
    (abi.MODEL_ACROBOT, abi.COST_FD, abi.F64, 200, {}),
    (abi.MODEL_ACROBOT, abi.COST_FD, abi.F64, 120, dict(u_min=[-1.5], u_max=[1.5])),
    (abi.MODEL_ACROBOT, abi.COST_ANALYTIC, abi.F32, 500, {}),
    (abi.MODEL_DOUBLE_INTEGRATOR, abi.COST_FD, abi.F64, 60, dict(goal=[1.0, 1.0, 0.0, 0.0])),
    (abi.MODEL_DOUBLE_INTEGRATOR, abi.COST_ANALYTIC, abi.F32, 60, dict(goal=[1.0, 1.0, 0.0, 0.0])),
])
def test_sixteen_lane_kernel_batch(model, cd, dtype, T, kw, monkeypatch):
    """the two-trajectories-per-warp instantiation of the warp kernel (ILQR_B200_LANES=16), batch of 96: FD mode, the
    double integrator (m = 2) and float, against the default engine bit for bit at every checkpoint"""
    B = 96
    n, m = abi.MODEL_DIMS[model]
    x0, u0 = make_inputs(515, B, T, n, m)
    dt = 0.02 if model == abi.MODEL_ACROBOT else 0.05
    ref = BatchILQR(model, T=T, B=B, dt=dt, cost_deriv=cd, dtype=dtype, **kw)
    monkeypatch.setenv("ILQR_B200_LANES", "16")
    s16 = BatchILQR(model, T=T, B=B, dt=dt, cost_deriv=cd, dtype=dtype, **kw)
    monkeypatch.delenv("ILQR_B200_LANES")
    ref.set_initial(x0, u0)
    s16.set_initial(x0, u0)
    done = 0
    for n_it in (1, 5, 101):
        ref.iterate(n_it - done)
        s16.iterate(n_it - done)
        done = n_it
        for f in ALL_FIELDS:
            assert np.array_equal(ref.get(f), s16.get(f)), (n_it, f)


@pytest.mark.parametrize("model,cd,dtype,T,kw", [
    (abi.MODEL_ACROBOT, abi.COST_ANALYTIC, abi.F64, 200, {}),
    (abi.MODEL_ACROBOT, abi.COST_FD, abi.F64, 120, dict(u_min=[-1.5], u_max=[1.5])),
    (abi.MODEL_ACROBOT, abi.COST_ANALYTIC, abi.F32, 300, {}),
    (abi.MODEL_DOUBLE_INTEGRATOR, abi.COST_FD, abi.F64, 60, dict(goal=[1.0, 1.0, 0.0, 0.0])),
])
@pytest.mark.parametrize("head", ["rows", "thread"])
def test_phase_engine_reroll_mode(model, cd, dtype, T, kw, head, monkeypatch):
    """the line search of large active sets: cost-only candidate rollouts, then ONE re-roll of the accepted candidate over
    xs / us (ilqr_phases.cuh: rollout_task kCostOnly / kInPlace) instead of eleven stored candidates and a copy.  Same
    operations, so the same bits as the store-all mode and as the warp engine, switching modes in mid-solve included."""
    B = 200
    n, m = abi.MODEL_DIMS[model]
    x0, u0 = make_inputs(808, B, T, n, m)
    dt = 0.02 if model == abi.MODEL_ACROBOT else 0.05
    monkeypatch.setenv("ILQR_B200_HANDOVER", "0")
    monkeypatch.setenv("ILQR_B200_ROWS_MAX", "0" if head == "thread" else "1000000")
    ref = BatchILQR(model, T=T, B=B, dt=dt, cost_deriv=cd, dtype=dtype, flags=abi.FLAG_ENGINE_WARP, **kw)
    ref.generate_trajectory(x0, u0)
    for reroll_min, check in (("0", "8"), ("120", "1")):      # always re-roll; re-roll until 120 are left, then store-all
        monkeypatch.setenv("ILQR_B200_REROLL_MIN", reroll_min)
        monkeypatch.setenv("ILQR_B200_CHECK_EVERY", check)
        s = BatchILQR(model, T=T, B=B, dt=dt, cost_deriv=cd, dtype=dtype, **kw)
        s.generate_trajectory(x0, u0)
        for f in ALL_FIELDS:
            assert np.array_equal(s.get(f), ref.get(f)), (reroll_min, f)


@pytest.mark.parametrize("model,cd,dtype,T,kw", [
    (abi.MODEL_ACROBOT, abi.COST_ANALYTIC, abi.F64, 200, {}),
    (abi.MODEL_ACROBOT, abi.COST_FD, abi.F64, 120, dict(u_min=[-1.5], u_max=[1.5])),
    (abi.MODEL_ACROBOT, abi.COST_ANALYTIC, abi.F32, 300, {}),
    (abi.MODEL_DOUBLE_INTEGRATOR, abi.COST_FD, abi.F64, 60, dict(goal=[1.0, 1.0, 0.0, 0.0])),
])
@pytest.mark.parametrize("stage_k,reroll", [(4, False), (1, False), (10, False), (3, True)])
def test_phase_engine_staged_line_search(model, cd, dtype, T, kw, stage_k, reroll, monkeypatch):
    """the line search of large active sets in two stages (ilqr_phases.cuh: PArgs::stage): the first stage_k candidates
    of every trajectory, then the remaining ones only where none of those passed.  The reference tries the candidates
    in order and stops at the first that passes (src/ilqr_core.cpp:186-214), so nothing it would have looked at is
    skipped: same bits as all-at-once and as the warp engine, counters (n_rollouts, alpha_index) included; also when
    the staging switches off in mid-solve and in the re-roll mode."""
    B = 200
    n, m = abi.MODEL_DIMS[model]
    x0, u0 = make_inputs(909, B, T, n, m)
    dt = 0.02 if model == abi.MODEL_ACROBOT else 0.05
    monkeypatch.setenv("ILQR_B200_HANDOVER", "0")
    ref = BatchILQR(model, T=T, B=B, dt=dt, cost_deriv=cd, dtype=dtype, flags=abi.FLAG_ENGINE_WARP, **kw)
    ref.generate_trajectory(x0, u0)
    monkeypatch.setenv("ILQR_B200_STAGE_K", str(stage_k))
    if reroll:
        monkeypatch.setenv("ILQR_B200_REROLL_MIN", "0")
    launches = []
    for stage_min, check in (("0", "8"), ("100", "1")):       # staged to the end; staged until 100 are left
        monkeypatch.setenv("ILQR_B200_STAGE_MIN", stage_min)
        monkeypatch.setenv("ILQR_B200_CHECK_EVERY", check)
        s = BatchILQR(model, T=T, B=B, dt=dt, cost_deriv=cd, dtype=dtype, **kw)
        s.generate_trajectory(x0, u0)
        for f in ALL_FIELDS:
            assert np.array_equal(s.get(f), ref.get(f)), (stage_min, f)
        launches.append(s.launch_count)
    assert launches[0] > launches[1]                          # two more launches per staged trip


@pytest.mark.parametrize("head", ["rows", "thread"])
@pytest.mark.parametrize("B", [77, 8192, 9000, 17000])
def test_phase_engine_ordered_active_list(B, head, monkeypatch):
    """the next trip's active list by order-preserving compaction (phase_compact_kernel) instead of atomic append: the
    same set of trajectories in ascending order, so the same bits; batch sizes below, across and above one pass of
    the compacting CTA (8192 entries)"""
    T = 60
    x0, u0 = make_inputs(31337, B, T, 4, 1)
    monkeypatch.setenv("ILQR_B200_HANDOVER", "0")
    monkeypatch.setenv("ILQR_B200_ROWS_MAX", "0" if head == "thread" else "1000000")
    ref = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC, flags=abi.FLAG_ENGINE_WARP)
    ref.generate_trajectory(x0, u0)
    for ordered_min in ("0", str(B // 2)):
        monkeypatch.setenv("ILQR_B200_ORDERED_MIN", ordered_min)
        monkeypatch.setenv("ILQR_B200_CHECK_EVERY", "1")
        s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
        s.generate_trajectory(x0, u0)
        for f in ALL_FIELDS:
            assert np.array_equal(s.get(f), ref.get(f)), (ordered_min, f)


@pytest.mark.parametrize("which", ["ILQR_B200_TEST_NO_CAND_MEMORY", "ILQR_B200_TEST_NO_PHASE_MEMORY"])
@pytest.mark.parametrize("cd", [abi.COST_ANALYTIC, abi.COST_FD])
def test_phase_engine_when_its_buffers_do_not_fit(which, cd, monkeypatch):
    """a batch whose per-trajectory work buffers do not fit in HBM beside its trajectories (ilqr_phase_launch.cuh):
    without the candidate buffers the line search re-rolls the accepted candidate; without the stored derivatives
    the persistent kernel runs the solve.  Neither is an error and both return the same bits.  (The variables
    make the allocations "fail" at a size the test can afford.)"""
    B, T = 150, 120
    x0, u0 = make_inputs(4242, B, T, 4, 1)
    monkeypatch.setenv("ILQR_B200_HANDOVER", "0")
    ref = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=cd)
    ref.generate_trajectory(x0, u0)
    monkeypatch.setenv(which, "1")
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=cd)
    s.generate_trajectory(x0, u0)
    for f in ALL_FIELDS:
        assert np.array_equal(s.get(f), ref.get(f)), f
    if which.endswith("PHASE_MEMORY"):
        assert s.launch_count < ref.launch_count          # one persistent launch instead of lockstep rounds
    else:
        assert s.launch_count > ref.launch_count          # the extra commit kernel of the re-roll mode
    s.warm_start(x0 + 0.01)
    ref.warm_start(x0 + 0.01)
    monkeypatch.delenv(which)                              # the decision, once taken, holds for the handle
    s.generate_trajectory()
    ref.generate_trajectory()
    for f in ALL_FIELDS:
        assert np.array_equal(s.get(f), ref.get(f)), f


@pytest.mark.parametrize("handover,check", [(150, 1), (250, 3), (40, 8)])
def test_phase_engine_hands_the_tail_to_the_persistent_kernel(handover, check, monkeypatch):
    """lockstep rounds while many trajectories run, then the persistent warp kernel for the survivors' remaining trips
    (the default for large batches): same bits as either engine alone, wherever the switch happens"""
    B, T = 300, 200
    x0, u0 = make_inputs(77, B, T, 4, 1)
    monkeypatch.setenv("ILQR_B200_HANDOVER", str(handover))
    monkeypatch.setenv("ILQR_B200_CHECK_EVERY", str(check))
    a = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
    b = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC, flags=abi.FLAG_ENGINE_WARP)
    a.generate_trajectory(x0, u0)
    b.generate_trajectory(x0, u0)
    for f in ALL_FIELDS:
        assert np.array_equal(a.get(f), b.get(f)), f
    assert a.launch_count > b.launch_count + 8        # it did run lockstep rounds first
    a.set_initial(x0, u0)
    b.set_initial(x0, u0)
    a.iterate(30)
    b.iterate(30)
    for f in ALL_FIELDS:
        assert np.array_equal(a.get(f), b.get(f)), f


@pytest.mark.parametrize("engine_flags,handover", [(0, "0"), (0, "100000"), (abi.FLAG_ENGINE_WARP, "0")])
@pytest.mark.parametrize("model,cd,T,kw", [(abi.MODEL_ACROBOT, abi.COST_ANALYTIC, 200, {}), (abi.MODEL_ACROBOT, abi.COST_FD, 200, {}),
                                           (abi.MODEL_DOUBLE_INTEGRATOR, abi.COST_FD, 60, dict(goal=[1.0, 1.0, 0.0, 0.0]))])
def test_fast_fma_build_within_1e_6(model, cd, T, kw, engine_flags, handover, monkeypatch):
    """ILQR_FLAG_FAST_FMA: the same kernels compiled WITH fused multiply-add contraction.  Not bit-identical to the
    reference's arithmetic any more, but within BASELINE's 1e-6 on K, k and cost of the ORACLE after 1 and 5 trips on
    every instance, on both engines."""
    monkeypatch.setenv("ILQR_B200_HANDOVER", handover)
    B = 128
    n, m = abi.MODEL_DIMS[model]
    x0, u0 = make_inputs(99, B, T, n, m)
    dt = 0.02 if model == abi.MODEL_ACROBOT else 0.05
    s = BatchILQR(model, T=T, B=B, dt=dt, cost_deriv=cd, flags=abi.FLAG_FAST_FMA | engine_flags, **kw)
    s.set_initial(x0, u0)
    done = 0
    for n_it in (1, 5):
        s.iterate(n_it - done)
        done = n_it
        ref = oracle_batch(model, x0, u0, dt, n_it, snap, cost_deriv=cd, **kw)
        g = gpu_snap(s)
        assert (g["alpha_index"] == ref["alpha_index"]).mean() >= 0.98
        for f in ("cost", "lam", "xs", "us", "K", "k"):
            # measured: 128 of 128 (acrobot, both modes); the double integrator in FD-cost mode has ONE instance of 128
            # that sits on a branch of the m = 2 boxQP after five trips (the FD cost Hessian divides the FMA rounding
            # difference by 4 eps^2), every other one within 1e-6
            close(g[f], ref[f], frac=1.0 if (n_it == 1 or model == abi.MODEL_ACROBOT) else 0.98)
    s.solve()
    assert (s.get("status") != abi.RUNNING).all()
    ref = oracle_batch(model, x0, u0, dt, 101, snap, cost_deriv=cd, **kw)
    close(s.get("cost"), ref["cost"], frac=0.93)


@pytest.mark.parametrize("engine_flags,handover", [(0, "0"), (abi.FLAG_ENGINE_WARP, "0")])
@pytest.mark.parametrize("model,flags,kw", [
    (abi.MODEL_ACROBOT, abi.FLAG_CLAMP_ROLLOUT, dict(u_min=[-1.5], u_max=[1.5])),
    (abi.MODEL_ACROBOT, abi.FLAG_ANALYTIC_DYN, {}),
    (abi.MODEL_ACROBOT, abi.FLAG_ANALYTIC_DYN | abi.FLAG_CLAMP_ROLLOUT, dict(u_min=[-1.5], u_max=[1.5])),
    (abi.MODEL_DOUBLE_INTEGRATOR, abi.FLAG_ANALYTIC_DYN | abi.FLAG_CLAMP_ROLLOUT, dict(goal=[1.0, 1.0, 0.0, 0.0]))])
def test_opt_in_modes_vs_oracle(model, flags, kw, engine_flags, handover, monkeypatch):
    """clamped rollouts (src/ilqr_core.cpp:322-329 "the right way") and closed-form dynamics Jacobians
    (notes.md:15,45), both off by default.  GPU against the oracle's implementation of the same modes: 1e-6 on every
    instance after 1 and 5 trips (acrobot; sin/cos differ), bit for bit for the double integrator."""
    monkeypatch.setenv("ILQR_B200_HANDOVER", handover)
    B, T = 48, 120
    n, m = abi.MODEL_DIMS[model]
    x0, u0 = make_inputs(4, B, T, n, m)
    dt = 0.02 if model == abi.MODEL_ACROBOT else 0.05
    s = BatchILQR(model, T=T, B=B, dt=dt, flags=flags | engine_flags, **kw)
    s.set_initial(x0, u0)
    done = 0
    for n_it in (1, 5):
        s.iterate(n_it - done)
        done = n_it
        ref = oracle_batch(model, x0, u0, dt, n_it, snap, flags=flags, **kw)
        g = gpu_snap(s)
        for f in ("cost", "lam", "xs", "us", "K", "k"):
            if model == abi.MODEL_DOUBLE_INTEGRATOR:
                assert np.array_equal(g[f], ref[f]), (n_it, f)
            else:   # a clamp is a kink: one instance in 48 has taken another line-search branch by the fifth trip
                close(g[f], ref[f], frac=1.0 if n_it == 1 else 0.95)
    if flags & abi.FLAG_CLAMP_ROLLOUT:
        lim = 1.5 if model == abi.MODEL_ACROBOT else 0.5
        s.solve()
        assert np.abs(s.get("us")).max() <= lim
        d = BatchILQR(model, T=T, B=B, dt=dt, **kw)                           # the default stays bug-compatible: unclamped
        d.generate_trajectory(x0, u0)
        if model == abi.MODEL_ACROBOT:
            assert np.abs(d.get("us")).max() > lim


def test_phase_engine_iterate_resume_and_warm_start(monkeypatch):
    """iterate in uneven chunks (active list rebuilt by each call), then warm start and continue: equal to one call"""
    monkeypatch.setenv("ILQR_B200_HANDOVER", "0")
    B, T = 96, 120
    x0, u0 = make_inputs(31, B, T, 4, 1)
    a = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
    b = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC, flags=abi.FLAG_ENGINE_WARP)
    for s in (a, b):
        s.set_initial(x0, u0)
    for chunk in (3, 1, 9, 2):
        a.iterate(chunk)
    b.iterate(15)
    for f in ALL_FIELDS:
        assert np.array_equal(a.get(f), b.get(f)), f
    for s in (a, b):
        s.warm_start(x0 + 0.01)
        s.iterate(0)
        s.iterate(6)
    for f in ALL_FIELDS:
        assert np.array_equal(a.get(f), b.get(f)), f


# ---------------------------------------------------------------------------------------------
# control-limited acrobot (BASELINE config 4): boxQP path hot, lambda-max exits
# ---------------------------------------------------------------------------------------------
def test_acrobot_control_limited(golden_solver):
    B, T = 16, 200
    x0, u0 = make_inputs(12345, B, T, 4, 1)
    kw = dict(u_min=[-1.5], u_max=[1.5])
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, **kw)
    s.set_initial(x0, u0)
    done = 0
    for n in (1, 5):
        s.iterate(n - done)
        done = n
        ref = oracle_batch(abi.MODEL_ACROBOT, x0, u0, 0.02, n, snap, **kw)
        g = gpu_snap(s)
        for f in ("cost", "lam", "xs", "us", "K", "k"):
            close(g[f], ref[f])
    s.solve()
    ref = oracle_batch(abi.MODEL_ACROBOT, x0, u0, 0.02, 101, snap, **kw)
    close(s.get("cost"), ref["cost"], frac=0.85)
    assert (s.get("status") == ref["status"]).mean() >= 0.75
    # the reference's golden: the rollouts are NOT clamped (src/ilqr_core.cpp:322-329)
    gold = golden_solver
    cases = ["acrobot_lim15_T200_b%d" % b for b in range(3)]
    close(s.get("cost")[:3], [gold[c + "/final_cost"] for c in cases], frac=0.66)
    assert np.abs(s.get("us")).max() > 1.5


# ---------------------------------------------------------------------------------------------
# double integrator: no sin/cos, so the CUDA path must equal the oracle bit for bit
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cost_deriv", [abi.COST_FD, abi.COST_ANALYTIC])
def test_double_integrator_bit_exact(cost_deriv):
    B, T, dt, goal = 12, 60, 0.05, [1.0, 1.0, 0.0, 0.0]
    x0, u0 = make_inputs(12345, B, T, 4, 2, canonical_first=False)
    s = BatchILQR(abi.MODEL_DOUBLE_INTEGRATOR, T=T, B=B, dt=dt, goal=goal, cost_deriv=cost_deriv)
    s.set_initial(x0, u0)
    ref0 = oracle_batch(abi.MODEL_DOUBLE_INTEGRATOR, x0, u0, dt, 0, snap, goal=goal, cost_deriv=cost_deriv)
    assert (s.get("cost") == ref0["cost"]).all() and (s.get("xs") == ref0["xs"]).all()
    done = 0
    for n in (1, 3, 8, 101):
        s.iterate(n - done)
        done = n
        ref = oracle_batch(abi.MODEL_DOUBLE_INTEGRATOR, x0, u0, dt, n, snap, goal=goal, cost_deriv=cost_deriv)
        g = gpu_snap(s)
        for f in ("trips", "status", "alpha_index", "cost", "lam", "xs", "us", "K", "k"):
            assert (g[f] == ref[f]).all(), (n, f, np.abs(np.asarray(g[f], float) - np.asarray(ref[f], float)).max())


def test_double_integrator_golden_cli(golden_solver):
    g = golden_solver
    c = "integrator_cli_T99"
    s = BatchILQR(abi.MODEL_DOUBLE_INTEGRATOR, T=99, B=1, dt=0.02, goal=list(g[c + "/goal"]))
    close(s.init_traj(g[c + "/x0"][None], g[c + "/u0"][None]), g[c + "/init_cost"][None], 1e-12, 0)
    s.backward_once(1.0)
    close(s.get("dV"), g[c + "/bw_dV"][None], 1e-8)
    close(s.get("k"), g[c + "/bw_k"][None], 1e-8, 1e-10)
    close(s.get("K"), g[c + "/bw_K"][None], 1e-7, 1e-9)
    s.generate_trajectory(g[c + "/x0"][None], g[c + "/u0"][None])
    close(s.get("cost"), g[c + "/final_cost"][None])


# ---------------------------------------------------------------------------------------------
# API semantics: warm start, iterate granularity, errors
# ---------------------------------------------------------------------------------------------
def test_warm_start_matches_oracle():
    B, T = 4, 120
    x0, u0 = make_inputs(777, B, T, 4, 1, canonical_first=False)
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02)
    s.set_initial(x0, u0)
    s.iterate(6)
    x0b = x0 + 0.01
    s.warm_start(x0b)
    cw = s.get("cost")
    s.iterate(3)
    for b in range(B):
        o = O.OracleSolver(abi.MODEL_ACROBOT, 0.02)
        o.init(x0[b], u0[b])
        o.iterate(6)
        close(cw[b:b + 1], [o.warm_start(x0b[b])])
        o.iterate(3)
        close(s.get("cost")[b:b + 1], [o.cost])
        close(s.get("xs")[b:b + 1], o.get("xs")[None])
        close(s.get("K")[b:b + 1], o.get("K")[None])


@pytest.mark.parametrize("case", ["acrobot_warm_b0", "acrobot_warm_b1"])
def test_warm_start_and_resume_against_reference_golden(golden_solver, case):
    """ilqr_warm_start = iLQR::generate_trajectory(x_0) (src/ilqr_core.cpp:65-76), ilqr_resume = a repeated
    generate_trajectory() (:78-102), lambda / dlambda carried over: against vectors written by the unmodified reference
    (replica checkpoints, and the terminal costs of the reference's OWN warm-start / continue calls)"""
    g = golden_solver
    x0, u0 = g[case + "/x0"], g[case + "/u0"]
    s = BatchILQR(abi.MODEL_ACROBOT, T=u0.shape[0], B=2, dt=float(g[case + "/dt"]))   # the same instance twice
    s.generate_trajectory(np.stack([x0, x0]), np.stack([u0, u0]))
    close(s.get("cost"), [g[case + "/first_cost"]] * 2)
    close(s.get("lambda"), [g[case + "/first_lambda"]] * 2, 1e-9, 1e-300)
    s.warm_start(np.stack([g[case + "/x0_warm"]] * 2))
    close(s.get("cost"), [g[case + "/warm_cost"]] * 2)
    close(s.get("xs"), np.stack([g[case + "/warm_xs"]] * 2))
    close(s.get("us"), np.stack([g[case + "/warm_us"]] * 2))
    done = 0
    for n in (1, 3, 10):
        s.iterate(n - done)
        done = n
        close(s.get("cost"), [g["%s/warm_it%d_cost" % (case, n)]] * 2)
        close(s.get("lambda"), [g["%s/warm_it%d_lambda" % (case, n)]] * 2, 1e-9, 1e-300)
        for f, name in (("K", "K"), ("k", "k"), ("xs", "xs"), ("us", "us")):
            close(s.get(f), np.stack([g["%s/warm_it%d_%s" % (case, n, name)]] * 2), 1e-5)
    s.solve()
    close(s.get("cost"), [g[case + "/warm_final_cost_native"]] * 2)
    s.generate_trajectory()                                                    # resume + solve
    close(s.get("cost"), [g[case + "/resume_final_cost_native"]] * 2)
    assert (s.get("status") != abi.RUNNING).all()


def test_resume_reenters_the_loop():
    """a solve that ran out of iterations continues for up to max_iter more after ilqr_resume (the reference's repeated
    generate_trajectory(), src/ilqr_core.cpp:78-102); without it a finished handle stays finished.  GPU == oracle."""
    p = abi.default_params()
    p.max_iter = 6
    B, T = 5, 80
    x0, u0 = make_inputs(5, B, T, 4, 1, canonical_first=False)
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, params=p)
    s.generate_trajectory(x0, u0)
    assert (s.get("status") == abi.EXIT_MAXITER).all() and (s.get("iters") == 6).all()
    c6 = s.get("cost").copy()
    s.solve()                                                                  # nothing is running: a no-op
    assert np.array_equal(s.get("cost"), c6) and (s.get("iters") == 6).all()
    s.generate_trajectory()
    assert (s.get("iters") > 6).all() and (s.get("cost") <= c6).all() and (s.get("cost") < c6).any()
    for b in range(B):
        o = O.OracleSolver(abi.MODEL_ACROBOT, 0.02, params=p)
        o.init(x0[b], u0[b])
        o.iterate(100)
        o.resume()
        o.iterate(100)
        close(s.get("cost")[b:b + 1], [o.cost])
        assert s.get("iters")[b] == o.count("loop_trips")


def test_iterate_granularity_and_determinism():
    """iterate(1) x N == iterate(N) bit for bit; a trajectory's result does not depend on the batch around it."""
    B, T = 64, 200
    x0, u0 = make_inputs(12345, B, T, 4, 1)
    a = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
    a.set_initial(x0, u0)
    a.iterate(7)
    b = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
    b.set_initial(x0, u0)
    for _ in range(7):
        b.iterate(1)
    for f in ("xs", "us", "K", "k", "cost", "lambda", "iters"):
        assert (a.get(f) == b.get(f)).all(), f
    sub = slice(5, 9)
    c = BatchILQR(abi.MODEL_ACROBOT, T=T, B=4, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
    c.set_initial(x0[sub], u0[sub])
    c.iterate(7)
    for f in ("xs", "us", "K", "k", "cost"):
        assert (a.get(f)[sub] == c.get(f)).all(), f


def test_errors():
    with pytest.raises(Exception):
        BatchILQR(model=7, T=10, B=1)
    s = BatchILQR(abi.MODEL_ACROBOT, T=10, B=2)
    with pytest.raises(Exception):
        s.iterate(1)  # before set_initial
    with pytest.raises(Exception):
        s.get("cost")


# ---------------------------------------------------------------------------------------------
# full-size properties (BASELINE config 2: B = 4096, T = 200, f64, analytic cost derivatives)
# ---------------------------------------------------------------------------------------------
def test_full_size_config2_properties():
    B, T = 4096, 200
    x0, u0 = make_inputs(12345, B, T, 4, 1)
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
    c0 = s.init_traj(x0, u0)
    s.solve()
    cost, status, trips = s.get("cost"), s.get("status"), s.get("iters")
    acc, rej = s.get("n_accept"), s.get("n_reject")
    assert (status != abi.RUNNING).all()
    assert np.isfinite(cost).all() and (cost <= c0 + 1e-9).all()          # monotone: only improving steps are accepted
    assert ((acc + rej == trips) | (status == abi.EXIT_GRAD)).all()
    assert (trips <= 100).all() and (trips >= 1).all()
    # rolling the returned controls out open-loop reproduces the returned states and cost (consistency of xs/us/cost)
    xs, us = s.get("xs"), s.get("us")
    idx = np.random.default_rng(0).choice(B, 24, replace=False)
    idx2 = np.arange(0, B, 8)  # 512 instances against the oracle's own solves
    for b in idx:
        o = O.OracleSolver(abi.MODEL_ACROBOT, 0.02, cost_deriv=abi.COST_ANALYTIC)
        c = o.init(xs[b, 0], us[b])
        close([c], cost[b:b + 1], 1e-9, 1e-9)
        close(o.get("xs")[None], xs[b:b + 1], 1e-8, 1e-8)
    # a sample of instances against the oracle's own solves
    r = O.solve_range(abi.make_desc(model=abi.MODEL_ACROBOT, T=T, dt=0.02, cost_deriv=abi.COST_ANALYTIC),
                      np.ascontiguousarray(x0[idx2]), np.ascontiguousarray(u0[idx2]), 0, len(idx2))
    close(cost[idx2], r["cost"], frac=0.97)
    close(cost[idx2], r["cost"], rtol=1e-3, frac=0.99)
    # re-running is bit-reproducible
    s2 = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
    s2.generate_trajectory(x0, u0)
    assert (s2.get("cost") == cost).all() and (s2.get("iters") == trips).all()


def _f32_inputs(B, T):
    """instances whose f32 rounding is exact in f64, so the oracle and the f32 kernels start from the same numbers"""
    x0, u0 = make_inputs(12345, B, T, 4, 1)
    return x0.astype(np.float32).astype(np.float64), u0.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("T,budget1,budget5", [(200, dict(K=1e-4, k=2e-3, cost=1e-3, same=0.95), dict(med_K=5e-4, frac=0.85)),
                                               (500, dict(K=1e-3, k=2e-2, cost=3e-2, same=0.9), dict(med_K=5e-3, frac=0.5))])
@pytest.mark.parametrize("head", ["rows", "thread"])
def test_f32_parity_vs_f64_oracle(T, budget1, budget5, head, monkeypatch):
    """BASELINE configs[2] arithmetic (f32, FD fx/fu, closed-form cost derivatives) against the f64 ORACLE on every
    instance, at a stated f32 budget.  The finite differences are formed in double and rounded to f32 (ilqr_core.cuh,
    FiniteDiff): with f32 differences the Jacobians carry 3e-5 of noise and K is off by 2e-3 (median) after ONE trip and
    by O(1) after five; with this scheme (measured on the CPU build of the same source, 24 instances) the first trip
    agrees to K 2e-5, k 5e-4, cost 2e-4 in the worst instance.  Later trips drift apart at f32 rounding amplified by the
    unstable recursion and a growing fraction of instances takes another line-search branch, exactly as f64-vs-f64 does
    at 1-ulp level (profiles/r2_attribution.md) but from a 1e-7 instead of a 1e-16 seed.  Gates = the GPU measurement
    on these 64 instances (profiles/r2_f32_parity.txt, tools/exp_f32_parity.py) with a margin: after one trip every
    same-branch instance within K 2.3e-5 / k 3.5e-4 / cost 1.9e-4 (T = 200) and 2.3e-4 / 5.6e-3 / 9.4e-3 (T = 500); after
    five trips the median K within 3.6e-5 / 5.1e-4 and 95 % / 67 % of the instances within 5e-2."""
    monkeypatch.setenv("ILQR_B200_ROWS_MAX", "0" if head == "thread" else "1000000")
    monkeypatch.setenv("ILQR_B200_HANDOVER", "0")
    B = 64
    x0, u0 = _f32_inputs(B, T)
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, dtype=abi.F32, cost_deriv=abi.COST_ANALYTIC)
    c0 = s.init_traj(x0, u0)
    ref0 = oracle_batch(abi.MODEL_ACROBOT, x0, u0, 0.02, 0, snap, cost_deriv=abi.COST_ANALYTIC)
    close(c0, ref0["cost"], 2e-4, 1e-3, frac=0.95)   # a 200- to 500-step rollout in f32
    close(c0, ref0["cost"], 2e-3, 1e-3)
    s.iterate(1)
    ref = oracle_batch(abi.MODEL_ACROBOT, x0, u0, 0.02, 1, snap, cost_deriv=abi.COST_ANALYTIC)
    g = gpu_snap(s)
    assert (g["alpha_index"] == ref["alpha_index"]).mean() >= budget1["same"]
    same = g["alpha_index"] == ref["alpha_index"]          # an instance on another line-search branch is not comparable
    close(g["K"][same], ref["K"][same], budget1["K"], 1e-6)
    close(g["k"][same], ref["k"][same], budget1["k"], 1e-6)
    close(g["cost"][same], ref["cost"][same], budget1["cost"], 0)
    s.iterate(4)
    ref = oracle_batch(abi.MODEL_ACROBOT, x0, u0, 0.02, 5, snap, cost_deriv=abi.COST_ANALYTIC)
    g = gpu_snap(s)
    eK = inst_err(g["K"], ref["K"], 1e-6)
    assert np.median(eK) <= budget5["med_K"], np.median(eK)
    assert (eK <= 5e-2).mean() >= budget5["frac"], (eK <= 5e-2).mean()
    assert (inst_err(g["cost"], ref["cost"], 0) <= 1e-2).mean() >= budget5["frac"]


def test_f32_config3_termination_statistics():
    """Where f32 solves end, next to the f64 oracle on the same instances.  At T = 500 the reference's own 100-iteration
    cap is what stops most solves IN F64 TOO (21 of 24 in the CPU measurement): the MAXITER share of configs[2] is a
    property of the problem, not of the arithmetic.  At T = 200 the f32 solves end within 1e-2 of the f64 cost for most
    instances and never above the initial cost."""
    B, T = 48, 200
    x0, u0 = _f32_inputs(B, T)
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, dtype=abi.F32, cost_deriv=abi.COST_ANALYTIC)
    c0 = s.init_traj(x0, u0)
    s.solve()
    ref = oracle_batch(abi.MODEL_ACROBOT, x0, u0, 0.02, 101, snap, cost_deriv=abi.COST_ANALYTIC)
    c1 = s.get("cost")
    assert np.isfinite(c1).all() and (c1 <= c0).all() and (s.get("status") != abi.RUNNING).all()
    assert (inst_err(c1, ref["cost"], 0) <= 1e-2).mean() >= 0.6
    assert abs(s.get("iters").mean() - ref["trips"].mean()) <= 12
    # T = 500: the iteration cap stops the f64 oracle as well
    B5, T5 = 12, 500
    x5, u5 = _f32_inputs(B5, T5)
    s5 = BatchILQR(abi.MODEL_ACROBOT, T=T5, B=B5, dt=0.02, dtype=abi.F32, cost_deriv=abi.COST_ANALYTIC)
    s5.generate_trajectory(x5, u5)
    ref5 = oracle_batch(abi.MODEL_ACROBOT, x5, u5, 0.02, 101, snap, cost_deriv=abi.COST_ANALYTIC)
    assert (ref5["status"] == abi.EXIT_MAXITER).mean() >= 0.6 and (s5.get("status") == abi.EXIT_MAXITER).mean() >= 0.6


# ---------------------------------------------------------------------------------------------
# the C++ host layer: the reference's UNCHANGED src/run_ilqr.cpp, acrobot.h, double_integrator.h
# compiled against ilqr_b200/host/ilqr.h (prebuilt in the dev container, ilqr_b200/host/Makefile)
# ---------------------------------------------------------------------------------------------
HOST_BUILD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ilqr_b200", "host", "_build")



@pytest.mark.skipif(not os.path.exists(os.path.join(HOST_BUILD, "run_iLQR")), reason="host binaries not built")
@pytest.mark.parametrize("which,case", [("acrobot", "acrobot_cli_T499"), ("integrator", "integrator_cli_T99")])
def test_reference_cli_runs_on_the_gpu_path(tmp_path, golden_solver, which, case):
    """BASELINE configs[0]: `./run_iLQR acrobot` — the reference's own main(), on libilqr_b200.so."""
    out = subprocess.run([os.path.join(HOST_BUILD, "run_iLQR"), which], cwd=tmp_path, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr
    assert "Run iLQR!" in out.stdout and "iLQR took:" in out.stdout          # src/run_ilqr.cpp:57,62
    # the result file (src/ilqr_core.cpp:300,414-431) has the reference's exact format — same header bytes, same row
    # structure, the terminal row ending in ", " with no newline — and is read by plot_results.py's own logic
    from ilqr_b200 import export
    g = golden_solver
    xs, us = g[case + "/final_xs"], g[case + "/final_us"]
    ours = open(tmp_path / "ilqr_result.csv", "rb").read()
    ref = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cli_%s.csv" % which), "rb").read()
    assert ours.split(b"\n")[0] == ref.split(b"\n")[0]
    assert ours.count(b"\n") == ref.count(b"\n") and ours.endswith(b", ") and not ours.endswith(b"\n")
    assert [l.count(b",") for l in ours.split(b"\n")] == [l.count(b",") for l in ref.split(b"\n")]
    states, controls = export.read_csv(tmp_path / "ilqr_result.csv", xs.shape[1], us.shape[1])
    assert states.shape == xs.shape and controls.shape == us.shape           # T + 1 states, T controls
    assert np.allclose(states[-1], xs[-1], atol=2e-6)                         # %f keeps 6 decimals
    assert np.allclose(states, xs, atol=5e-5) and np.allclose(controls, us, atol=5e-4)
    assert "Saved iLQR result to ilqr_result.csv" in out.stdout               # :430
    cost = float(out.stdout.split("cost ")[-1].split()[0])
    assert abs(cost - g[case + "/final_cost"]) <= 1e-6 * abs(cost)


@pytest.mark.skipif(not os.path.exists(os.path.join(HOST_BUILD, "batch_demo")), reason="host binaries not built")
def test_host_batch_entry_point(tmp_path, golden_solver):
    out = subprocess.run([os.path.join(HOST_BUILD, "batch_demo"), "32", "200"], cwd=tmp_path, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.strip().split("\n")]
    g = golden_solver
    ok = 0
    for b in range(6):
        cost = float(lines[b][3])
        ok += abs(cost - g["acrobot_T200_b%d/final_cost" % b]) <= 1e-6 * abs(cost)
    assert ok >= 4                                                            # FD-cost mode, bifurcations tolerated
    single = lines[6]
    assert abs(float(single[3]) - float(lines[0][3])) <= 1e-9 * abs(float(lines[0][3]))  # single API == batch entry 0


@pytest.mark.skipif(not os.path.exists(os.path.join(HOST_BUILD, "batch_demo")), reason="host binaries not built")
def test_host_batch_export(tmp_path):
    """iLQR::export_batch (whole batch, binary) and iLQR::output_to_csv(file, b) (one trajectory, the reference's CSV)"""
    from ilqr_b200 import export
    out = subprocess.run([os.path.join(HOST_BUILD, "batch_demo"), "24", "120", str(tmp_path / "res")], cwd=tmp_path,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    d = export.read_batch(tmp_path / "res.bin")
    assert d["xs"].shape == (24, 121, 4) and d["us"].shape == (24, 120, 1)
    lines = [l.split() for l in out.stdout.strip().split("\n")]
    for b in range(6):
        assert abs(float(lines[b][3]) - d["cost"][b]) <= 1e-9 * abs(d["cost"][b]) and int(lines[b][5]) == d["iters"][b]
    assert (d["status"] != abi.RUNNING).all()
    states, controls = export.read_csv(tmp_path / "res_b3.csv", 4, 1)
    assert np.abs(states - d["xs"][3]).max() <= 5.1e-7 and np.abs(controls - d["us"][3]).max() <= 5.1e-7
    # the same two files through the Python host layer
    x0, u0 = make_inputs(12345, 24, 120, 4, 1)
    s = BatchILQR(abi.MODEL_ACROBOT, T=120, B=24, dt=0.02)
    s.generate_trajectory(x0, u0)
    s.export_batch(tmp_path / "py.bin")
    s.output_to_csv(tmp_path / "py_b3.csv", 3)
    assert open(tmp_path / "py.bin", "rb").read() == open(tmp_path / "res.bin", "rb").read()
    assert open(tmp_path / "py_b3.csv", "rb").read() == open(tmp_path / "res_b3.csv", "rb").read()


@pytest.mark.skipif(not os.path.exists(os.path.join(HOST_BUILD, "bench_batch")), reason="host binaries not built")
def test_cpp_multi_gpu_host_path(tmp_path):
    """BatchSolver (ilqr_b200/host/batch_solver.h): one process, one handle + host thread per visible GPU, pinned inputs, one
    ncclAllGather of the final costs.  Its results are the Python host layer's, and iLQR::solve_batch gives the same
    answers through it (iLQR::devices) as through the plain handle."""
    import json
    import torch
    B, T = 1536, 200
    out = subprocess.run([os.path.join(HOST_BUILD, "bench_batch"), str(B), str(T), "1", "1"], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().split("\n")[-1])
    assert r["n_gpus"] == torch.cuda.device_count() and r["batch_total"] == B
    x0, u0 = make_inputs(12345, B, T, 4, 1)
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
    s.generate_trajectory(x0, u0)
    cost = s.get("cost")
    assert abs(r["cost0"] - cost[0]) <= 1e-12 * abs(cost[0])
    assert abs(r["cost_checksum"] - cost.sum()) <= 1e-11 * abs(cost.sum())        # every shard, every instance
    assert abs(r["trips_per_step"] - s.get("iters").sum()) < 0.5
    devs = ",".join(str(d) for d in range(torch.cuda.device_count()))
    a = subprocess.run([os.path.join(HOST_BUILD, "batch_demo"), "40", "120"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    b = subprocess.run([os.path.join(HOST_BUILD, "batch_demo"), "40", "120"], cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, ILQR_DEMO_DEVICES=devs))
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr

    def lines(t):
        return [l for l in t.split("\n") if not l.startswith("NCCL version")]
    assert lines(a.stdout) == lines(b.stdout)


# ---------------------------------------------------------------------------------------------
# edge shapes and non-default parameters: GPU == kernel source on the CPU, bit for bit
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model,T,B,dt", [(abi.MODEL_ACROBOT, 1, 1, 0.02), (abi.MODEL_ACROBOT, 5, 3, 0.02),
                                          (abi.MODEL_ACROBOT, 9, 7, 0.05), (abi.MODEL_ACROBOT, 33, 5, 0.02),
                                          (abi.MODEL_DOUBLE_INTEGRATOR, 7, 3, 0.05), (abi.MODEL_DOUBLE_INTEGRATOR, 17, 2, 0.1)])
def test_edge_shapes_bit_exact(model, T, B, dt):
    """horizons shorter than a tile, not a multiple of a tile (so unaligned tiles take the lane-copy path),
    batches that do not fill a CTA"""
    n, m = abi.MODEL_DIMS[model]
    x0, u0 = make_inputs(4242, B, T, n, m, canonical_first=False)
    kw = dict(goal=[0.5, -0.5, 0.0, 0.0]) if model == abi.MODEL_DOUBLE_INTEGRATOR else {}
    s = BatchILQR(model, T=T, B=B, dt=dt, **kw)
    s.generate_trajectory(x0, u0)
    g = gpu_snap(s)
    for b in range(B):
        e = E.EmuSolver(model, dt, **kw)
        e.init(x0[b], u0[b])
        e.iterate(1000)
        r = snap(e)
        for f in g:
            assert np.array_equal(np.asarray(g[f][b]), np.asarray(r[f])), (b, f)


@pytest.mark.parametrize("head", ["rows", "thread", "warp"])
@pytest.mark.parametrize("model,T,B,dt", [(abi.MODEL_ACROBOT, 1, 1, 0.02), (abi.MODEL_ACROBOT, 2, 5, 0.02), (abi.MODEL_ACROBOT, 9, 33, 0.05),
                                          (abi.MODEL_DOUBLE_INTEGRATOR, 1, 2, 0.05), (abi.MODEL_DOUBLE_INTEGRATOR, 7, 35, 0.1)])
def test_edge_shapes_phase_engine_bit_exact(model, T, B, dt, head, monkeypatch):
    """horizons of one and two timesteps, batches that do not fill a warp / a lane group / a CTA, forced through the
    lockstep phase kernels (a batch this small normally goes straight to the persistent kernel): GPU == kernel source on
    the CPU, bit for bit, both derivative modes"""
    monkeypatch.setenv("ILQR_B200_HANDOVER", "0")
    monkeypatch.setenv("ILQR_B200_ROWS_MAX", "0" if head == "thread" else "1000000")
    monkeypatch.setenv("ILQR_B200_WARP_PRE_MAX", "1000000" if head == "warp" else "0")
    n, m = abi.MODEL_DIMS[model]
    x0, u0 = make_inputs(77, B, T, n, m, canonical_first=False)
    kw = dict(goal=[0.5, -0.5, 0.0, 0.0]) if model == abi.MODEL_DOUBLE_INTEGRATOR else {}
    for cd in (abi.COST_FD, abi.COST_ANALYTIC):
        s = BatchILQR(model, T=T, B=B, dt=dt, cost_deriv=cd, **kw)
        s.generate_trajectory(x0, u0)
        g = gpu_snap(s)
        for b in range(min(B, 6)):
            e = E.EmuSolver(model, dt, cost_deriv=cd, **kw)
            e.init(x0[b], u0[b])
            e.iterate(1000)
            r = snap(e)
            for f in g:
                assert np.array_equal(np.asarray(g[f][b]), np.asarray(r[f])), (cd, b, f)


@pytest.mark.parametrize("T", [37, 38, 199, 200])
@pytest.mark.parametrize("lanes", ["32", "16"])
def test_shared_memory_layout_odd_and_even_horizons(T, lanes, monkeypatch):
    """Closed-form cost derivatives (the smallest scratch) with odd and even horizons, both lane decompositions: the
    group's mbarrier must not share its 16-byte granule with the last gradient-norm term whatever T is (it once did
    for odd T: compute-sanitizer synccheck "Barrier error: missing init", profiles/experiments/README.md)."""
    monkeypatch.setenv("ILQR_B200_LANES", lanes)
    B = 6
    x0, u0 = make_inputs(99, B, T, 4, 1, canonical_first=False)
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, cost_deriv=abi.COST_ANALYTIC)
    s.generate_trajectory(x0, u0)
    g = gpu_snap(s)
    for b in range(B):
        e = E.EmuSolver(abi.MODEL_ACROBOT, 0.02, cost_deriv=abi.COST_ANALYTIC, lanes=int(lanes))
        e.init(x0[b], u0[b])
        e.iterate(1000)
        r = snap(e)
        for f in g:
            assert np.array_equal(np.asarray(g[f][b]), np.asarray(r[f])), (b, f)


def test_non_default_parameters_bit_exact():
    """every tunable of ilqr_params reaches the kernels (shorter alpha table, other eps / tolerances / lambda schedule)"""
    p = abi.default_params()
    p.max_iter, p.n_alpha = 9, 5
    for i, a in enumerate((1.0, 0.3, 0.09, 0.027, 0.0081)):
        p.alpha[i] = a
    p.fd_eps, p.lambda_init, p.lambda_factor, p.tol_fun, p.z_min = 5e-4, 2.0, 2.0, 1e-4, 0.05
    p.qp_armijo, p.qp_step_dec, p.qp_clamp_tol = 0.2, 0.5, 1e-3
    B, T = 6, 60
    x0, u0 = make_inputs(99, B, T, 4, 1, canonical_first=False)
    kw = dict(u_min=[-2.0], u_max=[2.5])
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, params=p, **kw)
    s.generate_trajectory(x0, u0)
    g = gpu_snap(s)
    assert (g["trips"] <= 9).all()
    for b in range(B):
        e = E.EmuSolver(abi.MODEL_ACROBOT, 0.02, params=p, **kw)
        e.init(x0[b], u0[b])
        e.iterate(1000)
        r = snap(e)
        o = O.OracleSolver(abi.MODEL_ACROBOT, 0.02, params=p, **kw)
        o.init(x0[b], u0[b])
        o.iterate(1000)
        for f in g:
            assert np.array_equal(np.asarray(g[f][b]), np.asarray(r[f])), (b, f)
        assert abs(o.cost - g["cost"][b]) <= 1e-6 * abs(o.cost)


def test_f32_config3_shape():
    """BASELINE configs[2] arithmetic at its horizon (T = 500, f32, FD fx/fu, analytic cost derivatives), reduced batch"""
    B, T = 2048, 500
    x0, u0 = make_inputs(12345, B, T, 4, 1)
    s = BatchILQR(abi.MODEL_ACROBOT, T=T, B=B, dt=0.02, dtype=abi.F32, cost_deriv=abi.COST_ANALYTIC)
    c0 = s.init_traj(x0, u0)
    s.iterate(8)
    c1 = s.get("cost")
    assert np.isfinite(c1).all() and (c1 <= c0).all() and np.median(c1 / c0) < 0.7
    assert (s.get("iters") == 8).all()
