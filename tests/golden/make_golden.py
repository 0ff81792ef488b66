#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_oracle.so).

Run in the dev container, where /root/reference is mounted and `make -C oracle ref` has built the
probe:   python tests/golden/make_golden.py
The fixtures are committed; the GPU box never needs the reference to check against them.
Inputs come from include/ilqr_synth.h (seed 12345), generated here through the
real std::mt19937_64 so the fixture also pins that header.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refharness as R  # noqa: E402

SEED = 12345
CHECKPOINTS = (1, 5, 20)


def synth(B, T, n, m, x_scale=1.0, u_scale=0.5, canonical_first=True):
    """Same draw order as ilqr_synth_fill, using libstdc++'s own generator."""
    per = n + T * m
    u = R.std_uniform(SEED, B * per).reshape(B, per)
    x0 = x_scale * u[:, :n].copy()
    u0 = (u_scale * u[:, n:]).reshape(B, T, m).copy()
    if canonical_first:
        x0[0] = 0
        u0[0] = 0
    return x0, u0


def trace_instance(s, x0, u0, checkpoints=CHECKPOINTS, max_trace=100):
    """One reference solve with snapshots: first backward pass, N-iteration checkpoints, termination."""
    out = {}
    out["init_cost"] = s.init(x0, u0)
    out["xs_init"] = s.get("xs")
    # single backward pass at lambda = 1 from the initial trajectory
    out["bw_diverge"] = s.backward_once(1.0)
    for f in ("K", "k", "dV"):
        out["bw_" + f] = s.get(f)
    out["bw_Vx0"] = s.get("Vx")[0]
    out["bw_Vxx0"] = s.get("Vxx")[0]
    out["bw_gnorm"] = s.scalar("gnorm")
    for f in ("fx", "fu", "cx", "cu", "cxx", "cxu", "cuu"):
        d = s.get(f)
        out["d_%s_first" % f] = d[0]
        out["d_%s_mid" % f] = d[s.T // 2]
        out["d_%s_last" % f] = d[s.T]
    # rollout for alpha = 0.5012 after that backward pass
    out["ro_cost"] = s.rollout_once(0.5012)
    out["ro_xsT"] = s.get("xs")[s.T]
    out["ro_us"] = s.get("us")
    # full solve with per-iteration trace
    s.init(x0, u0)
    costs, alphas, lams, dlams, gnorms, dcosts, expecteds = [], [], [], [], [], [], []
    done = 0
    while s.count("status") == 0 and done < max_trace:
        s.iterate(1)
        done += 1
        costs.append(s.cost)
        alphas.append(s.count("alpha_index"))
        lams.append(s.scalar("lam"))
        dlams.append(s.scalar("dlam"))
        gnorms.append(s.scalar("gnorm"))
        dcosts.append(s.scalar("dcost"))
        expecteds.append(s.scalar("expected"))
        if done in checkpoints:
            for f in ("K", "k", "xs", "us"):
                out["it%d_%s" % (done, f)] = s.get(f)
            out["it%d_cost" % done] = s.cost
            out["it%d_dV" % done] = s.get("dV")
    out["trace_cost"] = np.array(costs)
    out["trace_alpha_index"] = np.array(alphas, dtype=np.int32)
    out["trace_lambda"] = np.array(lams)
    out["trace_dlambda"] = np.array(dlams)
    out["trace_gnorm"] = np.array(gnorms)
    out["trace_dcost"] = np.array(dcosts)
    out["trace_expected"] = np.array(expecteds)
    out["final_cost"] = s.cost
    out["final_status"] = s.count("status")
    out["final_trips"] = s.count("loop_trips")
    out["final_accepts"] = s.count("accepts")
    out["final_rejects"] = s.count("rejects")
    out["final_rollouts"] = s.count("rollouts")
    out["final_backwards"] = s.count("backwards")
    out["final_xs"] = s.get("xs")
    out["final_us"] = s.get("us")
    out["final_K"] = s.get("K")
    out["final_k"] = s.get("k")
    return out


def pack(cases):
    flat = {}
    for name, d in cases.items():
        for k, v in d.items():
            flat["%s/%s" % (name, k)] = np.asarray(v)
    return flat


def make_solver_fixture():
    cases = {}
    # acrobot, default +-5 limits, T=200: canonical + 5 random (SURVEY §8d instances 0..5)
    x0, u0 = synth(6, 200, 4, 1)
    for b in range(6):
        s = R.RefSolver(R.ACROBOT, 0.02)
        cases["acrobot_T200_b%d" % b] = dict(trace_instance(s, x0[b], u0[b]), x0=x0[b], u0=u0[b], dt=0.02)
    # control-limited acrobot (+-1.5, acrobot.h:38), BASELINE config 4
    for b in range(3):
        s = R.RefSolver(R.ACROBOT, 0.02, u_min=[-1.5], u_max=[1.5])
        cases["acrobot_lim15_T200_b%d" % b] = dict(trace_instance(s, x0[b], u0[b]), x0=x0[b], u0=u0[b], dt=0.02,
                                                   u_min=[-1.5], u_max=[1.5])
    # the reference CLI's acrobot instance (src/run_ilqr.cpp:39-54): T=499, zeros
    s = R.RefSolver(R.ACROBOT, 0.02)
    cases["acrobot_cli_T499"] = dict(trace_instance(s, np.zeros(4), np.zeros((499, 1))), x0=np.zeros(4),
                                     u0=np.zeros((499, 1)), dt=0.02)
    # the reference CLI's integrator instance (src/run_ilqr.cpp:18-37): T=99
    goal = [1.0, 0.5, 0.0, 0.0]
    s = R.RefSolver(R.DOUBLE_INTEGRATOR, 0.02, goal=goal)
    x0d = np.array([-1.0, 0.0, 0.0, -0.2])
    cases["integrator_cli_T99"] = dict(trace_instance(s, x0d, np.zeros((99, 2))), x0=x0d, u0=np.zeros((99, 2)),
                                       dt=0.02, goal=goal)
    # double integrator with random starts (exercises m=2 partial clamping)
    xd, ud = synth(4, 60, 4, 2, x_scale=1.0, u_scale=0.5, canonical_first=False)
    for b in range(3):
        s = R.RefSolver(R.DOUBLE_INTEGRATOR, 0.05, goal=[1.0, 1.0, 0.0, 0.0])
        cases["integrator_rand_T60_b%d" % b] = dict(trace_instance(s, xd[b], ud[b]), x0=xd[b], u0=ud[b], dt=0.05,
                                                    goal=[1.0, 1.0, 0.0, 0.0])
    # warm start / MPC: iLQR::generate_trajectory(x_0) after a finished solve, src/ilqr_core.cpp:65-76.
    # lambda / dlambda carry over from the first solve (TU statics, include/ilqr.h:17-18).  The *_native values come
    # from the reference's own generate_trajectory(x_0); the checkpoints from the probe's replica of its loop, which
    # tests/test_oracle_ref.py checks against the native call bit for bit.
    rng = np.random.default_rng(77)
    for b, (T, shift) in enumerate(((120, 0.01), (200, 0.05))):
        xw, uw = rng.uniform(-1, 1, 4), 0.5 * rng.uniform(-1, 1, (T, 1))
        x1 = xw + shift
        s = R.RefSolver(R.ACROBOT, 0.02)
        s.init(xw, uw)
        s.iterate(1000)
        d = dict(x0=xw, u0=uw, dt=0.02, x0_warm=x1, first_cost=s.cost, first_lambda=s.scalar("lam"),
                 first_dlambda=s.scalar("dlam"), first_trips=s.count("loop_trips"))
        d["warm_cost"] = s.warm_start(x1)
        d["warm_xs"] = s.get("xs")
        d["warm_us"] = s.get("us")
        done = 0
        for n in (1, 3, 10):
            s.iterate(n - done)
            done = n
            for f in ("K", "k", "xs", "us"):
                d["warm_it%d_%s" % (n, f)] = s.get(f)
            d["warm_it%d_cost" % n] = s.cost
            d["warm_it%d_lambda" % n] = s.scalar("lam")
        s.iterate(1000)
        d["warm_final_cost"], d["warm_final_trips"], d["warm_final_status"] = s.cost, s.count("loop_trips"), s.count("status")
        nat = R.RefSolver(R.ACROBOT, 0.02)
        nat.solve_native(xw, uw)
        nat.warm_native(x1)
        d["warm_final_cost_native"] = nat.cost
        d["warm_final_xs_native"] = nat.get("xs")
        # "continue": generate_trajectory() called once more on the finished warm solve (:78-102)
        nat.resume_native()
        d["resume_final_cost_native"] = nat.cost
        s.resume()
        s.iterate(1000)
        d["resume_final_cost"], d["resume_trips"] = s.cost, s.count("loop_trips")
        cases["acrobot_warm_b%d" % b] = d
    # a Model subclass of the USER's own (the plugin surface, include/model.h:6-21) solved by the reference's own iLQR
    # class: the pendulum of ilqr_b200/host/pendulum_model.h (n = 2, m = 1).  The GPU runs its device twin
    # (tests/user_models.py: PENDULUM, compiled at run time) against these vectors.
    rng = np.random.default_rng(5)
    for b in range(4):
        T = 150
        xp = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5)])
        up = 0.1 * rng.standard_normal((T, 1))
        s = R.RefSolver(R.PENDULUM, 0.05, goal=[np.pi, 0, 0, 0])
        cases["pendulum_T150_b%d" % b] = dict(trace_instance(s, xp, up, checkpoints=(1, 5, 20)), x0=xp, u0=up, dt=0.05,
                                              goal=np.pi)
    # keep the file small: the big per-checkpoint arrays only for a subset
    np.savez_compressed(os.path.join(HERE, "solver_golden.npz"), **pack(cases))
    return cases


def make_csv_fixture():
    """the result files the reference's own CLI writes (src/run_ilqr.cpp, output_to_csv src/ilqr_core.cpp:414-431)"""
    import shutil
    import subprocess
    import tempfile
    exe = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "run_iLQR")
    for which in ("acrobot", "integrator"):
        with tempfile.TemporaryDirectory() as d:
            subprocess.run([exe, which], cwd=d, check=True, stdout=subprocess.DEVNULL)
            shutil.copy(os.path.join(d, "ilqr_result.csv"), os.path.join(HERE, "ref_cli_%s.csv" % which))


def make_leaf_fixture():
    rng = np.random.default_rng(2024)
    out = {}
    # boxQP / quadclamp on random problems, m = 1..3, with and without active bounds
    qp = []
    for m in (1, 2, 3):
        for trial in range(40):
            A = rng.normal(size=(m, m))
            Q = A @ A.T + (0.02 + rng.uniform(0, 2)) * np.eye(m)
            c = rng.normal(size=m) * rng.choice([0.1, 1.0, 10.0])
            width = rng.choice([0.05, 0.5, 5.0])
            center = rng.normal(size=m) * 0.3
            lo, hi = center - width, center + width
            x0 = rng.normal(size=m) * rng.choice([0.0, 0.3, 3.0])
            res, x, vf, Rf = R.boxqp(Q, c, x0, lo, hi)
            Rpad = np.zeros((3, 3))
            if res != 6:  # fully clamped (result 6): the reference never factorises and its R is uninitialised memory
                Rpad[:Rf.shape[0], :Rf.shape[1]] = Rf
            qp.append(dict(m=m, Q=np.pad(Q, ((0, 3 - m), (0, 3 - m))), c=np.pad(c, (0, 3 - m)),
                           x0=np.pad(x0, (0, 3 - m)), lo=np.pad(lo, (0, 3 - m)), hi=np.pad(hi, (0, 3 - m)),
                           result=res, x=np.pad(x, (0, 3 - m)), v_free=np.pad(vf, (0, 3 - m)), R=Rpad,
                           r_dim=Rf.shape[0]))
    for k in qp[0]:
        out["qp_" + k] = np.array([q[k] for q in qp])
    # model evaluations + FD stencils at random points
    for name, model, kw in (("acrobot", R.ACROBOT, {}), ("integrator", R.DOUBLE_INTEGRATOR, dict(goal=[1.0, 0.5, 0.0, 0.0]))):
        s = R.RefSolver(model, 0.02, **kw)
        X = rng.uniform(-3, 3, size=(16, s.n))
        U = rng.uniform(-5, 5, size=(16, s.m))
        out[name + "_X"], out[name + "_U"] = X, U
        out[name + "_dyn"] = np.array([s.dynamics(x, u) for x, u in zip(X, U)])
        out[name + "_step"] = np.array([s.integrate(x, u, 0.02) for x, u in zip(X, U)])
        out[name + "_cost"] = np.array([s.model_cost(x, u) for x, u in zip(X, U)])
        out[name + "_final"] = np.array([s.final_cost(x) for x in X])
        for w in range(8):
            out["%s_fd%d" % (name, w)] = np.array([s.fd(w, x, u) for x, u in zip(X, U)])
    out["std_uniform_seed12345"] = R.std_uniform(SEED, 1000)
    np.savez_compressed(os.path.join(HERE, "leaf_golden.npz"), **out)


if __name__ == "__main__":
    assert R.available(), "build oracle/_ref first: make -C oracle ref"
    make_leaf_fixture()
    make_csv_fixture()
    cases = make_solver_fixture()
    for name, d in cases.items():
        if "init_cost" in d:
            print("%-28s init %.6f final %.12g trips %d status %s" % (
                name, d["init_cost"], d["final_cost"], d["final_trips"], R.STATUS[d["final_status"]]))
        else:
            print("%-28s first %.12g warm %.12g -> %.12g (native %.12g) resume %.12g (native %.12g)" % (
                name, d["first_cost"], d["warm_cost"], d["warm_final_cost"], d["warm_final_cost_native"],
                d["resume_final_cost"], d["resume_final_cost_native"]))
    for f in ("solver_golden.npz", "leaf_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
