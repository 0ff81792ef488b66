"""The run-time compiled model path (include/ilqr_b200.h: ilqr_register_model / ilqr_compile_model): the GPU side of
the reference's `class Model` plugin surface.  NVRTC needs no GPU, so registration and compilation are CPU tests; the
GPU tests run the compiled kernels."""
import os
import subprocess

import numpy as np
import pytest

from ilqr_b200 import abi
from ilqr_b200.solver import BatchILQR, make_inputs

import user_models as U


def test_register_and_compile_without_a_gpu():
    mid = abi.register_model("UserDoubleIntegrator", U.DOUBLE_INTEGRATOR_CLONE, 4, 2, [-0.5, -0.5], [0.5, 0.5])
    assert mid >= abi.MODEL_USER_BASE
    assert abi.model_dims(mid) == (4, 2)
    rc, log = abi.compile_model(mid, abi.F64, abi.COST_ANALYTIC)
    assert rc == 0, log
    pid = abi.register_model("Pendulum", U.PENDULUM, 2, 1, [-2.0], [2.0])
    assert pid == mid + 1 and abi.model_dims(pid) == (2, 1)
    rc, log = abi.compile_model(pid, abi.F64, abi.COST_FD)
    assert rc == 0, log


def test_a_model_that_does_not_compile_reports_the_compiler_log():
    bid = abi.register_model("Broken", U.BROKEN, 4, 1, [-1.0], [1.0])
    rc, log = abi.compile_model(bid)
    assert rc != 0 and "kConfigVars" in log
    wid = abi.register_model("UserDoubleIntegrator", U.DOUBLE_INTEGRATOR_CLONE, 4, 1, [-1.0], [1.0])  # wrong m
    rc, log = abi.compile_model(wid)
    assert rc != 0 and "differ from what ilqr_register_model was told" in log
    with pytest.raises(ValueError):
        abi.register_model("X", "struct X {};", 99, 1, [-1.0], [1.0])


FIELDS = ("xs", "us", "K", "k", "cost", "lambda", "iters", "status", "alpha_index", "dV", "Vx0", "Vxx0")


@pytest.mark.gpu
@pytest.mark.parametrize("cost_deriv", [abi.COST_ANALYTIC, abi.COST_FD])
def test_user_clone_of_a_builtin_model_is_bit_exact(cost_deriv):
    """The double integrator written out again as user source and compiled at run time == the built-in twin."""
    mid = abi.register_model("UserDoubleIntegrator", U.DOUBLE_INTEGRATOR_CLONE, 4, 2, [-0.5, -0.5], [0.5, 0.5])
    B, T, goal = 12, 99, [1.0, 0.5, 0.0, 0.0]
    x0, u0 = make_inputs(31, B, T, 4, 2, canonical_first=False)
    ref = BatchILQR(abi.MODEL_DOUBLE_INTEGRATOR, T=T, B=B, dt=0.02, cost_deriv=cost_deriv, goal=goal)
    usr = BatchILQR(mid, T=T, B=B, dt=0.02, cost_deriv=cost_deriv, model_params=goal)
    for s in (ref, usr):
        s.set_initial(x0, u0)
        s.iterate(3)
    for f in FIELDS:
        assert np.array_equal(ref.get(f), usr.get(f)), f
    for s in (ref, usr):
        s.warm_start(x0 + 0.01)
        s.solve()
    for f in FIELDS:
        assert np.array_equal(ref.get(f), usr.get(f)), f
    assert (usr.get("status") != abi.RUNNING).all()


@pytest.mark.gpu
def test_user_pendulum_swings_up_and_closed_form_matches_finite_differences():
    """A model the library does not ship (n = 2, m = 1: other matrix sizes in every phase)."""
    pid = abi.register_model("Pendulum", U.PENDULUM, 2, 1, [-2.0], [2.0])
    B, T = 64, 150
    rng = np.random.default_rng(5)
    x0 = np.stack([rng.uniform(-0.5, 0.5, B), rng.uniform(-0.5, 0.5, B)], axis=1)
    u0 = 0.1 * rng.standard_normal((B, T, 1))
    sols = {}
    for cd in (abi.COST_ANALYTIC, abi.COST_FD):
        s = BatchILQR(pid, T=T, B=B, dt=0.05, cost_deriv=cd, model_params=[np.pi])
        c0 = s.init_traj(x0, u0).copy()
        s.iterate(1)
        sols[cd] = dict(c0=c0, K1=s.get("K").copy(), k1=s.get("k").copy(), c1=s.get("cost").copy())
        s.solve()
        sols[cd].update(cost=s.get("cost").copy(), xs=s.get("xs").copy(), status=s.get("status").copy(), us=s.get("us").copy())
    a, f = sols[abi.COST_ANALYTIC], sols[abi.COST_FD]
    assert np.isfinite(a["cost"]).all() and (a["status"] != abi.RUNNING).all()
    assert (a["cost"] < a["c0"]).all() and np.median(a["cost"] / a["c0"]) < 0.1   # the solve did its job
    assert np.median(np.abs(a["xs"][:, -1, 0] - np.pi)) < 0.2       # most instances end near the upright position
    assert np.abs(a["us"]).max() > 1.0                              # and had to use torque to get there
    # the cost is quadratic, so central differences are exact up to rounding: gains after one backward pass agree
    assert np.allclose(a["K1"], f["K1"], rtol=1e-5, atol=1e-7) and np.allclose(a["k1"], f["k1"], rtol=1e-5, atol=1e-7)
    assert np.allclose(a["c1"], f["c1"], rtol=1e-6)


@pytest.mark.gpu
def test_user_pendulum_against_the_reference_solving_the_host_model(golden_solver):
    """SURVEY §8 (f2), oracle side: `class Pendulum : public Model` (ilqr_b200/host/pendulum_model.h, the user's HOST
    object) was solved by the UNMODIFIED reference's own iLQR class (oracle/ref_harness.cpp takes any Model*;
    tests/golden/make_golden.py wrote the vectors); its device twin, compiled at run time, must reproduce the reference's
    K, k, xs, us and cost after 1 and 5 trips to 1e-6 on every instance, and the terminal cost."""
    g = golden_solver
    cases = ["pendulum_T150_b%d" % b for b in range(4)]
    x0 = np.stack([g[c + "/x0"] for c in cases])
    u0 = np.stack([g[c + "/u0"] for c in cases])
    pid = abi.register_model("Pendulum", U.PENDULUM, 2, 1, [-2.0], [2.0])
    s = BatchILQR(pid, T=150, B=len(cases), dt=0.05, cost_deriv=abi.COST_FD, model_params=[float(g[cases[0] + "/goal"])])

    def close(a, b, rtol=1e-6, atol=1e-9):
        a, b = np.asarray(a, float).reshape(len(cases), -1), np.asarray(b, float).reshape(len(cases), -1)
        err = np.maximum(np.abs(a - b).max(1) - atol, 0) / np.maximum(np.abs(b).max(1), 1e-300)
        assert (err <= rtol).all(), err
    close(s.init_traj(x0, u0), [g[c + "/init_cost"] for c in cases], 1e-12, 0)
    s.backward_once(1.0)
    for f, name in (("K", "bw_K"), ("k", "bw_k"), ("dV", "bw_dV"), ("Vx0", "bw_Vx0"), ("Vxx0", "bw_Vxx0")):
        close(s.get(f), np.stack([g[c + "/" + name] for c in cases]), 1e-7)
    s.set_initial(x0, u0)
    done = 0
    for n in (1, 5):
        s.iterate(n - done)
        done = n
        for f in ("K", "k", "xs", "us"):
            close(s.get(f), np.stack([g["%s/it%d_%s" % (c, n, f)] for c in cases]))
        close(s.get("cost"), [g["%s/it%d_cost" % (c, n)] for c in cases])
    s.solve()
    close(s.get("cost"), [g[c + "/final_cost"] for c in cases], 1e-5)
    assert np.abs(s.get("iters") - np.array([int(g[c + "/final_trips"]) for c in cases])).max() <= 12


HOST_DEMO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ilqr_b200", "host", "_build",
                         "user_model_demo")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(HOST_DEMO), reason="host binaries not built")
def test_cpp_model_subclass_with_registered_twin(tmp_path):
    """`class Pendulum : public Model` (host virtuals) + its device twin registered from C++: `new iLQR(new Pendulum, dt)`
    verifies the twin against the host object on the GPU, solves, and agrees with the same problem driven from Python."""
    twin = tmp_path / "pendulum_twin.cu"
    twin.write_text(U.PENDULUM)
    T = 150
    out = subprocess.run([HOST_DEMO, str(twin), str(T)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    w = out.stdout.split()
    c0, cost, iters = float(w[w.index("initial") + 2]), float(w[w.index("final") + 2]), int(w[w.index("iterations") + 1])
    pid = abi.register_model("Pendulum", U.PENDULUM, 2, 1, [-2.0], [2.0])
    s = BatchILQR(pid, T=T, B=1, dt=0.05, cost_deriv=abi.COST_ANALYTIC, model_params=[3.141592653589793])
    x0 = np.array([[0.2, -0.1]])
    u0 = (0.1 * np.sin(0.3 * np.arange(T))).reshape(1, T, 1)
    assert abs(s.init_traj(x0, u0)[0] - c0) <= 1e-9 * abs(c0)
    s.solve()
    assert abs(s.get("cost")[0] - cost) <= 1e-9 * abs(cost) and int(s.get("iters")[0]) == iters
    # a twin that does not match the host object is refused
    bad = tmp_path / "bad_twin.cu"
    bad.write_text(U.PENDULUM.replace("S(9.81)", "S(9.0)"))
    out = subprocess.run([HOST_DEMO, str(bad), str(T)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 1 and "disagrees with the host Model object" in out.stderr
