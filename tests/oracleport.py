"""ctypes view of oracle/_build/liboracle_ilqr.so — the plain-C restatement (CPU oracle).

Test infrastructure only.  Mirrors tests/refharness.py so tests can swap one for the other.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ilqr_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_build", "liboracle_ilqr.so")
FIELDS = dict(xs=0, us=1, K=2, k=3, cost=4, dV=5, Vx=6, Vxx=7, fx=8, fu=9, cx=10, cu=11, cxx=12, cxu=13, cuu=14)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.orc_new.restype = vp
        L.orc_new.argtypes = [C.POINTER(abi.Desc)]
        L.orc_free.argtypes = [vp]
        L.orc_dims.argtypes = [vp, _ip, _ip]
        L.orc_init.restype = C.c_double
        L.orc_init.argtypes = [vp, _dp, _dp, C.c_int]
        L.orc_warm_start.restype = C.c_double
        L.orc_warm_start.argtypes = [vp, _dp]
        L.orc_iterate.argtypes = [vp, C.c_int]
        L.orc_backward_once.argtypes = [vp, C.c_double, C.c_int]
        L.orc_rollout_once.restype = C.c_double
        L.orc_rollout_once.argtypes = [vp, C.c_double]
        L.orc_get.argtypes = [vp, C.c_int, _dp]
        L.orc_scalar.restype = C.c_double
        L.orc_scalar.argtypes = [vp, C.c_int]
        L.orc_int.restype = C.c_long
        L.orc_int.argtypes = [vp, C.c_int]
        L.orc_dynamics.argtypes = [vp, _dp, _dp, _dp]
        L.orc_integrate.argtypes = [vp, _dp, _dp, C.c_double, _dp]
        L.orc_cost.restype = C.c_double
        L.orc_cost.argtypes = [vp, _dp, _dp]
        L.orc_final_cost.restype = C.c_double
        L.orc_final_cost.argtypes = [vp, _dp]
        L.orc_fd.argtypes = [vp, C.c_int, _dp, _dp, C.c_double, _dp]
        L.orc_boxqp.argtypes = [C.POINTER(abi.Params), C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _ip]
        L.orc_quadclamp.argtypes = [C.POINTER(abi.Params), C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip]
        L.orc_quadcost.restype = C.c_double
        L.orc_quadcost.argtypes = [C.c_int, _dp, _dp, _dp]
        L.orc_solve_range.restype = C.c_long
        L.orc_solve_range.argtypes = [C.POINTER(abi.Desc), C.c_long, C.c_long, _dp, _dp, C.c_int, _dp, _ip, _ip, _lp, _lp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _arr(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class OracleSolver:
    """Same surface as refharness.RefSolver, backed by the C restatement."""

    def __init__(self, model=abi.MODEL_ACROBOT, dt=0.02, goal=None, u_min=None, u_max=None,
                 cost_deriv=abi.COST_FD, params=None, flags=0):
        self.desc = abi.make_desc(model=model, dt=dt, goal=goal, u_min=u_min, u_max=u_max, cost_deriv=cost_deriv,
                                  params=params, flags=flags)
        self.h = lib().orc_new(C.byref(self.desc))
        assert self.h, "orc_new failed"
        n, m = C.c_int(), C.c_int()
        lib().orc_dims(self.h, C.byref(n), C.byref(m))
        self.n, self.m, self.dt = n.value, m.value, dt
        self.T = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_free(self.h)
            self.h = None

    def init(self, x0, u0):
        x0, u0 = _arr(x0), _arr(u0).reshape(-1, self.m)
        self.T = u0.shape[0]
        return lib().orc_init(self.h, _p(x0), _p(u0), self.T)

    def warm_start(self, x0):
        return lib().orc_warm_start(self.h, _p(_arr(x0)))

    def resume(self):
        lib().orc_resume.argtypes = [C.c_void_p]
        lib().orc_resume(self.h)

    def iterate(self, n):
        return lib().orc_iterate(self.h, n)

    def backward_once(self, lam=1.0, recompute=True):
        return lib().orc_backward_once(self.h, lam, int(recompute))

    def rollout_once(self, alpha):
        return lib().orc_rollout_once(self.h, alpha)

    def get(self, name):
        T, n, m = self.T, self.n, self.m
        shapes = dict(xs=(T + 1, n), us=(T, m), K=(T, m, n), k=(T, m), cost=(1,), dV=(2,), Vx=(T + 1, n),
                      Vxx=(T + 1, n, n), fx=(T + 1, n, n), fu=(T + 1, n, m), cx=(T + 1, n), cu=(T + 1, m),
                      cxx=(T + 1, n, n), cxu=(T + 1, n, m), cuu=(T + 1, m, m))
        out = np.empty(shapes[name], dtype=np.float64)
        cnt = lib().orc_get(self.h, FIELDS[name], _p(out))
        assert cnt == out.size, (name, cnt, out.size)
        return out

    @property
    def cost(self):
        return float(self.get("cost")[0])

    def scalar(self, name):
        return lib().orc_scalar(self.h, dict(lam=0, dlam=1, gnorm=2, dcost=3, expected=4, alpha=5, new_cost=6)[name])

    def count(self, name):
        return int(lib().orc_int(self.h, dict(iter=0, loop_trips=1, status=2, alpha_index=3, accepts=4, rejects=5,
                                              rollouts=6, backwards=7, derivs=8, T=9, diverge=10)[name]))

    def dynamics(self, x, u):
        out = np.empty(self.n)
        lib().orc_dynamics(self.h, _p(_arr(x)), _p(_arr(u)), _p(out))
        return out

    def integrate(self, x, u, dt):
        out = np.empty(self.n)
        lib().orc_integrate(self.h, _p(_arr(x)), _p(_arr(u)), dt, _p(out))
        return out

    def model_cost(self, x, u):
        return lib().orc_cost(self.h, _p(_arr(x)), _p(_arr(u)))

    def final_cost(self, x):
        return lib().orc_final_cost(self.h, _p(_arr(x)))

    def fd(self, which, x, u, dt=None):
        n, m = self.n, self.m
        shape = {0: (n, n), 1: (n, m), 2: (n,), 3: (m,), 4: (n,), 5: (n, n), 6: (m, m), 7: (n, n)}[which]
        out = np.empty(shape)
        lib().orc_fd(self.h, which, _p(_arr(x)), _p(_arr(u)), self.dt if dt is None else dt, _p(out))
        return out


def boxqp(Q, c, x0, lo, hi, params=None):
    Q, c, x0, lo, hi = map(_arr, (Q, c, x0, lo, hi))
    m = c.size
    x = np.empty(m)
    vf = np.zeros(m, dtype=np.int32)
    R = np.zeros(m * m)
    rd = C.c_int()
    res = lib().orc_boxqp(C.byref(params) if params is not None else None, m, _p(Q), _p(c), _p(x0), _p(lo), _p(hi),
                          _p(x), vf.ctypes.data_as(_ip), _p(R), C.byref(rd))
    r = rd.value
    return res, x, vf, R[:r * r].reshape(r, r)


def quadclamp(x0, d, Q, c, lo, hi, params=None):
    x0, d, Q, c, lo, hi = map(_arr, (x0, d, Q, c, lo, hi))
    m = c.size
    x = np.empty(m)
    v, ns = C.c_double(), C.c_int()
    failed = lib().orc_quadclamp(C.byref(params) if params is not None else None, m, _p(x0), _p(d), _p(Q), _p(c),
                                 _p(lo), _p(hi), _p(x), C.byref(v), C.byref(ns))
    return bool(failed), x, v.value, ns.value


def quadcost(Q, c, x):
    Q, c, x = map(_arr, (Q, c, x))
    return lib().orc_quadcost(c.size, _p(Q), _p(c), _p(x))


def solve_range(desc, x0, u0, b0, b1, max_trips=-1):
    """Sequential CPU solves of instances [b0, b1); returns dict of per-instance results."""
    x0, u0 = _arr(x0), _arr(u0)
    nb = b1 - b0
    cost = np.empty(nb)
    iters = np.zeros(nb, dtype=np.int32)
    status = np.zeros(nb, dtype=np.int32)
    acc = np.zeros(nb, dtype=np.int64)
    rej = np.zeros(nb, dtype=np.int64)
    total = lib().orc_solve_range(C.byref(desc), b0, b1, _p(x0), _p(u0), max_trips, _p(cost),
                                  iters.ctypes.data_as(_ip), status.ctypes.data_as(_ip),
                                  acc.ctypes.data_as(_lp), rej.ctypes.data_as(_lp))
    return dict(total_trips=total, cost=cost, iters=iters, status=status, n_accept=acc, n_reject=rej)
