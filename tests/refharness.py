"""ctypes view of oracle/_ref/libref_oracle.so — the UNMODIFIED reference behind a probe.

Test infrastructure only (see oracle/ref_harness.cpp).  The library exists wherever
`make -C oracle ref` has run; it is built in the dev container (where /root/reference is
mounted) and travels to the GPU box as a prebuilt file.  Nothing here reads /root/reference.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so")

ACROBOT, DOUBLE_INTEGRATOR = 0, 1
PENDULUM = 2  # the user-model example (ilqr_b200/host/pendulum_model.h): goal[0] = goal angle
STATUS = {0: "RUNNING", 1: "GRAD", 2: "TOLFUN", 3: "LAMBDA_MAX", 4: "MAXITER"}
FIELDS = dict(xs=0, us=1, K=2, k=3, cost=4, dV=5, Vx=6, Vxx=7, fx=8, fu=9, cx=10, cu=11, cxx=12, cxu=13, cuu=14)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_new.restype = C.c_void_p
        L.ref_new.argtypes = [C.c_int, _dp, C.c_double, _dp, _dp]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_dims.argtypes = [C.c_void_p, _ip, _ip]
        L.ref_init.restype = C.c_double
        L.ref_init.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        L.ref_iterate.argtypes = [C.c_void_p, C.c_int]
        L.ref_backward_once.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.ref_rollout_once.restype = C.c_double
        L.ref_rollout_once.argtypes = [C.c_void_p, C.c_double]
        L.ref_solve_native.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        L.ref_get.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_scalar.restype = C.c_double
        L.ref_scalar.argtypes = [C.c_void_p, C.c_int]
        L.ref_int.restype = C.c_long
        L.ref_int.argtypes = [C.c_void_p, C.c_int]
        L.ref_dynamics.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.ref_integrate.argtypes = [C.c_void_p, _dp, _dp, C.c_double, _dp]
        L.ref_cost.restype = C.c_double
        L.ref_cost.argtypes = [C.c_void_p, _dp, _dp]
        L.ref_final_cost.restype = C.c_double
        L.ref_final_cost.argtypes = [C.c_void_p, _dp]
        L.ref_boxqp.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _ip]
        L.ref_quadclamp.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip]
        L.ref_quadcost.restype = C.c_double
        L.ref_quadcost.argtypes = [C.c_int, _dp, _dp, _dp]
        L.ref_fd.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_double, _dp]
        L.ref_std_uniform.argtypes = [C.c_ulonglong, C.c_int, _dp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _arr(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class RefSolver:
    """One reference iLQR instance (`new iLQR(model, dt)`, src/run_ilqr.cpp:22-27,41-45)."""

    def __init__(self, model=ACROBOT, dt=0.02, goal=None, u_min=None, u_max=None):
        L = lib()
        g = _arr(goal) if goal is not None else None
        lo = _arr(u_min) if u_min is not None else None
        hi = _arr(u_max) if u_max is not None else None
        self.h = L.ref_new(model, _p(g) if g is not None else None, dt, _p(lo) if lo is not None else None,
                           _p(hi) if hi is not None else None)
        n, m = C.c_int(), C.c_int()
        L.ref_dims(self.h, C.byref(n), C.byref(m))
        self.n, self.m, self.dt = n.value, m.value, dt
        self.T = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_free(self.h)
            self.h = None

    def init(self, x0, u0):
        x0, u0 = _arr(x0), _arr(u0).reshape(-1, self.m)
        self.T = u0.shape[0]
        return lib().ref_init(self.h, _p(x0), _p(u0), self.T)

    def iterate(self, n):
        return lib().ref_iterate(self.h, n)

    def backward_once(self, lam=1.0, recompute=True):
        return lib().ref_backward_once(self.h, lam, int(recompute))

    def rollout_once(self, alpha):
        return lib().ref_rollout_once(self.h, alpha)

    def solve_native(self, x0, u0):
        x0, u0 = _arr(x0), _arr(u0).reshape(-1, self.m)
        self.T = u0.shape[0]
        cwd = os.getcwd()
        import tempfile
        with tempfile.TemporaryDirectory() as d:  # generate_trajectory drops ilqr_result.csv in cwd
            os.chdir(d)
            try:
                lib().ref_solve_native(self.h, _p(x0), _p(u0), self.T)
            finally:
                os.chdir(cwd)

    def _in_tmp(self, fn):
        cwd = os.getcwd()
        import tempfile
        with tempfile.TemporaryDirectory() as d:  # generate_trajectory drops ilqr_result.csv in cwd
            os.chdir(d)
            try:
                fn()
            finally:
                os.chdir(cwd)

    def warm_start(self, x0):
        """replica of generate_trajectory(x_0) up to the loop (src/ilqr_core.cpp:65-76); continue with iterate()"""
        L = lib()
        L.ref_warm_start.restype = C.c_double
        L.ref_warm_start.argtypes = [C.c_void_p, _dp]
        return L.ref_warm_start(self.h, _p(_arr(x0)))

    def warm_native(self, x0):
        """the reference's own generate_trajectory(x_0), start to finish"""
        L = lib()
        L.ref_warm_native.argtypes = [C.c_void_p, _dp]
        x0 = _arr(x0)
        self._in_tmp(lambda: L.ref_warm_native(self.h, _p(x0)))

    def resume(self):
        L = lib()
        L.ref_resume.argtypes = [C.c_void_p]
        L.ref_resume(self.h)

    def resume_native(self):
        L = lib()
        L.ref_resume_native.argtypes = [C.c_void_p]
        self._in_tmp(lambda: L.ref_resume_native(self.h))

    def get(self, name):
        T, n, m = self.T, self.n, self.m
        shapes = dict(xs=(T + 1, n), us=(T, m), K=(T, m, n), k=(T, m), cost=(1,), dV=(2,), Vx=(T + 1, n),
                      Vxx=(T + 1, n, n), fx=(T + 1, n, n), fu=(T + 1, n, m), cx=(T + 1, n), cu=(T + 1, m),
                      cxx=(T + 1, n, n), cxu=(T + 1, n, m), cuu=(T + 1, m, m))
        out = np.empty(shapes[name], dtype=np.float64)
        cnt = lib().ref_get(self.h, FIELDS[name], _p(out))
        assert cnt == out.size, (name, cnt, out.size)
        return out

    @property
    def cost(self):
        return float(self.get("cost")[0])

    def scalar(self, name):
        return lib().ref_scalar(self.h, dict(lam=0, dlam=1, gnorm=2, dcost=3, expected=4, alpha=5, new_cost=6)[name])

    def count(self, name):
        return int(lib().ref_int(self.h, dict(iter=0, loop_trips=1, status=2, alpha_index=3, accepts=4, rejects=5,
                                              rollouts=6, backwards=7, derivs=8, T=9)[name]))

    # leaf probes
    def dynamics(self, x, u):
        out = np.empty(self.n)
        lib().ref_dynamics(self.h, _p(_arr(x)), _p(_arr(u)), _p(out))
        return out

    def integrate(self, x, u, dt):
        out = np.empty(self.n)
        lib().ref_integrate(self.h, _p(_arr(x)), _p(_arr(u)), dt, _p(out))
        return out

    def model_cost(self, x, u):
        return lib().ref_cost(self.h, _p(_arr(x)), _p(_arr(u)))

    def final_cost(self, x):
        return lib().ref_final_cost(self.h, _p(_arr(x)))

    def fd(self, which, x, u, dt=None):
        n, m = self.n, self.m
        shape = {0: (n, n), 1: (n, m), 2: (n,), 3: (m,), 4: (n,), 5: (n, n), 6: (m, m), 7: (n, n)}[which]
        out = np.empty(shape)
        lib().ref_fd(self.h, which, _p(_arr(x)), _p(_arr(u)), self.dt if dt is None else dt, _p(out))
        return out


def boxqp(Q, c, x0, lo, hi):
    Q, c, x0, lo, hi = map(_arr, (Q, c, x0, lo, hi))
    m = c.size
    x = np.empty(m)
    vf = np.zeros(m, dtype=np.int32)
    R = np.zeros(m * m)
    rd = C.c_int()
    res = lib().ref_boxqp(m, _p(Q), _p(c), _p(x0), _p(lo), _p(hi), _p(x), vf.ctypes.data_as(_ip), _p(R), C.byref(rd))
    r = rd.value
    return res, x, vf, R[:r * r].reshape(r, r)


def quadclamp(x0, d, Q, c, lo, hi):
    x0, d, Q, c, lo, hi = map(_arr, (x0, d, Q, c, lo, hi))
    m = c.size
    x = np.empty(m)
    v, ns = C.c_double(), C.c_int()
    failed = lib().ref_quadclamp(m, _p(x0), _p(d), _p(Q), _p(c), _p(lo), _p(hi), _p(x), C.byref(v), C.byref(ns))
    return bool(failed), x, v.value, ns.value


def quadcost(Q, c, x):
    Q, c, x = map(_arr, (Q, c, x))
    return lib().ref_quadcost(c.size, _p(Q), _p(c), _p(x))


def std_uniform(seed, count):
    out = np.empty(count)
    lib().ref_std_uniform(seed, count, _p(out))
    return out
