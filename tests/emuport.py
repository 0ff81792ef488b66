"""ctypes view of tests/_build/libilqr_emu.so — the kernel core (ilqr_b200/csrc/ilqr_core.cuh)
compiled for the CPU with the warp phases run lane by lane (tests/emu/ilqr_emu.cpp).

Test infrastructure only: it lets the CPU suite check the kernel source's control flow and
arithmetic order bit-for-bit against the oracle.  The product never loads it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ilqr_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "tests", "_build", "libilqr_emu.so")            # same sin/cos as the CUDA build
LIB_PATH_LIBM = os.path.join(ROOT, "tests", "_build", "libilqr_emu_libm.so")  # platform libm sin/cos (= the oracle's)
SRC = os.path.join(ROOT, "tests", "emu", "ilqr_emu.cpp")
DEPS = [SRC] + [os.path.join(ROOT, "ilqr_b200", "csrc", f) for f in
                ("ilqr_core.cuh", "ilqr_phases.cuh", "boxqp.cuh", "models.cuh", "trig.cuh", "params.h")] + [os.path.join(ROOT, "include", "ilqr_b200.h")]
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_libs = {}


LIB_PATH_NOISE = os.path.join(ROOT, "tests", "_build", "libilqr_emu_noise.so")  # libm sin/cos with 1-ulp noise (experiments)


def _path(libm):
    return LIB_PATH_NOISE if libm == "noise" else (LIB_PATH_LIBM if libm else LIB_PATH)


def build(libm=False):
    path = _path(libm)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    defs = ["-DILQR_TRIG_LIBM", "-DILQR_TRIG_NOISE"] if libm == "noise" else (["-DILQR_TRIG_LIBM"] if libm else [])
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-Wall",
                           "-Wno-unknown-pragmas", "-Wno-maybe-uninitialized", "-I" + os.path.join(ROOT, "include")] +
                          defs + ["-o", path, SRC])


def lib(libm=False):
    if libm not in _libs:
        path = _path(libm)
        if (not os.path.exists(path)) or any(os.path.getmtime(d) > os.path.getmtime(path) for d in DEPS):
            build(libm)
        L = C.CDLL(path)
        vp = C.c_void_p
        L.emu_new.restype = vp
        L.emu_new.argtypes = [C.POINTER(abi.Desc)]
        L.emu_free.argtypes = [vp]
        L.emu_dims.argtypes = [vp, _ip, _ip]
        L.emu_init.restype = C.c_double
        L.emu_init.argtypes = [vp, _dp, _dp, C.c_int]
        L.emu_warm_start.restype = C.c_double
        L.emu_warm_start.argtypes = [vp, _dp]
        L.emu_iterate.argtypes = [vp, C.c_int]
        L.emu_backward_once.argtypes = [vp, C.c_double]
        L.emu_rollout_once.restype = C.c_double
        L.emu_rollout_once.argtypes = [vp, C.c_double]
        L.emu_get.argtypes = [vp, C.c_int, _dp]
        L.emu_scalar.restype = C.c_double
        L.emu_scalar.argtypes = [vp, C.c_int]
        L.emu_int.restype = C.c_long
        L.emu_int.argtypes = [vp, C.c_int]
        L.emu_boxqp.argtypes = [C.POINTER(abi.Params), C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _ip]
        _libs[libm] = L
    return _libs[libm]


def _p(a):
    return a.ctypes.data_as(_dp)


def _arr(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class EmuSolver:
    """Same surface as oracleport.OracleSolver (Vx/Vxx: timestep 0 only)."""

    def __init__(self, model=abi.MODEL_ACROBOT, dt=0.02, goal=None, u_min=None, u_max=None,
                 cost_deriv=abi.COST_FD, params=None, dtype=abi.F64, libm=False, lanes=32, flags=0):
        self.L = lib(libm)
        self.L.emu_set_lanes(int(lanes))  # 32: one trajectory per warp; 16: two per warp; 1: the phase engine (ilqr_phases.cuh)
        self.desc = abi.make_desc(model=model, dt=dt, goal=goal, u_min=u_min, u_max=u_max, cost_deriv=cost_deriv,
                                  params=params, dtype=dtype, flags=flags)
        self.h = self.L.emu_new(C.byref(self.desc))
        assert self.h, "emu_new failed"
        n, m = C.c_int(), C.c_int()
        self.L.emu_dims(self.h, C.byref(n), C.byref(m))
        self.n, self.m, self.dt = n.value, m.value, dt
        self.T = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.emu_free(self.h)
            self.h = None

    def init(self, x0, u0):
        x0, u0 = _arr(x0), _arr(u0).reshape(-1, self.m)
        self.T = u0.shape[0]
        return self.L.emu_init(self.h, _p(x0), _p(u0), self.T)

    def warm_start(self, x0):
        return self.L.emu_warm_start(self.h, _p(_arr(x0)))

    def iterate(self, n):
        return self.L.emu_iterate(self.h, n)

    def resume(self):
        self.L.emu_resume.argtypes = [C.c_void_p]
        self.L.emu_resume(self.h)

    def backward_once(self, lam=1.0):
        return self.L.emu_backward_once(self.h, lam)

    def rollout_once(self, alpha):
        return self.L.emu_rollout_once(self.h, alpha)

    def get(self, name):
        T, n, m = self.T, self.n, self.m
        shapes = dict(xs=(T + 1, n), us=(T, m), K=(T, m, n), k=(T, m), cost=(1,), dV=(2,), Vx0=(n,), Vxx0=(n, n))
        ids = dict(xs=0, us=1, K=2, k=3, cost=4, dV=5, Vx0=6, Vxx0=7)
        out = np.empty(shapes[name], dtype=np.float64)
        cnt = self.L.emu_get(self.h, ids[name], _p(out))
        assert cnt == out.size, (name, cnt, out.size)
        return out

    @property
    def cost(self):
        return float(self.get("cost")[0])

    def scalar(self, name):
        return self.L.emu_scalar(self.h, dict(lam=0, dlam=1, gnorm=2, dcost=3, expected=4, alpha=5, new_cost=6)[name])

    def count(self, name):
        return int(self.L.emu_int(self.h, dict(iter=0, loop_trips=1, status=2, alpha_index=3, accepts=4, rejects=5,
                                              rollouts=6, backwards=7, derivs=8, T=9, diverge=10)[name]))


def boxqp(Q, c, x0, lo, hi, params=None, generic=False):
    Q, c, x0, lo, hi = map(_arr, (Q, c, x0, lo, hi))
    m = c.size
    x = np.empty(m)
    vf = np.zeros(m, dtype=np.int32)
    R = np.zeros(m * m)
    rd = C.c_int()
    res = lib().emu_boxqp(C.byref(params) if params is not None else None, m, int(generic), _p(Q), _p(c), _p(x0),
                          _p(lo), _p(hi), _p(x), vf.ctypes.data_as(_ip), _p(R), C.byref(rd))
    r = rd.value
    return res, x, vf, R[:r * r].reshape(r, r)
