"""Host-side mirror of the reference's solver interface for a BATCH of problem instances.

`BatchILQR` plays the role of `class iLQR` (reference include/ilqr.h:28-107) for B independent
instances at once and forwards every call through the C ABI of libilqr_b200.so
(include/ilqr_b200.h); method names follow the reference: `init_traj`, `generate_trajectory`
(fresh / warm-start / continue overloads, include/ilqr.h:49-54), plus read access to the results
the reference keeps private (xs, us, K, k, cost_s ...).  There is no CPU implementation here:
if the CUDA library is missing or no GPU is present the constructor raises.
"""
import ctypes as C

import numpy as np

from . import abi


class ILQRError(RuntimeError):
    pass


class BatchILQR:
    """`new iLQR(model, dt)` for B instances (include/ilqr.h:30-44)."""

    def __init__(self, model=abi.MODEL_ACROBOT, T=200, B=1, dt=0.02, dtype=abi.F64, cost_deriv=abi.COST_FD,
                 device=0, u_min=None, u_max=None, goal=None, params=None, model_params=None, flags=0):
        self.lib = abi.load()
        self.desc = abi.make_desc(model=model, T=T, B=B, dt=dt, dtype=dtype, cost_deriv=cost_deriv, device=device,
                                  u_min=u_min, u_max=u_max, goal=goal, params=params, flags=flags)
        if model_params is not None:  # ilqr_desc.model_params[0..15]: what a registered model's functions see as `mp`
            for i, v in enumerate(model_params):
                self.desc.model_params[i] = float(v)
        self.n, self.m = abi.model_dims(model)
        self.T, self.B = int(T), int(B)
        self.np_dtype = np.float32 if dtype == abi.F32 else np.float64
        self.h = C.c_void_p()
        rc = self.lib.ilqr_create(C.byref(self.desc), C.byref(self.h))
        if rc != 0:
            msg = self.lib.ilqr_last_error(None)
            self.h = None
            raise ILQRError("ilqr_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))

    # -- plumbing ------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.ilqr_last_error(self.h)
            raise ILQRError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "h", None):
            self.lib.ilqr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _host(self, a, shape):
        a = np.ascontiguousarray(np.asarray(a, dtype=self.np_dtype))
        if a.shape != shape:
            a = np.ascontiguousarray(a.reshape(shape))
        return a

    # -- the reference's public methods, batched -------------------------------------------------
    def init_traj(self, x0, u0):
        """iLQR::init_traj (src/ilqr_core.cpp:11-56); returns the initial costs [B]."""
        self.set_initial(x0, u0)
        return self.get("cost")

    def set_initial(self, x0, u0):
        x0 = self._host(x0, (self.B, self.n))
        u0 = self._host(u0, (self.B, self.T, self.m))
        self._keep = (x0, u0)
        self._check(self.lib.ilqr_set_initial(self.h, x0.ctypes.data, u0.ctypes.data, 0), "ilqr_set_initial")

    def set_initial_device(self, x0_ptr, u0_ptr):
        """x0 / u0 already on the handle's device, in the handle's dtype and the ABI layout."""
        self._check(self.lib.ilqr_set_initial(self.h, C.c_void_p(x0_ptr), C.c_void_p(u0_ptr), 1), "ilqr_set_initial")

    def generate_trajectory(self, x0=None, u0=None):
        """The three overloads of iLQR::generate_trajectory (include/ilqr.h:49-51):
        (x0, u0) fresh solve, (x0) warm start from the previous solution, () continue."""
        if x0 is not None and u0 is not None:
            self.set_initial(x0, u0)
        elif x0 is not None:
            self.warm_start(x0)
        else:
            self.resume()
        self.solve()
        return self

    def resume(self):
        """re-enter the loop like a second call of iLQR::generate_trajectory() (src/ilqr_core.cpp:78-102): loop counter 0,
        derivatives refreshed, lambda / dlambda carried over"""
        self._check(self.lib.ilqr_resume(self.h), "ilqr_resume")

    def warm_start(self, x0):
        x0 = self._host(x0, (self.B, self.n))
        self._keep_x0 = x0
        self._check(self.lib.ilqr_warm_start(self.h, x0.ctypes.data, 0), "ilqr_warm_start")

    def iterate(self, n_iters):
        self._check(self.lib.ilqr_iterate(self.h, int(n_iters)), "ilqr_iterate")

    def solve(self):
        self._check(self.lib.ilqr_solve(self.h), "ilqr_solve")

    def backward_once(self, lam=1.0):
        self._check(self.lib.ilqr_backward_once(self.h, float(lam)), "ilqr_backward_once")

    def rollout_once(self, alpha):
        self._check(self.lib.ilqr_rollout_once(self.h, float(alpha)), "ilqr_rollout_once")

    def sync(self):
        self._check(self.lib.ilqr_sync(self.h), "ilqr_sync")

    @property
    def stream(self):
        return self.lib.ilqr_stream(self.h)

    @property
    def launch_count(self):
        return int(self.lib.ilqr_launch_count(self.h))

    # -- results ---------------------------------------------------------------------------------
    def field_shape(self, name):
        B, T, n, m = self.B, self.T, self.n, self.m
        return dict(xs=(B, T + 1, n), us=(B, T, m), K=(B, T, m, n), k=(B, T, m), cost=(B,), dV=(B, 2), Vx0=(B, n),
                    Vxx0=(B, n, n)).get(name, (B,))

    def get(self, name, out=None):
        fid, kind = abi.FIELDS[name]
        dt = np.int32 if kind == "i" else self.np_dtype
        if out is None:
            out = np.empty(self.field_shape(name), dtype=dt)
        self._check(self.lib.ilqr_get(self.h, fid, out.ctypes.data, 0), "ilqr_get(%s)" % name)
        return out

    # -- result files (ilqr_b200/export.py) --------------------------------------------------------
    def output_to_csv(self, filename, b=0):
        """iLQR::output_to_csv (src/ilqr_core.cpp:414-431) for trajectory b of the batch, byte-compatible"""
        from . import export
        export.write_csv(filename, self.get("xs")[b], self.get("us")[b])

    def export_batch(self, filename):
        """the whole batch (xs, us, cost, iterations, status) in one binary file"""
        from . import export
        export.write_batch(filename, self.get("xs"), self.get("us"), self.get("cost"), self.get("iters"), self.get("status"))

    def get_device(self, name, dst_ptr):
        fid, _ = abi.FIELDS[name]
        self._check(self.lib.ilqr_get(self.h, fid, C.c_void_p(dst_ptr), 1), "ilqr_get(%s)" % name)


def make_inputs(seed, B, T, n, m, x_scale=1.0, u_scale=0.5, canonical_first=True):
    """The synthetic instances of SURVEY.md §8d (include/ilqr_synth.h), as f64 arrays."""
    lib = abi.load()
    x0 = np.empty((B, n))
    u0 = np.empty((B, T, m))
    rc = lib.ilqr_make_inputs(seed, B, T, n, m, x_scale, u_scale, int(canonical_first),
                              x0.ctypes.data_as(C.POINTER(C.c_double)), u0.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise ILQRError("ilqr_make_inputs failed (%d)" % rc)
    return x0, u0
