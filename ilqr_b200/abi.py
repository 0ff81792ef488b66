"""ctypes mirror of include/ilqr_b200.h (the C ABI of libilqr_b200.so).

The structs here are shared by the product wrapper (ilqr_b200.solver) and by the test-side
wrapper of the CPU oracle (tests/oracleport.py), which checks the same API semantics.
Loading the CUDA library fails loudly when it has not been built: there is no CPU fallback.
"""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("ILQR_B200_LIB") or os.path.join(ROOT, "ilqr_b200", "libilqr_b200.so")

MAX_N, MAX_M, MAX_ALPHA = 8, 4, 16
MODEL_ACROBOT, MODEL_DOUBLE_INTEGRATOR = 0, 1
MODEL_USER_BASE = 100  # ids of models registered at run time (register_model)
F64, F32 = 0, 1
COST_FD, COST_ANALYTIC = 0, 1
FLAG_ENGINE_WARP = 1  # ilqr_desc.flags
FLAG_CLAMP_ROLLOUT, FLAG_ANALYTIC_DYN = 2, 4
FLAG_FAST_FMA = 8
RUNNING, EXIT_GRAD, EXIT_TOLFUN, EXIT_LAMBDA_MAX, EXIT_MAXITER = 0, 1, 2, 3, 4
STATUS_NAMES = {0: "RUNNING", 1: "GRAD", 2: "TOLFUN", 3: "LAMBDA_MAX", 4: "MAXITER"}

# ilqr_get field ids: name -> (id, kind) with kind 'r' = handle dtype, 'i' = int32
FIELDS = {
    "xs": (0, "r"), "us": (1, "r"), "K": (2, "r"), "k": (3, "r"), "cost": (4, "r"), "dV": (5, "r"),
    "Vx0": (6, "r"), "Vxx0": (7, "r"), "lambda": (8, "r"), "dlambda": (9, "r"), "gnorm": (10, "r"),
    "iters": (11, "i"), "status": (12, "i"), "alpha_index": (13, "i"), "n_accept": (14, "i"),
    "n_reject": (15, "i"), "n_backward": (16, "i"), "diverge": (17, "i"),
}


class Params(C.Structure):
    _fields_ = [
        ("max_iter", C.c_int32), ("n_alpha", C.c_int32),
        ("tol_fun", C.c_double), ("tol_grad", C.c_double),
        ("lambda_init", C.c_double), ("dlambda_init", C.c_double), ("lambda_factor", C.c_double),
        ("lambda_max", C.c_double), ("lambda_min", C.c_double), ("z_min", C.c_double),
        ("grad_lambda_gate", C.c_double),
        ("alpha", C.c_double * MAX_ALPHA),
        ("qp_max_iter", C.c_int32), ("reserved0", C.c_int32),
        ("qp_min_grad", C.c_double), ("qp_min_rel_improve", C.c_double), ("qp_step_dec", C.c_double),
        ("qp_min_step", C.c_double), ("qp_armijo", C.c_double), ("qp_clamp_tol", C.c_double),
        ("fd_eps", C.c_double),
    ]


class Desc(C.Structure):
    _fields_ = [
        ("model_id", C.c_int32), ("dtype", C.c_int32), ("cost_deriv", C.c_int32), ("device", C.c_int32),
        ("T", C.c_int32), ("override_limits", C.c_int32), ("flags", C.c_int32), ("reserved1", C.c_int32),
        ("B", C.c_int64), ("dt", C.c_double),
        ("u_min", C.c_double * MAX_M), ("u_max", C.c_double * MAX_M),
        ("model_params", C.c_double * 16),
        ("params", Params),
    ]


# The reference's constants (include/ilqr.h:14-25, include/boxqp.h:19-24,61-64, include/finite_diff.h:9).
REFERENCE_ALPHA = (1.0000, 0.5012, 0.2512, 0.1259, 0.0631, 0.0316, 0.0158, 0.0079, 0.0040, 0.0020, 0.0010)


def default_params():
    p = Params()
    p.max_iter, p.n_alpha = 100, 11
    p.tol_fun = p.tol_grad = 1e-6
    p.lambda_init = p.dlambda_init = 1.0
    p.lambda_factor, p.lambda_max, p.lambda_min, p.z_min = 1.6, 1e11, 1e-8, 0.0
    p.grad_lambda_gate = 1e-5
    for i, a in enumerate(REFERENCE_ALPHA):
        p.alpha[i] = a
    p.qp_max_iter = 100
    p.qp_min_grad = p.qp_min_rel_improve = 1e-8
    p.qp_step_dec, p.qp_min_step, p.qp_armijo, p.qp_clamp_tol = 0.6, 1e-22, 0.1, 1e-4
    p.fd_eps = 1e-3
    return p


MODEL_DIMS = {MODEL_ACROBOT: (4, 1), MODEL_DOUBLE_INTEGRATOR: (4, 2)}


def make_desc(model=MODEL_ACROBOT, T=200, B=1, dt=0.02, dtype=F64, cost_deriv=COST_FD, device=0, u_min=None,
              u_max=None, goal=None, params=None, flags=0):
    d = Desc()
    d.model_id, d.dtype, d.cost_deriv, d.device = model, dtype, cost_deriv, device
    d.flags = int(flags)
    d.T, d.B, d.dt = int(T), int(B), float(dt)
    if (u_min is None) != (u_max is None):
        raise ValueError("give both u_min and u_max or neither")
    if u_min is not None:
        d.override_limits = 1
        for j, (lo, hi) in enumerate(zip(list(u_min), list(u_max))):
            d.u_min[j], d.u_max[j] = lo, hi
    if goal is not None:
        for i, g in enumerate(goal):
            d.model_params[i] = g
    d.params = params if params is not None else default_params()
    return d


_lib = None

EXPORTS = [
    "ilqr_default_params", "ilqr_model_info", "ilqr_create", "ilqr_destroy", "ilqr_last_error",
    "ilqr_set_initial", "ilqr_warm_start", "ilqr_resume", "ilqr_iterate", "ilqr_solve", "ilqr_backward_once",
    "ilqr_rollout_once", "ilqr_get", "ilqr_sync", "ilqr_stream", "ilqr_launch_count", "ilqr_make_inputs",
    "ilqr_version", "ilqr_register_model", "ilqr_compile_model", "ilqr_measure_fp64",
]


def load():
    """dlopen libilqr_b200.so (built by __graft_entry__.build / `make -C ilqr_b200/csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
            "g.build()').  There is no CPU fallback for the solver.")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.ilqr_default_params.argtypes = [C.POINTER(Params)]
    L.ilqr_model_info.argtypes = [C.c_int32, ip, ip, dp, dp]
    L.ilqr_create.argtypes = [C.POINTER(Desc), C.POINTER(vp)]
    L.ilqr_destroy.argtypes = [vp]
    L.ilqr_last_error.restype = C.c_char_p
    L.ilqr_last_error.argtypes = [vp]
    L.ilqr_set_initial.argtypes = [vp, vp, vp, C.c_int]
    L.ilqr_warm_start.argtypes = [vp, vp, C.c_int]
    L.ilqr_resume.argtypes = [vp]
    L.ilqr_iterate.argtypes = [vp, C.c_int]
    L.ilqr_solve.argtypes = [vp]
    L.ilqr_backward_once.argtypes = [vp, C.c_double]
    L.ilqr_rollout_once.argtypes = [vp, C.c_double]
    L.ilqr_get.argtypes = [vp, C.c_int, vp, C.c_int]
    L.ilqr_sync.argtypes = [vp]
    L.ilqr_stream.restype = vp
    L.ilqr_stream.argtypes = [vp]
    L.ilqr_launch_count.restype = C.c_int64
    L.ilqr_launch_count.argtypes = [vp]
    L.ilqr_make_inputs.argtypes = [C.c_uint64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                   C.c_int, dp, dp]
    L.ilqr_version.restype = C.c_char_p
    L.ilqr_measure_fp64.argtypes = [C.c_int32, dp, dp, dp]
    L.ilqr_register_model.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, dp, dp, ip]
    L.ilqr_compile_model.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.c_size_t]
    _lib = L
    return L


def model_dims(model):
    """(n, m) of a built-in or registered model."""
    if model in MODEL_DIMS:
        return MODEL_DIMS[model]
    n, m = C.c_int32(), C.c_int32()
    if load().ilqr_model_info(model, C.byref(n), C.byref(m), None, None) != 0:
        raise ValueError("unknown model id %r" % (model,))
    return n.value, m.value


def register_model(struct_name, cuda_source, n, m, u_min, u_max):
    """A user model as CUDA source (the device twin of a `Model` subclass, include/ilqr_b200.h: ilqr_register_model).
    Returns the model id for BatchILQR(model=...).  Compiled by NVRTC when a handle of it first launches."""
    lo = (C.c_double * MAX_M)(*[float(v) for v in u_min])
    hi = (C.c_double * MAX_M)(*[float(v) for v in u_max])
    mid = C.c_int32()
    rc = load().ilqr_register_model(struct_name.encode(), cuda_source.encode(), int(n), int(m), lo, hi, C.byref(mid))
    if rc != 0:
        raise ValueError("ilqr_register_model failed (%d): %s" % (rc, (load().ilqr_last_error(None) or b"?").decode()))
    return mid.value


def compile_model(model, dtype=F64, cost_deriv=COST_ANALYTIC):
    """NVRTC-compile a registered model now (no GPU needed); returns (rc, compiler log)."""
    log = C.create_string_buffer(1 << 16)
    rc = load().ilqr_compile_model(int(model), int(dtype), int(cost_deriv), log, len(log))
    return rc, log.value.decode(errors="replace")
