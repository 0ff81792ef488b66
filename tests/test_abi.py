"""CPU: the C-ABI library loads and exports every symbol include/ilqr_b200.h declares; the calls that
need no GPU behave; the ctypes mirror of the structs matches the header's layout."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from ilqr_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ilqr_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ilqr_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(abi.LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "ilqr_b200", "csrc")])
    return abi.load()


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    assert set(names) == set(abi.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None


def test_struct_layout_matches_header():
    """compile a probe of sizeof/offsetof with gcc and compare with the ctypes mirror"""
    src = os.path.join(ROOT, "tests", "_build", "layout.c")
    exe = os.path.join(ROOT, "tests", "_build", "layout")
    os.makedirs(os.path.dirname(src), exist_ok=True)
    with open(src, "w") as f:
        f.write('#include <stdio.h>\n#include <stddef.h>\n#include "ilqr_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n",'
                'sizeof(ilqr_params), sizeof(ilqr_desc), offsetof(ilqr_params, qp_max_iter), offsetof(ilqr_params, fd_eps),'
                'offsetof(ilqr_desc, B), offsetof(ilqr_desc, params));return 0;}\n')
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), "-o", exe, src])
    got = [int(v) for v in subprocess.check_output([exe]).split()]
    exp = [C.sizeof(abi.Params), C.sizeof(abi.Desc), abi.Params.qp_max_iter.offset, abi.Params.fd_eps.offset,
           abi.Desc.B.offset, abi.Desc.params.offset]
    assert got == exp


def test_defaults_are_the_reference_constants(lib):
    p = abi.Params()
    assert lib.ilqr_default_params(C.byref(p)) == 0
    q = abi.default_params()
    for name, _ in abi.Params._fields_:
        a, b = getattr(p, name), getattr(q, name)
        assert (list(a) == list(b)) if hasattr(a, "__len__") else (a == b), name
    assert p.max_iter == 100 and p.n_alpha == 11 and p.lambda_factor == 1.6 and p.lambda_max == 1e11   # ilqr.h:14-22
    assert list(p.alpha)[:11] == list(abi.REFERENCE_ALPHA)                                               # ilqr.h:24
    assert p.qp_max_iter == 100 and p.qp_step_dec == 0.6 and p.qp_armijo == 0.1 and p.qp_min_step == 1e-22  # boxqp.h:19-24
    assert p.fd_eps == 1e-3                                                                              # finite_diff.h:9


def test_model_info(lib):
    n, m = C.c_int32(), C.c_int32()
    lo, hi = (C.c_double * 4)(), (C.c_double * 4)()
    assert lib.ilqr_model_info(abi.MODEL_ACROBOT, C.byref(n), C.byref(m), lo, hi) == 0
    assert (n.value, m.value, lo[0], hi[0]) == (4, 1, -5.0, 5.0)            # acrobot.h:27-28,37
    assert lib.ilqr_model_info(abi.MODEL_DOUBLE_INTEGRATOR, C.byref(n), C.byref(m), lo, hi) == 0
    assert (n.value, m.value, lo[1], hi[1]) == (4, 2, -0.5, 0.5)            # double_integrator.h:16-17,25-26
    assert lib.ilqr_model_info(9, C.byref(n), C.byref(m), lo, hi) < 0


def test_make_inputs_is_the_shared_generator(lib):
    import bench
    from ilqr_b200.solver import make_inputs
    x0, u0 = make_inputs(12345, 5, 9, 4, 1)
    y0, v0 = bench.synth_inputs_cpu(5, 9, 12345)
    assert np.array_equal(x0, y0) and np.array_equal(u0, v0)
    assert (x0[0] == 0).all() and (u0[0] == 0).all() and np.abs(x0[1:]).max() <= 1 and np.abs(u0[1:]).max() <= 0.5


def test_create_fails_loudly_without_gpu_or_on_bad_arguments(lib):
    import torch
    h = C.c_void_p()
    bad = abi.make_desc(model=5, T=10, B=1)
    assert lib.ilqr_create(C.byref(bad), C.byref(h)) == -1 and b"model" in lib.ilqr_last_error(None)
    for kw in (dict(T=0), dict(B=0), dict(dt=0.0)):
        d = abi.make_desc(**{**dict(T=10, B=1, dt=0.02), **kw})
        assert lib.ilqr_create(C.byref(d), C.byref(h)) == -1
    if not torch.cuda.is_available():
        ok = abi.make_desc(T=10, B=1)
        rc = lib.ilqr_create(C.byref(ok), C.byref(h))
        assert rc == -2 and b"no CPU path" in lib.ilqr_last_error(None)   # ILQR_E_CUDA: never a silent fallback
    assert lib.ilqr_version().startswith(b"ilqr_b200")
