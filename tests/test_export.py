"""Result files: the reference's CSV format byte for byte, and the whole-batch binary file.

The reference writes `ilqr_result.csv` at the end of every solve (src/ilqr_core.cpp:300,414-431) and its
plot_results.py:5-21 reads it back, telling the terminal row by the blank that ends it.  tests/golden/ref_cli_*.csv are
the files the reference's own CLI wrote (tests/golden/make_golden.py)."""
import os
import subprocess

import numpy as np
import pytest

from ilqr_b200 import export

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSV_TOOL = os.path.join(ROOT, "ilqr_b200", "host", "_build", "csv_tool")
CASES = [("acrobot", "acrobot_cli_T499", 4, 1), ("integrator", "integrator_cli_T99", 4, 2)]


@pytest.mark.parametrize("which,case,n,m", CASES)
def test_python_writer_equals_reference_file(tmp_path, golden_solver, which, case, n, m):
    """same trajectory (the golden final xs / us are the reference's own) -> the same bytes"""
    g = golden_solver
    out = tmp_path / "ours.csv"
    export.write_csv(out, g[case + "/final_xs"], g[case + "/final_us"])
    ref = open(os.path.join(ROOT, "tests", "golden", "ref_cli_%s.csv" % which), "rb").read()
    assert open(out, "rb").read() == ref
    assert not ref.endswith(b"\n") and ref.endswith(b", ")  # the terminal row has no newline (:427)


@pytest.mark.skipif(not os.path.exists(CSV_TOOL), reason="host binaries not built")
@pytest.mark.parametrize("which,case,n,m", CASES)
def test_cpp_host_writer_equals_reference_file(tmp_path, golden_solver, which, case, n, m):
    """the C++ host layer's writer (ilqr_b200/host/ilqr_host.cpp: ilqr_write_csv, behind iLQR::output_to_csv)"""
    g = golden_solver
    xs, us = g[case + "/final_xs"], g[case + "/final_us"]
    raw = tmp_path / "traj.bin"
    with open(raw, "wb") as f:
        f.write(np.ascontiguousarray(xs, dtype="<f8").tobytes())
        f.write(np.ascontiguousarray(us, dtype="<f8").tobytes())
    out = tmp_path / "host.csv"
    subprocess.check_call([CSV_TOOL, str(us.shape[0]), str(n), str(m), str(raw), str(out)])
    assert open(out, "rb").read() == open(os.path.join(ROOT, "tests", "golden", "ref_cli_%s.csv" % which), "rb").read()


@pytest.mark.parametrize("which,case,n,m", CASES)
def test_plot_results_reader_logic(golden_solver, which, case, n, m):
    """plot_results.read_trajectory's logic on the reference's file: T + 1 states, T controls, values to 6 decimals"""
    g = golden_solver
    states, controls = export.read_csv(os.path.join(ROOT, "tests", "golden", "ref_cli_%s.csv" % which), n, m)
    xs, us = g[case + "/final_xs"], g[case + "/final_us"]
    assert states.shape == xs.shape and controls.shape == us.shape
    assert np.abs(states - xs).max() <= 5.1e-7 and np.abs(controls - us).max() <= 5.1e-7


def test_batch_file_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    B, T, n, m = 7, 13, 4, 2
    xs, us = rng.normal(size=(B, T + 1, n)), rng.normal(size=(B, T, m))
    cost, iters, status = rng.uniform(size=B), rng.integers(1, 100, B), rng.integers(1, 5, B)
    p = tmp_path / "batch.bin"
    export.write_batch(p, xs, us, cost, iters, status)
    assert os.path.getsize(p) == 64 + 8 * (xs.size + us.size + B) + 4 * 2 * B
    d = export.read_batch(p)
    assert np.array_equal(d["xs"], xs) and np.array_equal(d["us"], us) and np.array_equal(d["cost"], cost)
    assert np.array_equal(d["iters"], iters) and np.array_equal(d["status"], status)
    with open(p, "r+b") as f:
        f.write(b"NOTILQR!")
    with pytest.raises(ValueError):
        export.read_batch(p)
