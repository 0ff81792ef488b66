"""Result files of the batched solver.

* one trajectory in the reference's CSV format, byte for byte what `iLQR::output_to_csv` writes
  (reference src/ilqr_core.cpp:414-431): header `x1, ..., xn, u0, ..., um` (one control name too many, :418-419),
  a `%f` row per knot whose last control ends the line, and a terminal row of states only that ends in ", " with no
  newline — the blank that `plot_results.py:15` uses to recognise it;
* the whole batch in one little-endian binary file: 64-byte header (magic "ILQRB200", uint32 version = 1,
  dtype (0 = f64), n, m, T, reserved, uint64 B at offset 32), then xs [B][T+1][n], us [B][T][m], cost [B] as f64 and
  iterations [B], status [B] as int32.  The C++ host layer writes the same file (iLQR::export_batch).
"""
import csv
import struct

import numpy as np

MAGIC = b"ILQRB200"


def write_csv(path, xs, us):
    xs, us = np.asarray(xs, dtype=np.float64), np.asarray(us, dtype=np.float64)
    T, n, m = us.shape[0], xs.shape[1], us.shape[1]
    assert xs.shape[0] == T + 1
    out = ["".join("x%d, " % i for i in range(1, n + 1)) + "".join("u%d, " % j for j in range(m)) + "u%d\n" % m]
    for t in range(T):
        out.append("".join("%f, " % v for v in xs[t]) + "".join("%f, " % v for v in us[t, :-1]) + "%f\n" % us[t, -1])
    out.append("".join("%f, " % v for v in xs[T]))
    with open(path, "w") as f:
        f.write("".join(out))


def read_csv(path, n_states, n_controls):
    """The logic of the reference's plot_results.read_trajectory (plot_results.py:5-21)."""
    states, controls = [], []
    with open(path) as f:
        for i, row in enumerate(csv.reader(f)):
            if i == 0:
                continue
            states.append([float(v) for v in row[:n_states]])
            if row[-1] != " ":  # the terminal row ends in ", "
                controls.append([float(v) for v in row[-n_controls:]])
    return np.array(states), np.array(controls)


def write_batch(path, xs, us, cost, iters, status):
    xs, us = np.ascontiguousarray(xs, dtype="<f8"), np.ascontiguousarray(us, dtype="<f8")
    B, T, n, m = xs.shape[0], us.shape[1], xs.shape[2], us.shape[2]
    head = MAGIC + struct.pack("<6I", 1, 0, n, m, T, 0) + struct.pack("<Q", B) + b"\0" * 24
    assert len(head) == 64
    with open(path, "wb") as f:
        f.write(head)
        f.write(xs.tobytes())
        f.write(us.tobytes())
        f.write(np.ascontiguousarray(cost, dtype="<f8").tobytes())
        f.write(np.ascontiguousarray(iters, dtype="<i4").tobytes())
        f.write(np.ascontiguousarray(status, dtype="<i4").tobytes())


def read_batch(path):
    with open(path, "rb") as f:
        head = f.read(64)
        if len(head) != 64 or head[:8] != MAGIC:
            raise ValueError("%s: not an ilqr_b200 batch file" % path)
        version, dtype, n, m, T, _ = struct.unpack("<6I", head[8:32])
        (B,) = struct.unpack("<Q", head[32:40])
        if version != 1 or dtype != 0:
            raise ValueError("%s: unsupported version / dtype" % path)

        def take(count, dt):
            a = np.frombuffer(f.read(count * np.dtype(dt).itemsize), dtype=dt)
            if a.size != count:
                raise ValueError("%s: truncated" % path)
            return a

        xs = take(B * (T + 1) * n, "<f8").reshape(B, T + 1, n)
        us = take(B * T * m, "<f8").reshape(B, T, m)
        return dict(xs=xs, us=us, cost=take(B, "<f8"), iters=take(B, "<i4"), status=take(B, "<i4"))
