"""CPU, world_size 2 over gloo: the N > 1 host logic of bench.py (ilqr_b200/shard.py) — block ownership,
the single gather of final costs in rank order, and max-time / sum-count reduction."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ilqr_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    with shard.stdout_to_stderr():  # as bench.py brings its process group up
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dist.barrier()
    try:
        B = 6
        # every rank "solves" its own block: cost of instance b is a function of its GLOBAL index
        lo, hi = shard.shard_bounds(B * world, world, rank)
        local = torch.arange(lo, hi, dtype=torch.float64) * 1.5 + 0.25
        got = shard.gather_final_costs(local, dst=0)
        t, c = shard.reduce_step_stats([10.0 + rank, 3.0 - rank], [100 + rank, 7], torch.device("cpu"))
        per = shard.gather_per_rank([50.0 + rank, 1000 + rank], torch.device("cpu"))
        assert per == [[50.0 + r, 1000.0 + r] for r in range(world)]
        if rank == 0:
            out.put(("costs", got.numpy().tolist()))
            out.put(("stats", t, c))
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    for total, world in ((4096, 1), (1048576, 8), (10, 3), (7, 8)):
        blocks = [shard.shard_bounds(total, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == total
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1
    assert shard.shard_bounds(1048576, 8, 3) == (393216, 524288)   # BASELINE configs[4]: 131 072 per GPU
    assert shard.rank_seed(12345, 0) == 12345 and shard.rank_seed(12345, 5) == 12350


def test_gather_of_final_costs_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    msgs = [out.get(timeout=10) for _ in range(2)]
    costs = next(m for m in msgs if m[0] == "costs")[1]
    stats = next(m for m in msgs if m[0] == "stats")
    assert np.allclose(costs, np.arange(12) * 1.5 + 0.25)          # rank order == global instance order
    assert stats[1] == [11.0, 3.0] and stats[2] == [201.0, 14.0]    # MAX of times, SUM of counts


def test_single_process_is_identity():
    x = torch.arange(4, dtype=torch.float64)
    assert shard.gather_final_costs(x) is x


def test_cpp_host_partition_equals_python_partition():
    """BatchSolver::shard_bounds (ilqr_b200/host/batch_solver.cpp, the one-process multi-GPU host path) cuts a batch into
    the same contiguous blocks as shard.shard_bounds (the torchrun path)"""
    import os
    import subprocess
    from ilqr_b200 import shard
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ilqr_b200", "host", "_build", "bench_batch")
    if not os.path.exists(exe):
        import pytest
        pytest.skip("host binaries not built")
    for total, world in ((1048576, 8), (4096, 3), (7, 8), (100, 1)):
        out = subprocess.run([exe, "--shards", str(total), str(world)], capture_output=True, text=True, check=True).stdout.split()
        got = [(int(out[2 * r]), int(out[2 * r + 1])) for r in range(world)]
        assert got == [shard.shard_bounds(total, world, r) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == total


def test_stdout_to_stderr_redirect(tmp_path):
    """shard.stdout_to_stderr: what a native library writes to file descriptor 1 inside the block lands on stderr, and
    stdout is back afterwards (bench.py wraps the NCCL start-up in it: NCCL prints its version banner on stdout)"""
    import subprocess
    import sys
    code = ("import ctypes, os, sys\n"
            "sys.path.insert(0, %r)\n"
            "from ilqr_b200 import shard\n"
            "libc = ctypes.CDLL(None)\n"
            "print('before', flush=True)\n"
            "with shard.stdout_to_stderr():\n"
            "    libc.puts(b'banner from a native library')\n"
            "    libc.fflush(None)\n"
            "print('{\"json\": 1}', flush=True)\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split("\n")[:2] == ["before", '{"json": 1}']
    assert "banner from a native library" in out.stderr
