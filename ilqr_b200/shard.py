"""Multi-GPU plumbing for the batched solve: independent problem instances shard trivially, so each rank
(one process per GPU) owns a contiguous block of the global batch and nothing is exchanged during the
solve; the ONLY collective is one gather of the final per-instance costs to rank 0 (BASELINE.json north_star,
SURVEY.md §8e).  Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
import contextlib
import os
import sys

import torch
import torch.distributed as dist


@contextlib.contextmanager
def stdout_to_stderr():
    """Inside the block, file descriptor 1 points at stderr: what native libraries print on STDOUT while a process
    group comes up (NCCL's "NCCL version ..." banner, printed at NCCL_DEBUG=WARN and VERSION) lands on stderr, so a
    program whose stdout is a protocol (bench.py: ONE JSON line) keeps it clean.  Restored on exit."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def shard_bounds(total, world, rank):
    """contiguous block [lo, hi) of `total` instances owned by `rank`; blocks differ by at most one"""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_seed(seed, rank):
    """Synthetic instances of rank r: the generator of include/ilqr_synth.h seeded with seed + r, so every
    rank solves different problems without materialising the global batch on each of them."""
    return int(seed) + int(rank)


def gather_final_costs(local_cost, dst=0):
    """ONE collective per solve: rank `dst` receives every rank's final costs (equal block sizes), concatenated in
    rank order; other ranks receive None.  With world size 1 it is the identity."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_cost
    world, rank = dist.get_world_size(), dist.get_rank()
    bufs = [torch.empty_like(local_cost) for _ in range(world)] if rank == dst else None
    dist.gather(local_cost, bufs, dst=dst)
    return torch.cat(bufs) if rank == dst else None


def reduce_step_stats(elapsed_ms, counts, device):
    """whole-job timing: MAX over ranks of the device-timed durations, SUM over ranks of the trip counters"""
    t = torch.tensor(list(elapsed_ms), dtype=torch.float64, device=device)
    c = torch.tensor(list(counts), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.tolist(), c.tolist()


def gather_per_rank(values, device):
    """[world][len(values)]: every rank's own numbers (device-timed ms, trip counts), for the bench line's breakdown"""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return [o.tolist() for o in out]
    return [t.tolist()]
