#!/usr/bin/env python
"""bench.py — batched iLQR iterations/sec (acrobot, T=200) on N B200s, next to the CPU reference.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--config cfg2|cfg3|cfg4|cfg5]

One STEP = one pass of the hot path over one batch of synthetic instances (SURVEY.md §8d,
include/ilqr_synth.h, seed 12345): every instance is initialised (`iLQR::init_traj`) and solved to
termination (`iLQR::generate_trajectory`) — derivative sweeps, backward passes with boxQP, line
searches, lambda schedule.  The metric counts the loop trips of src/ilqr_core.cpp:103-288 actually
executed, summed over the batch, per second.

  value  inputs already resident in HBM; timed with CUDA events on the library's stream.
  e2e    the same through the C ABI with HOST buffers: pinned x0/u0 -> H2D -> solve -> D2H of the
         final costs and trip counts, copies inside the timed region.
  roofline  one ilqr_solve (all its kernels): ALGORITHMIC bytes (40 096 B per accepted and
         32 064 B per rejected trip at n=4, m=1, T=200, f64) / its CUDA-event duration, against the
         measured HBM copy bandwidth in MEASURED_PEAKS.json; `traffic` = DRAM bytes of one solve from the
         committed ncu counters (profiles/traffic.json).  `secondary` = the ceiling that actually binds:
         useful flops against the fp64 issue rate measured on this GPU (ilqr_measure_fp64),
         with warp instructions per trip / lanes per instruction from the same ncu capture.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/libref_oracle.so, built from /root/reference
         by oracle/Makefile) on the host cores, one forked worker per core, on a bounded sample of
         the same batch; `terminal_cost_vs_gpu` compares the two on that sample, and `like_for_like`
         repeats it with the GPU in the reference's own derivative mode (finite-difference costs) plus
         K, k, cost after 5 trips.
  e2e_cpp_host  (N = 1) the same batch through the C++ host layer (ilqr_b200/host: BatchSolver, pinned inputs, one
         ncclAllGather), end to end, as reported by ilqr_b200/host/_build/bench_batch.
  per_rank  (N > 1) every rank's own ms per step and trips: what the max over ranks is the max of.
  extras  (default line only) short runs of BASELINE configs[2], [3], the configs[4] shard and the opt-in
         FMA build, each with its rate, exit-status histogram and a parity sample against the reference.

N > 1 (torchrun, one rank per GPU): every rank solves its own B instances (weak scaling, no
data-path collective) and the final costs are gathered to rank 0 with ONE NCCL gather per step.
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "batched iLQR iterations/sec (acrobot T=200)"
UNIT = "iterations/s"
SEED = 12345

CONFIGS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "cfg2": dict(workload="acrobot batch=4096 T=200 fp64, analytic cost derivatives, FD fx/fu (BASELINE configs[1])",
                 B=4096, T=200, cost_deriv="analytic", limits=None),
    # configs[2]: f32, longer horizon, finite-difference fx/fu, closed-form cost derivatives
    "cfg3": dict(workload="acrobot batch=65536 T=500 fp32, finite-difference fx/fu, analytic cost derivatives (BASELINE configs[2])",
                 B=65536, T=500, cost_deriv="analytic", limits=None, dtype="f32"),
    # configs[3]: control-limited, full finite differences
    "cfg4": dict(workload="control-limited acrobot (+-1.5) batch=8192 T=200 fp64, full finite differences (BASELINE configs[3])",
                 B=8192, T=200, cost_deriv="fd", limits=1.5),
    # configs[4] per-GPU shard: 1 048 576 / 8
    "cfg5": dict(workload="acrobot batch=131072 per GPU T=200 fp64, analytic cost derivatives (BASELINE configs[4] shard)",
                 B=131072, T=200, cost_deriv="analytic", limits=None),
    # the headline configuration on the opt-in build with fused multiply-add contraction (ILQR_FLAG_FAST_FMA = 8): within
    # 1e-6 of the oracle after a handful of trips, not bit-identical to the reference's no-FMA arithmetic
    "cfg2_fast_fma": dict(workload="acrobot batch=4096 T=200 fp64, analytic cost derivatives, ILQR_FLAG_FAST_FMA (opt-in; not the headline)",
                          B=4096, T=200, cost_deriv="analytic", limits=None, flags=8),
}


def bytes_per_trip(n, m, T, s):
    """algorithmic HBM bytes of one accepted / rejected loop trip of one trajectory."""
    acc = s * (3 * ((T + 1) * n + T * m) + 2 * T * m * (n + 1))
    rej = s * (2 * ((T + 1) * n + T * m) + 2 * T * m * (n + 1))
    return acc, rej


# ------------------------------------------------------------------------------------------------
# CPU reference arm
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, b0, b1, x0, u0, T, limits, cost_deriv = args
    if kind == "reference":
        import refharness as R
        L = R.lib()
        L.ref_solve_range.restype = C.c_long
        L.ref_solve_range.argtypes = [C.c_void_p, C.c_long, C.c_long, R._dp, R._dp, C.c_int, C.c_int, R._dp, R._ip, R._ip]
        lo = [-limits] if limits else None
        hi = [limits] if limits else None
        s = R.RefSolver(R.ACROBOT, 0.02, u_min=lo, u_max=hi)
        cost = np.empty(b1 - b0)
        iters = np.zeros(b1 - b0, dtype=np.int32)
        status = np.zeros(b1 - b0, dtype=np.int32)
        t0 = time.perf_counter()
        total = L.ref_solve_range(s.h, b0, b1, R._p(x0), R._p(u0), T, -1, R._p(cost), iters.ctypes.data_as(R._ip),
                                  status.ctypes.data_as(R._ip))
        return total, time.perf_counter() - t0, cost
    import oracleport as O
    from ilqr_b200 import abi
    kw = dict(u_min=[-limits], u_max=[limits]) if limits else {}
    desc = abi.make_desc(model=abi.MODEL_ACROBOT, T=T, dt=0.02,
                         cost_deriv=abi.COST_ANALYTIC if cost_deriv == "analytic" else abi.COST_FD, **kw)
    t0 = time.perf_counter()
    r = O.solve_range(desc, x0, u0, b0, b1)
    return int(r["total_trips"]), time.perf_counter() - t0, r["cost"]


def cpu_reference_kind():
    import refharness as R
    return "reference" if R.available() else "port"


def run_cpu_sample(cfg, x0, u0, n_sample, cores):
    """Solve instances [0, n_sample) with the reference on `cores` forked workers; returns
    (iterations, wall seconds, costs)."""
    kind = cpu_reference_kind()
    n_sample = min(n_sample, x0.shape[0])
    x0 = np.ascontiguousarray(x0[:n_sample])
    u0 = np.ascontiguousarray(u0[:n_sample])
    cores = max(1, min(cores, n_sample))
    bounds = np.linspace(0, n_sample, cores + 1).astype(int)
    jobs = [(kind, int(bounds[i]), int(bounds[i + 1]), x0, u0, cfg["T"], cfg["limits"], cfg["cost_deriv"])
            for i in range(cores) if bounds[i + 1] > bounds[i]]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(jobs)) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    iters = sum(r[0] for r in res)
    cost = np.concatenate([r[2] for r in res])
    return kind, iters, wall, cost, len(jobs)


def synth_inputs(B, T, seed):
    """include/ilqr_synth.h through the library's own generator when it is built, else numpy-free C port."""
    from ilqr_b200.solver import make_inputs
    return make_inputs(seed, B, T, 4, 1)


def synth_inputs_cpu(B, T, seed):
    """Same generator without loading the CUDA library (the reference arm must not touch our engine)."""
    src = os.path.join(ROOT, "tests", "_build", "libsynth.so")
    if not os.path.exists(src):
        os.makedirs(os.path.dirname(src), exist_ok=True)
        csrc = os.path.join(ROOT, "tests", "_build", "synth.c")
        with open(csrc, "w") as f:
            f.write('#include "ilqr_synth.h"\nvoid synth(unsigned long long seed, long B, int T, int n, int m, double xs, double us, int c, double *x0, double *u0) { ilqr_synth_fill(seed, (size_t)B, T, n, m, xs, us, c, x0, u0); }\n')
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", src, csrc])
    L = C.CDLL(src)
    dp = C.POINTER(C.c_double)
    L.synth.argtypes = [C.c_ulonglong, C.c_long, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, dp, dp]
    x0 = np.empty((B, 4))
    u0 = np.empty((B, T, 1))
    L.synth(seed, B, T, 4, 1, 1.0, 0.5, 1, x0.ctypes.data_as(dp), u0.ctypes.data_as(dp))
    return x0, u0


def reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B = args.batch or cfg["B"]
    per_core = args.cpu_per_core
    n_sample = min(B, cores * per_core)
    x0, u0 = synth_inputs_cpu(n_sample, cfg["T"], SEED)
    times, iters = [], []
    kind = cpu_reference_kind()
    if kind == "reference":
        import refharness as R
        R.lib()  # dlopen oracle/_ref/libref_oracle.so in THIS process before forking the workers
    for step in range(args.warmup + args.steps):
        kind, it, wall, _, used = run_cpu_sample(cfg, x0, u0, n_sample, cores)
        if step >= args.warmup:
            times.append(wall)
            iters.append(it)
    total_t = sum(times)
    value = sum(iters) / total_t
    sample = "first %d of %d instances per step, solved to termination, %d forked workers" % (n_sample, B, used)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "sample": sample, "sample_instances": n_sample, "timer": "host wall clock (CPU arm)",
                   "cost_deriv": "fd (reference has no other)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
                power.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def flops_per_trip(T, cost_fd):
    """SURVEY.md §8d "secondary ceiling": useful flops of one accepted / rejected loop trip of one acrobot trajectory
    (n = 4, m = 1; a sin or cos counted as ONE flop, as the survey does).  Per timestep: 10 Euler steps of ~75 flops for
    the finite-difference Jacobians, (FD-cost mode) 70 cost evaluations of ~21 flops plus the stencil arithmetic,
    ~750 for the backward step with its boxQP, 11 candidate rollouts + commit of ~118 each.  A rejected trip reuses the
    derivatives of the last accepted one."""
    deriv = 10 * 75 + (70 * 21 + 80 if cost_fd else 0)
    back, search = 750, 12 * 118
    return T * (deriv + back + search), T * (back + search)


class Runner:
    """One configuration on this rank's GPU: resident / end-to-end / fixed-N timings through the C ABI."""

    def __init__(self, cfg, B, rank, local, dev, world, gather):
        import torch
        from ilqr_b200 import abi, shard
        from ilqr_b200.solver import BatchILQR
        self.torch, self.abi = torch, abi
        self.cfg, self.B, self.T, self.world, self.dev = cfg, B, cfg["T"], world, dev
        cd = abi.COST_ANALYTIC if cfg["cost_deriv"] == "analytic" else abi.COST_FD
        kw = dict(u_min=[-cfg["limits"]], u_max=[cfg["limits"]]) if cfg["limits"] else {}
        self.x0, self.u0 = synth_inputs(B, self.T, shard.rank_seed(SEED, rank))  # every rank owns different instances
        self.f32 = cfg.get("dtype") == "f32"
        self.sbytes, np_t, self.th_t = (4, np.float32, torch.float32) if self.f32 else (8, np.float64, torch.float64)
        self.solver = BatchILQR(abi.MODEL_ACROBOT, T=self.T, B=B, dt=0.02, cost_deriv=cd, device=local,
                                dtype=abi.F32 if self.f32 else abi.F64, flags=cfg.get("flags", 0), **kw)
        self.stream = torch.cuda.ExternalStream(self.solver.stream, device=dev)
        self.x0_h = torch.from_numpy(self.x0.astype(np_t)).pin_memory()  # the handle's dtype: set / get are plain copies
        self.u0_h = torch.from_numpy(self.u0.astype(np_t)).pin_memory()
        self.cost_h = torch.empty(B, dtype=self.th_t).pin_memory()
        self.iters_h = torch.empty(B, dtype=torch.int32).pin_memory()
        self.x0_d, self.u0_d = self.x0_h.to(dev), self.u0_h.to(dev)
        self.cost_d = torch.empty(B, dtype=self.th_t, device=dev)
        self.gather = gather
        torch.cuda.synchronize()

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def step_resident(self):
        """inputs resident in HBM -> final costs resident in HBM (rank 0 after the gather)"""
        s = self.solver
        e0, e1, e2, e3 = self.ev(), self.ev(), self.ev(), self.ev()
        with self.torch.cuda.stream(self.stream):
            e0.record()
            s.set_initial_device(self.x0_d.data_ptr(), self.u0_d.data_ptr())
            e1.record()
            s.solve()
            e2.record()
            s.get_device("cost", self.cost_d.data_ptr())
            self.gather(self.cost_d)  # the one collective of the job (NCCL over NVLink when N > 1)
            e3.record()
        return e0, e1, e2, e3

    def step_e2e(self):
        """host buffers in, host results out, through the C ABI"""
        s = self.solver
        e0, e3 = self.ev(), self.ev()
        with self.torch.cuda.stream(self.stream):
            e0.record()
            s.set_initial(self.x0_h.numpy(), self.u0_h.numpy())
            s.solve()
            s.get("cost", out=self.cost_h.numpy())
            s.get("iters", out=self.iters_h.numpy())
            if self.world > 1:
                self.cost_d.copy_(self.cost_h, non_blocking=True)
                self.gather(self.cost_d)
            e3.record()
        return e0, e3

    def fixed_n(self, n_fixed, reps, flush):
        """every instance runs exactly N trips: the engine's rate without the ragged-termination tail"""
        s = self.solver
        ms, trips = [], 0
        for _ in range(reps):
            flush()
            with self.torch.cuda.stream(self.stream):
                s.set_initial_device(self.x0_d.data_ptr(), self.u0_d.data_ptr())
                f0, f1 = self.ev(), self.ev()
                f0.record()
                s.iterate(n_fixed)
                f1.record()
            s.sync()
            ms.append(f0.elapsed_time(f1))
            trips += int(s.get("iters").sum())
        return sum(ms), trips


def parity_sample(cfg, x0, u0, gpu_cost, n_sample, cores):
    """the first n_sample instances solved by the reference on the host cores, against the GPU's terminal costs"""
    kind, it, wall, cpu_cost, used = run_cpu_sample(cfg, x0, u0, n_sample, cores)
    g = np.asarray(gpu_cost[:len(cpu_cost)], dtype=np.float64)
    rel = np.abs(g - cpu_cost) / np.maximum(np.abs(cpu_cost), 1e-300)
    return kind, it, wall, used, {"instances": int(len(cpu_cost)), "median_rel_diff": float(np.median(rel)),
                                  "frac_within_1e-6": float((rel <= 1e-6).mean()), "frac_within_1e-3": float((rel <= 1e-3).mean()),
                                  "frac_within_1e-2": float((rel <= 1e-2).mean())}


def _ref_checkpoint_worker(args):
    """K, k, cost of the reference after n trips, instance by instance (the probe's replica loop)"""
    b0, b1, x0, u0, limits, n = args
    import refharness as R
    out = []
    for b in range(b0, b1):
        s = R.RefSolver(R.ACROBOT, 0.02, u_min=[-limits] if limits else None, u_max=[limits] if limits else None)
        s.init(x0[b], u0[b])
        s.iterate(n)
        out.append((s.get("K"), s.get("k"), s.cost))
    return out


def like_for_like(cfg, B, x0, u0, n_sample, cores, local):
    """Same instances, same derivative mode on both sides: the GPU with the reference's finite-difference cost
    derivatives (the reference has no other) against the reference itself — K, k, cost after 5 trips on every instance
    of a small sample and the terminal cost on the CPU sample."""
    from ilqr_b200 import abi
    from ilqr_b200.solver import BatchILQR
    import refharness as R
    if not R.available():
        return None
    n5 = min(64, n_sample)
    kw = dict(u_min=[-cfg["limits"]], u_max=[cfg["limits"]]) if cfg["limits"] else {}
    s = BatchILQR(abi.MODEL_ACROBOT, T=cfg["T"], B=n_sample, dt=0.02, cost_deriv=abi.COST_FD, device=local, **kw)
    s.set_initial(x0[:n_sample], u0[:n_sample])
    s.iterate(5)
    K, k, c5 = s.get("K")[:n5], s.get("k")[:n5], s.get("cost")[:n5]
    s.solve()
    gpu_final = s.get("cost")
    bounds = np.linspace(0, n5, min(cores, n5) + 1).astype(int)
    jobs = [(int(bounds[i]), int(bounds[i + 1]), x0, u0, cfg["limits"], 5) for i in range(len(bounds) - 1) if bounds[i + 1] > bounds[i]]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        ref = [r for chunk in pool.map(_ref_checkpoint_worker, jobs) for r in chunk]

    def inst(a, b):  # max error over the instance's array relative to its scale, 1e-9 absolute floor
        return max(np.abs(a - b).max() - 1e-9, 0.0) / max(np.abs(b).max(), 1e-300)
    eK = np.array([inst(K[i], ref[i][0]) for i in range(n5)])
    ek = np.array([inst(k[i], ref[i][1]) for i in range(n5)])
    ec = np.array([abs(c5[i] - ref[i][2]) / abs(ref[i][2]) for i in range(n5)])
    fd_cfg = dict(cfg, cost_deriv="fd")
    _, _, _, _, term = parity_sample(fd_cfg, x0, u0, gpu_final, n_sample, cores)
    return {"cost_deriv": "fd on both sides (the reference has no other)",
            "after_5_trips": {"instances": int(n5), "K_frac_within_1e-6": float((eK <= 1e-6).mean()), "K_worst": float(eK.max()),
                              "k_frac_within_1e-6": float((ek <= 1e-6).mean()), "k_worst": float(ek.max()),
                              "cost_frac_within_1e-6": float((ec <= 1e-6).mean())},
            "terminal_cost": term}


def ours_arm(args, cfg):
    import torch
    import torch.distributed as dist
    from ilqr_b200 import abi, shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the solver has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        with shard.stdout_to_stderr():  # NCCL prints its version banner on STDOUT, ahead of the one JSON line
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()  # the communicator (and its banner) now, not at the first collective of the run
    B, T = args.batch or cfg["B"], cfg["T"]
    n, m = 4, 1
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def gather_costs(cost_d):
        return shard.gather_final_costs(cost_d, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    r = Runner(cfg, B, rank, local, dev, world, gather_costs)
    solver = r.solver

    def do_flush(stream=r.stream):
        with torch.cuda.stream(stream):
            flush_buf.zero_()

    for _ in range(max(args.warmup, 3)):
        r.step_resident()
        r.step_e2e()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: resident inputs ------------------------------------------------------
    barrier()
    wall0 = time.perf_counter()
    evs, trips, acc, rej, launches = [], 0, 0, 0, 0
    status_hist = np.zeros(5, dtype=np.int64)
    for _ in range(args.steps):
        do_flush()
        l0 = solver.launch_count
        evs.append(r.step_resident())
        launches += solver.launch_count - l0  # init_traj kernel + the solve's kernels + cost gather, per step
        solver.sync()
        trips += int(solver.get("iters").sum())
        acc += int(solver.get("n_accept").sum())
        rej += int(solver.get("n_reject").sum())
        status_hist += np.bincount(solver.get("status"), minlength=5)[:5]
    barrier()
    wall_resident = time.perf_counter() - wall0
    step_ms = [e[0].elapsed_time(e[3]) for e in evs]
    solve_ms = [e[1].elapsed_time(e[2]) for e in evs]
    # ---- timed region 2: end to end ---------------------------------------------------------------
    barrier()
    evs2 = []
    trips_e2e = 0
    for _ in range(args.steps):
        do_flush()
        evs2.append(r.step_e2e())
        torch.cuda.synchronize()
        trips_e2e += int(r.iters_h.sum())
    barrier()
    e2e_ms = [e[0].elapsed_time(e[1]) for e in evs2]
    clocks = sampler.stop() if rank == 0 else None
    # ---- extra (not the headline): fixed-N mode
    N_FIXED = 15
    fixed_ms, fixed_trips = r.fixed_n(N_FIXED, min(args.steps, 3), do_flush)

    # whole-job numbers: sum of trips over ranks / max time over ranks
    t_res, cnt = shard.reduce_step_stats([sum(step_ms), sum(e2e_ms), sum(solve_ms), fixed_ms],
                                         [trips, trips_e2e, acc, rej, fixed_trips], dev)
    per_rank = shard.gather_per_rank([sum(step_ms) / args.steps, sum(e2e_ms) / args.steps, trips / args.steps], dev)
    if rank == 0:
        value = cnt[0] / (t_res[0] * 1e-3)
        e2e = cnt[1] / (t_res[1] * 1e-3)
        sbytes = r.sbytes
        b_acc, b_rej = bytes_per_trip(n, m, T, sbytes)
        alg_bytes_per_launch = (cnt[2] * b_acc + cnt[3] * b_rej) / world / args.steps  # per GPU per solve
        solve_s = t_res[2] * 1e-3 / args.steps
        achieved = alg_bytes_per_launch / solve_s / 1e9
        traffic, traffic_src, prof = None, None, {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes of ONE solve, from ncu captures
        if os.path.exists(tpath) and world == 1:
            tj = json.load(open(tpath)).get(args.config)
            if tj and tj.get("batch") == B:
                traffic, traffic_src, prof = tj["dram_bytes_per_launch"], tj["source"], tj
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # compute-side ceiling: the measured fp64 issue rate of this GPU against the useful flops of the solve
        fma, mul, add = C.c_double(), C.c_double(), C.c_double()
        abi.load().ilqr_measure_fp64(local, C.byref(fma), C.byref(mul), C.byref(add))
        f_acc, f_rej = flops_per_trip(T, cfg["cost_deriv"] == "fd")
        useful = (cnt[2] * f_acc + cnt[3] * f_rej) / world / args.steps
        no_fma_peak = 0.5 * (mul.value + add.value)  # the library is built without FMA contraction: one flop per issue slot
        secondary = {
            "bound": "fp32 issue rate" if r.f32 else "fp64 issue rate (dependent chains: boxQP sqrt/div, sin/cos, 4-term dot products)",
            "measured_fp64_ops_per_s": {"fma": fma.value, "mul": mul.value, "add": add.value,
                                        "how": "ilqr_measure_fp64: 8 independent chains per thread, 8 x 256 threads per SM, CUDA events"},
            "peak_tflops_no_fma": no_fma_peak / 1e12, "peak_tflops_fma": 2 * fma.value / 1e12,
            "useful_mflop_per_accepted_trip": f_acc / 1e6, "useful_mflop_per_rejected_trip": f_rej / 1e6,
            "achieved_tflops": useful / solve_s / 1e12, "frac_of_no_fma_peak": useful / solve_s / no_fma_peak,
            "warp_instructions_per_trip": prof.get("warp_instructions_per_trip"), "lanes_per_instruction": prof.get("lanes_per_instruction"),
            "fp64_pipe_busy_pct": prof.get("fp64_pipe_busy_pct"),
            "note": "useful flops: SURVEY.md §8d count (a sin/cos = 1 flop); instruction figures from the committed ncu capture named in traffic_source",
        } if not r.f32 else {"bound": "fp32 issue rate", "useful_mflop_per_accepted_trip": f_acc / 1e6,
                             "achieved_tflops": useful / solve_s / 1e12}
        engine = os.environ.get("ILQR_B200_ENGINE", "phase")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": t_res[0] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if r.f32 else "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"], "batch_per_gpu": B, "T": T, "seed": SEED, "engine": engine,
                       "step": "init_traj + generate_trajectory to termination for every instance; one NCCL gather of final costs when N>1",
                       "l2": "512 MiB buffer written between timed steps (L2 flush)",
                       "cost_deriv": cfg["cost_deriv"] + (" (closed form; the reference arm can only do finite differences, see cpu_baseline.like_for_like)"
                                                          if cfg["cost_deriv"] == "analytic" else " (finite differences, as the reference)"),
                       "trips_per_step": cnt[0] / args.steps, "accepted": cnt[2] / args.steps, "rejected": cnt[3] / args.steps,
                       "exit_status_histogram": {abi.STATUS_NAMES[i]: int(v) for i, v in enumerate(status_hist) if v},
                       "wall_s_resident_region": wall_resident,
                       "fixed_n_mode": {"trips_per_instance": N_FIXED, "value": cnt[4] / (t_res[3] * 1e-3), "unit": UNIT,
                                        "note": "every instance runs exactly N trips: the engine's rate without the ragged-termination tail"}},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": t_res[1] / args.steps,
                    "h2d_bytes_per_step": int(world * (r.x0_h.numel() + r.u0_h.numel()) * sbytes),
                    "d2h_bytes_per_step": int(world * (B * sbytes + B * 4))},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm",
                         "kernel": "one ilqr_solve: the phase kernels (sweep / backward / rollout / accept / compact) of every lockstep trip, then the persistent warp kernel for the survivors"
                                   if engine != "warp" else "ilqr_warp_kernel (op_iterate)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "kernel_ms": solve_s * 1e3,
                         "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                         "limited_by": "not HBM: the fp64 arithmetic of the finite differences, boxQP and rollouts (see secondary)",
                         "secondary": secondary},
            "clocks": clocks,
        }
        if world > 1:  # what the max over ranks is the max OF: every rank solves different instances to termination
            line["per_rank"] = {"ms_per_step": [round(p[0], 3) for p in per_rank], "e2e_ms_per_step": [round(p[1], 3) for p in per_rank],
                                "trips_per_step": [int(p[2]) for p in per_rank],
                                "note": "value = sum of trips / max of ms; a rank's time is set by its own slowest instances"}
        cores = os.cpu_count() or 1
        if world == 1 and not args.no_cpu:
            n_sample = min(B, cores * args.cpu_per_core)
            kind, it, wall, used, term = parity_sample(cfg, r.x0, r.u0, r.cost_h.numpy(), n_sample, cores)
            line["cpu_baseline"] = {"value": it / wall, "unit": UNIT, "cores": used, "kind": kind,
                                    "cost_deriv": "fd (the reference has no other)",
                                    "sample": "first %d of %d instances, solved to termination, %d forked workers, %.1f s wall"
                                              % (n_sample, B, used, wall),
                                    # NB: this compares the GPU in the workload's own derivative mode with the reference's FD
                                    # mode; like_for_like below uses FD on both sides
                                    "terminal_cost_vs_gpu": term}
            try:
                line["cpu_baseline"]["like_for_like"] = like_for_like(cfg, B, r.x0, r.u0, n_sample, cores, local)
            except Exception as e:  # never lose the headline to the checker
                line["cpu_baseline"]["like_for_like"] = {"error": repr(e)}
        # the same workload through the C++ host layer (one process, BatchSolver: pinned inputs, one ncclAllGather), end to end
        cpp = os.path.join(ROOT, "ilqr_b200", "host", "_build", "bench_batch")
        if world == 1 and os.path.exists(cpp) and cfg["cost_deriv"] == "analytic" and not cfg["limits"] and not r.f32:
            try:
                vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v]
                env = dict(os.environ, CUDA_VISIBLE_DEVICES=vis[local] if local < len(vis) else str(local))  # this rank's GPU only
                out = subprocess.run([cpp, str(B), str(T), "3", "2"], capture_output=True, text=True, timeout=600, env=env)
                last = [l for l in out.stdout.strip().split("\n") if l.startswith("{")][-1]
                line["e2e_cpp_host"] = json.loads(last)
            except Exception as e:
                line["e2e_cpp_host"] = {"error": repr(e)}
        if world == 1 and not args.no_extras and args.config == "cfg2" and not args.batch:
            del r, solver
            line["extras"] = {}
            for name in ("cfg3", "cfg4", "cfg5", "cfg2_fast_fma"):
                try:
                    line["extras"][name] = run_extra(name, local, dev, do_flush, gather_costs, cores, args)
                except Exception as e:
                    line["extras"][name] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extra(name, local, dev, do_flush, gather, cores, args):
    """a short run of another BASELINE configuration inside the default line (not the headline)"""
    import torch
    from ilqr_b200 import abi
    cfg = CONFIGS[name]
    B, T = cfg["B"], cfg["T"]
    r = Runner(cfg, B, 0, local, dev, 1, gather)
    s = r.solver
    for _ in range(2):
        r.step_resident()
    torch.cuda.synchronize()
    evs, trips, acc, rej = [], 0, 0, 0
    hist = np.zeros(5, dtype=np.int64)
    steps = 2
    for _ in range(steps):
        do_flush(r.stream)
        evs.append(r.step_resident())
        s.sync()
        trips += int(s.get("iters").sum())
        acc += int(s.get("n_accept").sum())
        rej += int(s.get("n_reject").sum())
        hist += np.bincount(s.get("status"), minlength=5)[:5]
    ms = sum(e[0].elapsed_time(e[3]) for e in evs)
    solve_ms = sum(e[1].elapsed_time(e[2]) for e in evs)
    fixed_ms, fixed_trips = r.fixed_n(15, 1, lambda: do_flush(r.stream))
    b_acc, b_rej = bytes_per_trip(4, 1, T, r.sbytes)
    peak = 6553.6
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak = json.load(open(pp))["hbm_gbs"]
    achieved = (acc * b_acc + rej * b_rej) / (solve_ms * 1e-3) / 1e9
    out = {"workload": cfg["workload"], "value": trips / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "trips_per_step": trips / steps, "accepted": acc / steps, "rejected": rej / steps,
           "exit_status_histogram": {abi.STATUS_NAMES[i]: int(v // steps) for i, v in enumerate(hist) if v},
           "maxiter_fraction": float(hist[abi.EXIT_MAXITER]) / max(1, int(hist.sum())),
           "fixed_n_mode_value": fixed_trips / (fixed_ms * 1e-3),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak},
           "dtype": "f32" if r.f32 else "f64"}
    if not args.no_cpu:
        n_sample = min(B, max(cores, min(cores * 4, 64)))
        cost = s.get("cost")
        kind, it, wall, used, term = parity_sample(cfg, r.x0, r.u0, cost, n_sample, cores)
        out["cpu_reference"] = {"value": it / wall, "unit": UNIT, "cores": used, "kind": kind, "cost_deriv": "fd",
                                "terminal_cost_vs_gpu": term}
    del r, s
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override instances per GPU")
    ap.add_argument("--cpu-per-core", type=int, default=0,
                    help="CPU baseline: instances per host core in the sample (default: 128 at T=200, scaled down with the horizon)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of configs[2..4] inside the default line")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.cpu_per_core <= 0:  # ~7 s of reference work per core at T = 200; the reference's cost per iteration grows with T
        args.cpu_per_core = max(4, int(128 * (200.0 / cfg["T"]) ** 2))
    if args.impl == "reference":
        reference_arm(args, cfg)
    else:
        ours_arm(args, cfg)


if __name__ == "__main__":
    main()
