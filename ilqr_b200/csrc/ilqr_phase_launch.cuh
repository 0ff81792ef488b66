/*
 * ilqr_phase_launch.cuh — host side of the batch-lockstep engine (ilqr_phases.cuh) for one built-in model: buffers,
 * the per-trip round of launches, and the read-back that stops launching once every trajectory has terminated.
 * Included by the model's phase translation unit only.
 */
#ifndef ILQR_PHASE_LAUNCH_CUH_
#define ILQR_PHASE_LAUNCH_CUH_

#include <stdlib.h>

#include "ilqr_variant.h"
#include "ilqr_host.h"
#include "ilqr_phases.cuh"
#include "params.h"

namespace ilqr {

constexpr int kPhaseCheckEvery = 8; /* trips between read-backs of the active count (4 for small batches) */
/* Running trajectories at or below which the lockstep rounds stop and the persistent warp-per-trajectory kernel
 * finishes the solve.  A lockstep round costs its latency floor (~0.45 ms: the slowest trajectory's backward pass and
 * rollout, one after the other) however few trajectories are left, and every trajectory waits for the slowest; the
 * persistent kernel advances each survivor at its own pace (~0.25 ms per trip) once they all fit the machine at once
 * (2368 resident warps).  Measured on BASELINE configs[1] (gpurun_out r2h/r2i): 59.7 ms lockstep only, 48.8 ms with
 * the hand-over at 3200, 53.3 ms persistent kernel only. */
constexpr long long kPhaseHandover = 3200;
/* Running trajectories above which the line search is staged (ilqr_phases.cuh: PArgs::stage), and the size of the
 * first stage.  62 % of the line searches of the synthetic batch accept one of the first four candidates (23 / 16 /
 * 14 / 9 %; 28 % reject all eleven), so staging rolls out 40 % fewer candidates and writes a third of the candidate
 * bytes (configs[4] shard: 870 GB of DRAM traffic per solve instead of 1203).  OFF by default (environment
 * ILQR_B200_STAGE_MIN turns it on): it was slower at every size measured — configs[4] shard 11.1 -> 9.5 M it/s (k = 4),
 * 10.1 (k = 6); configs[2] 5.7 -> 3.8; configs[1] 3.50 -> 3.11 (gpurun_out/st3, profiles/experiments/README.md).  A
 * rollout launch costs its 200-step chain however few candidates it carries, so two stages pay it twice; the re-roll
 * of the steps accepted in stage 2 is one thread per trajectory at DRAM latency per timestep (65 ms per solve for 0.6 %
 * of the instructions); and the backward phase lost 16 % on an active list that is no longer in ascending order. */
constexpr long long kPhaseStageMin = 1LL << 62;
constexpr int kPhaseStageK = 4;
/* Running trajectories above which the next active list is built by order-preserving compaction (ilqr_phases.cuh:
 * phase_compact_kernel; one more launch per trip, ~10 us) instead of atomic append.  The 32 trajectories of a warp of
 * the thread-per-trajectory kernels are then neighbours in memory (1 MB of Jacobians instead of a sample of ~80 MB:
 * TLB reach and open DRAM pages; ncu long-scoreboard stalls of the backward phase 0.73 -> 0.26 per issue at trip 2 and
 * its time over a solve 173 -> 113 ms): configs[4] shard 11.0 -> 12.4 M it/s, configs[2] 5.74 -> 5.96, configs[3]
 * +1 %, configs[1] +0.5 % (gpurun_out/ord1, acc1). */
constexpr long long kPhaseOrderedMin = 2048;
/* Running trajectories above which the line search stores no candidates (cost-only rollouts + one re-roll of the
 * accepted candidate, ilqr_phases.cuh: rollout_task).  OFF by default (environment ILQR_B200_REROLL_MIN turns it on):
 * it removes the candidate buffers (88 KB per trajectory at T = 200) and 40 % of the DRAM traffic of configs[4]
 * (483 of 1203 GB per solve are candidate stores, ten of eleven never read), but the extra rollout — one thread per
 * accepted trajectory, a 200-step chain at low occupancy — costs more time than the stores: configs[4] shard 11.0 ->
 * 9.2 M it/s, configs[2] 5.7 -> 3.4 M it/s (gpurun_out/r2t).  For batches whose candidate buffers would not fit. */
constexpr long long kPhaseRerollMin = 1LL << 62;

inline void phase_release(ilqr_handle *h) {
  void **bufs[] = {&h->phF, &h->phC, &h->phCandX, &h->phCandU, &h->phNewcost, &h->phGterm, (void **)&h->phAct, (void **)&h->phNact};
  for (void **b : bufs) {
    if (*b) cudaFree(*b);
    *b = nullptr;
  }
  h->phReady = false;
}

/* The per-trajectory buffers of the phase kernels.  Returns kPhaseNoMemory (and releases what it had taken) when the
 * batch's stored derivatives do not fit beside its trajectories: the caller then runs the persistent kernel, whose
 * work buffers are per resident warp, not per trajectory — same results, any batch the trajectories themselves fit. */
constexpr int kPhaseNoMemory = 1000;
template <class Model, typename S, int CD>
int phase_prepare(ilqr_handle *h) {
  if (h->phReady) return ILQR_OK;
  constexpr size_t N = Model::N, M = Model::M, NM = N + M, NCF = NM + NM * NM;
  const size_t B = (size_t)h->desc.B, T = (size_t)h->desc.T;
  if (B > (size_t)1 << 30) return kPhaseNoMemory; /* the active lists index trajectories with 32-bit integers */
  bool ok = cudaMalloc(&h->phF, B * T * NM * N * sizeof(S)) == cudaSuccess;
  if (ok && CD == kCostFD) ok = cudaMalloc(&h->phC, B * T * NCF * sizeof(S)) == cudaSuccess;
  ok = ok && cudaMalloc(&h->phNewcost, B * kMaxAlpha * sizeof(S)) == cudaSuccess;
  ok = ok && cudaMalloc(&h->phGterm, B * T * sizeof(S)) == cudaSuccess;
  ok = ok && cudaMalloc((void **)&h->phAct, 5 * B * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMalloc((void **)&h->phNact, 4 * sizeof(int)) == cudaSuccess;
  if (!ok) {
    const cudaError_t e = cudaGetLastError(); /* also clears it */
    phase_release(h);
    if (e == cudaErrorMemoryAllocation) return kPhaseNoMemory;
    return ilqr_fail(h, ILQR_E_CUDA, std::string("phase buffers: ") + cudaGetErrorString(e));
  }
  if (!h->phHostCount) CU(h, cudaMallocHost((void **)&h->phHostCount, 2 * sizeof(int)));
  for (int i = 0; i < 2; i++)
    if (!h->phEvent[i]) CU(h, cudaEventCreateWithFlags(&h->phEvent[i], cudaEventDisableTiming));
  h->phReady = true;
  return ILQR_OK;
}

template <class Model, typename S, int CD>
int phase_iterate_t(ilqr_handle *h, int n_iters) {
  const int rc = phase_prepare<Model, S, CD>(h);
  if (rc == kPhaseNoMemory || getenv("ILQR_B200_TEST_NO_PHASE_MEMORY")) { /* (the variable: the tests' way to take this branch) */
    phase_release(h);
    h->engine_warp = true; /* for the rest of this handle's life */
    return h->desc.model_id == ILQR_MODEL_ACROBOT ? ILQR_ENTRY(ilqr_launch_acrobot)(h, kOpIterate, n_iters, 0.0)
                                                  : ILQR_ENTRY(ilqr_launch_double_integrator)(h, kOpIterate, n_iters, 0.0);
  }
  if (rc != ILQR_OK) return rc;
  PArgs<S> a;
  if (make_solve_params<S>(h->desc, &a.P) != 0) return ilqr_fail(h, ILQR_E_INVALID, "bad parameters");
  a.x0 = (const S *)h->x0;
  a.xs = (S *)h->xs;
  a.us = (S *)h->us;
  a.K = (S *)h->K;
  a.k = (S *)h->k;
  a.Vx0 = (S *)h->Vx0;
  a.Vxx0 = (S *)h->Vxx0;
  a.st = (TrajState<S> *)h->st;
  a.buf.F = (S *)h->phF;
  a.buf.C = (S *)h->phC;
  a.buf.cand_x = (S *)h->phCandX;
  a.buf.cand_u = (S *)h->phCandU;
  a.buf.newcost = (S *)h->phNewcost;
  a.buf.gterm = (S *)h->phGterm;
  a.buf.act = h->phAct;
  a.buf.n_act = h->phNact;
  a.B = h->desc.B;
  a.parity = 0;
  a.reroll = 0;
  a.force_sweep = 1;
  cudaStream_t st = h->stream;
  CU(h, cudaMemsetAsync(h->phNact, 0, 4 * sizeof(int), st));
  {
    const unsigned blocks = (unsigned)((a.B + 255) / 256);
    phase_begin_kernel<S><<<blocks, 256, 0, st>>>(a);
    h->launches++;
  }
  long long bound = a.B; /* upper bound of the active count: it never grows */
  /* the head of a trip (sweep + backward) runs one thread per trajectory when the active set can fill the machine that
   * way, 8 lanes per trajectory below that, optionally one warp per trajectory (ilqr_phases.cuh).  The crossover,
   * measured with ordered active lists (gpurun_out/rm1): closed-form cost derivatives — configs[4] shard 12.59 /
   * 12.66 / 12.71 / 12.66 M it/s at 6144 / 8192 / 12288 / 24576, configs[1] (4096 trajectories) 3.10 M one thread
   * each against 3.61 M; finite-difference cost derivatives (each lane also streams its row of C) — configs[3]
   * (8192 trajectories) 2.55 M it/s at 6144 against 2.46 M at 8192 and above */
  long long warp_pre_max = 0, rows_max = CD == kCostFD ? 6144 : 12288;
  if (const char *e = getenv("ILQR_B200_WARP_PRE_MAX")) warp_pre_max = atoll(e);
  if (const char *e = getenv("ILQR_B200_ROWS_MAX")) rows_max = atoll(e);
  int rows_gpw = 4;
  if (const char *e = getenv("ILQR_B200_ROWS_GPW")) rows_gpw = atoi(e);
  /* below this many running trajectories the lockstep rounds stop and the persistent warp-per-trajectory kernel
   * finishes the solve: every trajectory then advances at its own pace instead of the pace of the slowest */
  long long handover = kPhaseHandover;
  if (const char *e = getenv("ILQR_B200_HANDOVER")) handover = atoll(e);
  int check_every = h->desc.B <= 16384 ? 4 : kPhaseCheckEvery;
  /* above this many running trajectories the line search keeps no candidates and re-rolls the accepted one
   * (ilqr_phases.cuh: rollout_task) */
  long long reroll_min = kPhaseRerollMin;
  if (const char *e = getenv("ILQR_B200_REROLL_MIN")) reroll_min = atoll(e);
  /* above this many running trajectories the line search is staged: first stage_k candidates, then the rest */
  long long stage_min = kPhaseStageMin;
  if (const char *e = getenv("ILQR_B200_STAGE_MIN")) stage_min = atoll(e);
  /* above this many running trajectories the next active list keeps the order of this one (phase_compact_kernel) */
  long long ordered_min = kPhaseOrderedMin;
  if (const char *e = getenv("ILQR_B200_ORDERED_MIN")) ordered_min = atoll(e);
  int stage_k = kPhaseStageK;
  if (const char *e = getenv("ILQR_B200_STAGE_K")) stage_k = atoi(e);
  if (const char *e = getenv("ILQR_B200_CHECK_EVERY")) check_every = atoi(e) > 0 ? atoi(e) : check_every;
  constexpr int N = Model::N, M = Model::M;
  const size_t pre_smem = warp_smem_bytes<typename Core<Model, S, CD, WarpExec<N, M, S, 32>>::Sc, S>(h->desc.T) * kWarpsPerCta;
  if (pre_smem > 48 * 1024)
    CU(h, cudaFuncSetAttribute(phase_pre_warp_kernel<Model, S, CD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem));
  a.P.bulk_f = ((size_t)h->phF % 16 == 0) && (((size_t)h->desc.T * (N + M) * N * sizeof(S)) % 16 == 0) &&
               (((size_t)kTileB * (N + M) * N * sizeof(S)) % 16 == 0);
  {
    constexpr size_t ncf = Scratch<N, M, S, CD>::NCF;
    a.P.bulk_c = CD == kCostFD && ((size_t)h->phC % 16 == 0) && (((size_t)h->desc.T * ncf * sizeof(S)) % 16 == 0) &&
                 (((size_t)kTileB * ncf * sizeof(S)) % 16 == 0);
  }
  const int max_trips = n_iters < h->desc.params.max_iter + 1 ? n_iters : h->desc.params.max_iter + 1;
  const int na = h->desc.params.n_alpha;
  int pending = -1; /* slot of a read-back in flight */
  int trip = 0;
  for (; trip < max_trips && bound > handover; trip++) {
    a.parity = trip & 1;
    a.force_sweep = trip == 0;
    if (bound <= warp_pre_max) { /* few trajectories: 32 lanes each (sweep + backward in one kernel) */
      phase_pre_warp_kernel<Model, S, CD><<<(unsigned)((bound + kWarpsPerCta - 1) / kWarpsPerCta), kThreads, pre_smem, st>>>(a);
      h->launches += 1;
    } else if (bound <= rows_max) { /* a few thousand: 8 lanes each, one row of the Q-function per lane */
      phase_sweep_kernel<Model, S, CD><<<(unsigned)bound, kSweepThreads, 0, st>>>(a);
      if (rows_gpw == 1) {
        constexpr int per_cta = (kRowThreads / 32) * 1;
        phase_backward_rows_kernel<Model, S, CD, 1><<<(unsigned)((bound + per_cta - 1) / per_cta), kRowThreads, 0, st>>>(a);
      } else if (rows_gpw == 2) {
        constexpr int per_cta = (kRowThreads / 32) * 2;
        phase_backward_rows_kernel<Model, S, CD, 2><<<(unsigned)((bound + per_cta - 1) / per_cta), kRowThreads, 0, st>>>(a);
      } else {
        constexpr int per_cta = (kRowThreads / 32) * 4;
        phase_backward_rows_kernel<Model, S, CD, 4><<<(unsigned)((bound + per_cta - 1) / per_cta), kRowThreads, 0, st>>>(a);
      }
      h->launches += 2;
    } else {
      phase_sweep_kernel<Model, S, CD><<<(unsigned)bound, kSweepThreads, 0, st>>>(a);
      phase_backward_kernel<Model, S, CD><<<(unsigned)((bound + kBackwardThreads - 1) / kBackwardThreads), kBackwardThreads, 0, st>>>(a);
      h->launches += 2;
    }
    a.reroll = bound > reroll_min || h->phNoCand;
    if (!a.reroll && !h->phCandX) { /* the candidate buffers, on first use */
      const size_t B = (size_t)h->desc.B, T = (size_t)h->desc.T;
      const bool ok = !getenv("ILQR_B200_TEST_NO_CAND_MEMORY") && cudaMalloc(&h->phCandX, B * T * na * N * sizeof(S)) == cudaSuccess &&
                      cudaMalloc(&h->phCandU, B * T * na * M * sizeof(S)) == cudaSuccess;
      if (!ok) { /* they do not fit: cost-only rollouts and one re-roll of the accepted candidate instead — same results */
        (void)cudaGetLastError();
        if (h->phCandX) cudaFree(h->phCandX);
        h->phCandX = h->phCandU = nullptr;
        h->phNoCand = a.reroll = true;
      }
      a.buf.cand_x = (S *)h->phCandX;
      a.buf.cand_u = (S *)h->phCandU;
    }
    /* the line search: all candidates at once, or — large active sets — the first stage_k, then the rest (cost only)
     * for the trajectories where none of those passed (ilqr_phases.cuh: PArgs::stage) */
    const bool staged = bound > stage_min && stage_k > 0 && stage_k < na;
    a.ordered = !staged && bound > ordered_min;
    for (int stage = staged ? 1 : 0; stage <= (staged ? 2 : 0); stage++) {
      a.stage = stage;
      a.cand_lo = stage == 2 ? stage_k : 0;
      a.cand_hi = stage == 1 ? stage_k : na;
      a.keep = !a.reroll && stage != 2;
      const long long nthr = bound * (a.cand_hi - a.cand_lo); /* stage 2: an upper bound, its list is on the device */
      if (a.keep)
        phase_rollout_kernel<Model, S, CD, Phases<Model, S, CD>::kToCand>
            <<<(unsigned)((nthr + kRolloutThreads - 1) / kRolloutThreads), kRolloutThreads, 0, st>>>(a);
      else
        phase_rollout_kernel<Model, S, CD, Phases<Model, S, CD>::kCostOnly>
            <<<(unsigned)((nthr + kRolloutThreads - 1) / kRolloutThreads), kRolloutThreads, 0, st>>>(a);
      phase_accept_kernel<Model, S, CD><<<(unsigned)((bound * 32 + kAcceptThreads - 1) / kAcceptThreads), kAcceptThreads, 0, st>>>(a);
      h->launches += 2;
    }
    if (a.ordered) {
      phase_compact_kernel<S><<<1, kCompactThreads, 0, st>>>(a);
      h->launches += 1;
    }
    if (a.reroll || staged) { /* the accepted steps whose rollouts were not kept (list 3; the grid is an upper bound) */
      phase_commit_kernel<Model, S, CD><<<(unsigned)((bound + kRolloutThreads - 1) / kRolloutThreads), kRolloutThreads, 0, st>>>(a);
      h->launches += 1;
    }
    if ((trip + 1) % check_every == 0) {
      if (pending >= 0) { /* the count of kPhaseCheckEvery trips ago: by now it has almost always arrived */
        CU(h, cudaEventSynchronize(h->phEvent[pending]));
        bound = h->phHostCount[pending];
      }
      const int slot = pending < 0 ? 0 : pending ^ 1;
      CU(h, cudaMemcpyAsync(&h->phHostCount[slot], h->phNact + (a.parity ^ 1), sizeof(int), cudaMemcpyDeviceToHost, st));
      CU(h, cudaEventRecord(h->phEvent[slot], st));
      pending = slot;
    }
  }
  CU(h, cudaGetLastError());
  if (bound > 0 && trip < max_trips) { /* hand the survivors to the persistent kernel for their remaining trips */
    const int left = n_iters - trip;
    return h->desc.model_id == ILQR_MODEL_ACROBOT ? ILQR_ENTRY(ilqr_launch_acrobot)(h, kOpIterate, left, 0.0)
                                                  : ILQR_ENTRY(ilqr_launch_double_integrator)(h, kOpIterate, left, 0.0);
  }
  return ILQR_OK;
}

template <class Model>
int phase_iterate(ilqr_handle *h, int n_iters) {
  const bool fd = h->desc.cost_deriv == ILQR_COST_FD;
  if (h->desc.dtype == ILQR_F32)
    return fd ? phase_iterate_t<Model, float, kCostFD>(h, n_iters) : phase_iterate_t<Model, float, kCostAnalytic>(h, n_iters);
  return fd ? phase_iterate_t<Model, double, kCostFD>(h, n_iters) : phase_iterate_t<Model, double, kCostAnalytic>(h, n_iters);
}

}  // namespace ilqr
#endif
