/*
 * ilqr_kernel.cuh — the solver kernel: launch geometry, argument block and `ilqr_warp_kernel`, the persistent
 * warp-per-trajectory driver around Core (ilqr_core.cuh).
 *
 * A header so that two translation units share it: ilqr_b200.cu (the built-in models, compiled by nvcc when the
 * library is built) and the run-time compiled unit of a user model (ilqr_register_model: NVRTC for sm_100a, the same
 * source with the user's struct as `Model`).  Everything here must therefore compile under NVRTC as well: no host
 * headers, no host code.
 */
#ifndef ILQR_KERNEL_CUH_
#define ILQR_KERNEL_CUH_

#include "ilqr_core.cuh"

namespace ilqr {

constexpr int kWarpsPerCta = 4;
constexpr int kThreads = kWarpsPerCta * 32;
#ifndef ILQR_MIN_BLOCKS
/* Resident CTAs per SM the register allocation must allow.  4 (126 registers, 16 warps per SM) is faster than 7
 * (72 registers, the whole BASELINE batch of 4096 resident at once) even though the batch then needs a second
 * wave: a trajectory is a serial chain, a lone warp issues in order, and under the 72-register cap ptxas sinks
 * every shared-memory load next to its use, so each dependent step pays the full load latency (measured, one
 * warp: 425 -> 308 us per loop trip; configs[1] solve 75.5 -> 67.1 ms). */
#define ILQR_MIN_BLOCKS 4
#endif

enum Op { kOpInit = 0, kOpWarm = 1, kOpIterate = 2, kOpBackwardOnce = 3, kOpRolloutOnce = 4 };

template <typename S>
struct KArgs {
  SolveParams<S> P;
  const S *x0;
  S *xs, *us, *K, *k, *Vx0, *Vxx0;
  TrajState<S> *st;
  S *slotF, *slotC, *slotCandX, *slotCandU; /* per-warp work buffers, see SlotPtrs */
  unsigned long long *queue;
  long long B;
  int op;
  int n_iters;
  S scalar; /* lambda (backward_once) or alpha (rollout_once) */
};

/* shared memory of one lane group: its scratch, then T gradient-norm terms, then (in the last 16-byte granule, which
 * nothing else may touch: a stray store into an mbarrier word corrupts its phase) the group's mbarrier */
template <class Sc, typename S>
__host__ __device__ inline size_t warp_smem_bytes(int T) {
  return ((sizeof(Sc) + (size_t)T * sizeof(S) + 15) & ~(size_t)15) + 16;
}

/* G = lanes per trajectory: 32 (one trajectory per warp) or 16 (two per warp; used for batches large enough
 * to fill the machine that way).  With G = 16 the two halves of a warp run the SAME code on their own
 * trajectories, so the phases where at most 16 lanes have work — the boxQP lane and the 11 line-search
 * candidates, more than half of all instructions — are issued once for two trajectories.  The iterate
 * operation advances both halves one loop trip at a time so they stay in lockstep; a half whose trajectory
 * has terminated pulls the next instance from the queue at the trip boundary. */
template <class Model, typename S, int CD, int G>
__global__ void __launch_bounds__(kThreads, G == 32 ? ILQR_MIN_BLOCKS : 4) ilqr_warp_kernel(const __grid_constant__ KArgs<S> a) {
  constexpr int N = Model::N, M = Model::M;
  constexpr int kGroupsPerCta = kThreads / G;
  using Ex = WarpExec<N, M, S, G>;
  using CoreT = Core<Model, S, CD, Ex>;
  using Sc = typename CoreT::Sc;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t per_group = warp_smem_bytes<Sc, S>(a.P.T);
  const int group = threadIdx.x / G;
  unsigned char *mine = smem_raw + group * per_group;
  Sc &sc = *reinterpret_cast<Sc *>(mine);
  const int T = a.P.T;
  const size_t slot = (size_t)blockIdx.x * kGroupsPerCta + group;
  SlotPtrs<S> sl;
  sl.F = a.slotF + slot * (size_t)T * (N + M) * N;
  sl.C = a.slotC ? a.slotC + slot * (size_t)T * Sc::NCF : nullptr;
  sl.cand_x = a.slotCandX + slot * (size_t)a.P.n_alpha * T * N;
  sl.cand_u = a.slotCandU + slot * (size_t)a.P.n_alpha * T * M;
  sl.gterm = reinterpret_cast<S *>(mine + sizeof(Sc));
  Ex ex;
  ex.lane = threadIdx.x & (G - 1);
  const int leader = (threadIdx.x & 31) & ~(G - 1); /* first lane of this group within the warp */
  ex.mask = G == 32 ? 0xffffffffu : (0xffffu << leader);
  ex.init_barrier(reinterpret_cast<unsigned long long *>(mine + per_group - 16));
  auto next_instance = [&]() -> long long {
    unsigned long long b = 0;
    if (ex.lane == 0) b = atomicAdd(a.queue, 1ULL);
    return (long long)__shfl_sync(ex.mask, b, leader);
  };
  auto pointers = [&](long long b) {
    TrajPtrs<S> tr;
    tr.x0 = a.x0 + b * N;
    tr.xs = a.xs + b * (size_t)(T + 1) * N;
    tr.us = a.us + b * (size_t)T * M;
    tr.K = a.K + b * (size_t)T * M * N;
    tr.k = a.k + b * (size_t)T * M;
    tr.Vx0 = a.Vx0 + b * N;
    tr.Vxx0 = a.Vxx0 + b * N * N;
    tr.st = a.st + b;
    return tr;
  };
  CoreT core(a.P, sc, ex, pointers(0), sl);
  if (G == 32 || a.op != kOpIterate) {
    for (;;) {
      const long long b = next_instance();
      if (b >= a.B) break;
      core.tr = pointers(b);
      switch (a.op) {
        case kOpInit: core.op_init(); break;
        case kOpWarm: core.op_warm_start(); break;
        case kOpIterate: core.op_iterate(a.n_iters); break;
        case kOpBackwardOnce: core.op_backward_once(a.scalar); break;
        case kOpRolloutOnce: core.op_rollout_once(a.scalar); break;
        default: break;
      }
      __syncwarp(ex.mask);
    }
  } else {
    bool has = false, drained = false;
    for (;;) {
      if (!has && !drained) {
        const long long b = next_instance();
        if (b < a.B) {
          core.tr = pointers(b);
          core.iterate_begin(a.n_iters);
          has = true;
        } else {
          drained = true;
        }
      }
      if (!__any_sync(0xffffffffu, has)) break; /* both halves meet here once per trip */
      if (has && !core.iterate_trip()) {
        core.iterate_end();
        has = false;
      }
    }
  }
#if defined(ILQR_PHASE_CLOCKS)
  if (a.op == kOpIterate && blockIdx.x == 0 && threadIdx.x == 0) {
    static const char *names[16] = {"deriv_sweep", "bw_A1_W", "bw_A2_Q", "bw_B_boxqp+V", "bw_B2_V(m>1)", "-", "bw_tile_load",
                                    "bw_tile_flush", "bw_terminal", "gnorm+test", "rollout_cand_compute", "accept_test", "commit",
                                    "lambda_sched", "roll_tile_load", "trip_head"};
    for (int i = 0; i < 16; i++)
      printf("phase %2d %-22s cycles %12llu  count %8llu  avg %8.1f\n", i, names[i], g_phase_clk[i], g_phase_cnt[i],
             g_phase_cnt[i] ? (double)g_phase_clk[i] / (double)g_phase_cnt[i] : 0.0);
  }
#endif
}

}  // namespace ilqr
#endif
