/*
 * ilqr_launch.cuh — launch geometry and dispatch (dtype, derivative mode, lane decomposition) of the solver kernel for
 * one built-in model; included by the model's translation unit only.
 */
#ifndef ILQR_LAUNCH_CUH_
#define ILQR_LAUNCH_CUH_

#include <type_traits>

#include "ilqr_host.h"
#include "ilqr_kernel.cuh"
#include "params.h"

namespace ilqr {

template <class Model, typename S, int CD, int G>
int launch_t(ilqr_handle *h, int op, int n_iters, double scalar) {
  constexpr int N = Model::N, M = Model::M;
  constexpr int kGroupsPerCta = kThreads / G;
  KArgs<S> a;
  if (make_solve_params<S>(h->desc, &a.P) != 0) return ilqr_fail(h, ILQR_E_INVALID, "bad parameters");
  a.x0 = (const S *)h->x0;
  a.xs = (S *)h->xs;
  a.us = (S *)h->us;
  a.K = (S *)h->K;
  a.k = (S *)h->k;
  a.Vx0 = (S *)h->Vx0;
  a.Vxx0 = (S *)h->Vxx0;
  a.st = (TrajState<S> *)h->st;
  a.queue = h->queue;
  a.B = h->desc.B;
  a.op = op;
  a.n_iters = n_iters;
  a.scalar = S(scalar);
  auto kern = ilqr_warp_kernel<Model, S, CD, G>;
  const size_t smem = warp_smem_bytes<typename Core<Model, S, CD, WarpExec<N, M, S, G>>::Sc, S>(h->desc.T) * kGroupsPerCta;
  if (smem > 48 * 1024) CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1) return ilqr_fail(h, ILQR_E_CUDA, "kernel does not fit on an SM");
  long long want = (h->desc.B + kGroupsPerCta - 1) / kGroupsPerCta;
  long long cap = (long long)per_sm * h->num_sms;
  const int grid = (int)(want < cap ? want : cap);
  /* the work buffers of the resident warps (Jacobian columns, FD cost derivatives, line-search candidates) */
  const long long slots = (long long)grid * kGroupsPerCta;
  if (slots > h->slots) {
    CU(h, cudaStreamSynchronize(h->stream));
    void **bufs[] = {&h->slotF, &h->slotC, &h->slotCandX, &h->slotCandU};
    const size_t T = (size_t)h->desc.T, na = (size_t)h->desc.params.n_alpha;
    const size_t per[] = {T * (N + M) * N, CD == kCostFD ? T * Scratch<N, M, S, CD>::NCF : 0, na * T * N, na * T * M};
    for (int i = 0; i < 4; i++) {
      if (*bufs[i]) CU(h, cudaFree(*bufs[i]));
      *bufs[i] = nullptr;
      if (per[i]) CU(h, cudaMalloc(bufs[i], per[i] * sizeof(S) * (size_t)slots));
    }
    h->slots = slots;
  }
  a.P.bulk_f = ((size_t)h->slotF % 16 == 0) && (((size_t)h->desc.T * (N + M) * N * sizeof(S)) % 16 == 0) &&
               (((size_t)kTileB * (N + M) * N * sizeof(S)) % 16 == 0);
  {
    constexpr size_t ncf = Scratch<N, M, S, CD>::NCF;
    a.P.bulk_c = CD == kCostFD && ((size_t)h->slotC % 16 == 0) && (((size_t)h->desc.T * ncf * sizeof(S)) % 16 == 0) &&
                 (((size_t)kTileB * ncf * sizeof(S)) % 16 == 0);
  }
  a.slotF = (S *)h->slotF;
  a.slotC = (S *)h->slotC;
  a.slotCandX = (S *)h->slotCandX;
  a.slotCandU = (S *)h->slotCandU;
  CU(h, cudaMemsetAsync(h->queue, 0, sizeof(unsigned long long), h->stream));
  kern<<<grid, kThreads, smem, h->stream>>>(a);
  CU(h, cudaGetLastError());
  h->launches++;
  return ILQR_OK;
}

template <class Model, typename S>
int launch_cd(ilqr_handle *h, int op, int n_iters, double scalar) {
#if defined(ILQR_EXPERIMENT_BUILD) /* experiments only: acrobot f64 analytic, to keep the build short */
  if (!(std::is_same<Model, Acrobot>::value && std::is_same<S, double>::value && h->desc.cost_deriv == ILQR_COST_ANALYTIC))
    return ilqr_fail(h, ILQR_E_INVALID, "experiment build: acrobot f64 analytic only");
  if constexpr (std::is_same<Model, Acrobot>::value && std::is_same<S, double>::value) {
    const bool pack16 = h->lanes == 16;
    return pack16 ? launch_t<Model, S, kCostAnalytic, 16>(h, op, n_iters, scalar) : launch_t<Model, S, kCostAnalytic, 32>(h, op, n_iters, scalar);
  } else {
    return ILQR_E_INVALID;
  }
#else
  /* Two trajectories per warp (16 lanes each) only when asked for (ILQR_B200_LANES=16: experiments and the tests of
   * that decomposition).  Round 1 chose it for batches that fill the machine; those now run on the batch-lockstep
   * phase kernels (ilqr_phases.cuh), which do the same job — issue the narrow phases once for several trajectories —
   * for every phase, so the default paths are the 32-lane kernel and the phase kernels. */
  const bool pack = h->lanes == 16;
  if (h->desc.cost_deriv == ILQR_COST_ANALYTIC)
    return pack ? launch_t<Model, S, kCostAnalytic, 16>(h, op, n_iters, scalar) : launch_t<Model, S, kCostAnalytic, 32>(h, op, n_iters, scalar);
  return pack ? launch_t<Model, S, kCostFD, 16>(h, op, n_iters, scalar) : launch_t<Model, S, kCostFD, 32>(h, op, n_iters, scalar);
#endif
}
template <class Model>
int launch_s(ilqr_handle *h, int op, int n_iters, double scalar) {
  if (h->desc.dtype == ILQR_F32) return launch_cd<Model, float>(h, op, n_iters, scalar);
  return launch_cd<Model, double>(h, op, n_iters, scalar);
}
}  // namespace ilqr
#endif
