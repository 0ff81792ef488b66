/*
 * ilqr_phase_launch.cuh — host side of the batch-lockstep engine (ilqr_phases.cuh) for one built-in model: buffers,
 * the per-trip round of launches, and the read-back that stops launching once every trajectory has terminated.
 * Included by the model's phase translation unit only.
 */
#ifndef ILQR_PHASE_LAUNCH_CUH_
#define ILQR_PHASE_LAUNCH_CUH_

#include "ilqr_host.h"
#include "ilqr_phases.cuh"
#include "params.h"

namespace ilqr {

constexpr int kPhaseCheckEvery = 8; /* trips between read-backs of the active count */

template <class Model, typename S, int CD>
int phase_prepare(ilqr_handle *h) {
  if (h->phReady) return ILQR_OK;
  constexpr size_t N = Model::N, M = Model::M, NM = N + M, NCF = NM + NM * NM;
  const size_t B = (size_t)h->desc.B, T = (size_t)h->desc.T, na = (size_t)h->desc.params.n_alpha;
  CU(h, cudaMalloc(&h->phF, B * T * NM * N * sizeof(S)));
  if (CD == kCostFD) CU(h, cudaMalloc(&h->phC, B * T * NCF * sizeof(S)));
  CU(h, cudaMalloc(&h->phCandX, B * T * na * N * sizeof(S)));
  CU(h, cudaMalloc(&h->phCandU, B * T * na * M * sizeof(S)));
  CU(h, cudaMalloc(&h->phNewcost, B * kMaxAlpha * sizeof(S)));
  CU(h, cudaMalloc((void **)&h->phAct, 2 * B * sizeof(int)));
  CU(h, cudaMalloc((void **)&h->phNact, 2 * sizeof(int)));
  CU(h, cudaMallocHost((void **)&h->phHostCount, 2 * sizeof(int)));
  for (int i = 0; i < 2; i++) CU(h, cudaEventCreateWithFlags(&h->phEvent[i], cudaEventDisableTiming));
  h->phReady = true;
  return ILQR_OK;
}

template <class Model, typename S, int CD>
int phase_iterate_t(ilqr_handle *h, int n_iters) {
  const int rc = phase_prepare<Model, S, CD>(h);
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
  a.buf.act = h->phAct;
  a.buf.n_act = h->phNact;
  a.B = h->desc.B;
  a.parity = 0;
  a.force_sweep = 1;
  cudaStream_t st = h->stream;
  CU(h, cudaMemsetAsync(h->phNact, 0, 2 * sizeof(int), st));
  {
    const unsigned blocks = (unsigned)((a.B + 255) / 256);
    phase_begin_kernel<S><<<blocks, 256, 0, st>>>(a);
    h->launches++;
  }
  long long bound = a.B; /* upper bound of the active count: it never grows */
  const int max_trips = n_iters < h->desc.params.max_iter + 1 ? n_iters : h->desc.params.max_iter + 1;
  const int na = h->desc.params.n_alpha;
  int pending = -1; /* slot of a read-back in flight */
  for (int trip = 0; trip < max_trips && bound > 0; trip++) {
    a.parity = trip & 1;
    a.force_sweep = trip == 0;
    phase_sweep_kernel<Model, S, CD><<<(unsigned)bound, kSweepThreads, 0, st>>>(a);
    phase_backward_kernel<Model, S, CD><<<(unsigned)((bound + kBackwardThreads - 1) / kBackwardThreads), kBackwardThreads, 0, st>>>(a);
    phase_rollout_kernel<Model, S, CD><<<(unsigned)((bound * na + kRolloutThreads - 1) / kRolloutThreads), kRolloutThreads, 0, st>>>(a);
    phase_accept_kernel<Model, S, CD><<<(unsigned)((bound * 32 + kAcceptThreads - 1) / kAcceptThreads), kAcceptThreads, 0, st>>>(a);
    h->launches += 4;
    if ((trip + 1) % kPhaseCheckEvery == 0) {
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
