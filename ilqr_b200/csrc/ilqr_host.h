/*
 * ilqr_host.h — what the translation units of libilqr_b200.so share on the host side: the handle, error plumbing, and
 * the per-model launch entry points.  The kernels of each built-in model are compiled in a translation unit of their
 * own (ilqr_model_acrobot.cu, ilqr_model_double_integrator.cu) so that the library builds in parallel; ilqr_b200.cu
 * holds the C ABI, the run-time compiled user-model path and everything that does not depend on a model.
 */
#ifndef ILQR_HOST_H_
#define ILQR_HOST_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/ilqr_b200.h"

struct ilqr_handle {
  ilqr_desc desc;
  int n = 0, m = 0;
  size_t ssize = 8;
  cudaStream_t stream = nullptr;
  void *x0 = nullptr, *xs = nullptr, *us = nullptr, *K = nullptr, *k = nullptr, *Vx0 = nullptr, *Vxx0 = nullptr,
       *st = nullptr, *tmp = nullptr;
  void *slotF = nullptr, *slotC = nullptr, *slotCandX = nullptr, *slotCandU = nullptr; /* per resident warp */
  long long slots = 0;
  int lanes = 0; /* 0: choose by batch size; 16 / 32: forced (environment ILQR_B200_LANES, for experiments and tests) */
  /* batch-lockstep phase engine (ilqr_phases.cuh): per-trajectory work arrays, active lists, trip bookkeeping */
  bool engine_warp = false; /* ILQR_FLAG_ENGINE_WARP or environment ILQR_B200_ENGINE=warp */
  void *phF = nullptr, *phC = nullptr, *phCandX = nullptr, *phCandU = nullptr, *phNewcost = nullptr, *phGterm = nullptr;
  int *phAct = nullptr, *phNact = nullptr;
  int *phHostCount = nullptr; /* pinned: active-list lengths read back while the trips run */
  cudaEvent_t phEvent[2] = {nullptr, nullptr};
  bool phReady = false;
  bool phNoCand = false; /* the candidate buffers did not fit: the line search re-rolls the accepted candidate */
  unsigned long long *queue = nullptr;
  int num_sms = 0;
  int64_t launches = 0;
  bool initialised = false;
  std::string err;
};


/* records msg as the handle's (or, h == NULL, the thread's ilqr_create) last error and returns code */
int ilqr_fail(ilqr_handle *h, int code, const std::string &msg);
#define CU(h, call)                                                                                     \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return ilqr_fail(h, ILQR_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));             \
  } while (0)

/* one launch of the solver kernel of a built-in model for the handle's dtype / derivative mode / batch size */
int ilqr_launch_acrobot(ilqr_handle *h, int op, int n_iters, double scalar);
int ilqr_launch_double_integrator(ilqr_handle *h, int op, int n_iters, double scalar);
/* up to n_iters loop trips for every running instance on the phase engine (ilqr_phase_launch.cuh) */
int ilqr_phase_iterate_acrobot(ilqr_handle *h, int n_iters);
int ilqr_phase_iterate_double_integrator(ilqr_handle *h, int n_iters);
/* the same four from the translation units built with fused multiply-add contraction (ilqr_variant.h) */
int ilqr_launch_acrobot_fma(ilqr_handle *h, int op, int n_iters, double scalar);
int ilqr_launch_double_integrator_fma(ilqr_handle *h, int op, int n_iters, double scalar);
int ilqr_phase_iterate_acrobot_fma(ilqr_handle *h, int n_iters);
int ilqr_phase_iterate_double_integrator_fma(ilqr_handle *h, int n_iters);

#endif
