/*
 * ilqr_b200.cu — kernels and the extern "C" ABI (include/ilqr_b200.h) of libilqr_b200.so.
 *
 * One persistent warp per trajectory: every warp of the grid pulls problem instances from an
 * atomic queue and runs the whole operation for that instance (ilqr_core.cuh) with its working
 * set in its private slice of shared memory; warps never synchronise with each other.  The grid
 * is sized to fill the 148 SMs at the kernel's occupancy, so ragged per-trajectory iteration
 * counts are absorbed by the queue.  One launch per ABI call (`ilqr_solve` = one launch).
 *
 * HBM layout (per handle, scalar type S = f64 or f32), trajectory-major and contiguous in t so a
 * warp streams its own trajectory with coalesced tile copies; identical to the host layout of
 * the ABI, so set/get are plain copies:
 *     x0 [B][n]   xs [B][T+1][n]   us [B][T][m]   K [B][T][m][n]   k [B][T][m]
 *     Vx0 [B][n]  Vxx0 [B][n][n]   st [B] (TrajState: cost, lambda, dlambda, dV, counters ...)
 *
 * The kernels of the built-in models are compiled in translation units of their own (ilqr_model_*.cu through
 * ilqr_launch.cuh); this file holds the C ABI, the run-time compiled user-model path (ilqr_rtc.inc) and everything
 * that does not depend on a model.
 *
 * There is no CPU path in this library: every entry point that computes launches a kernel.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <type_traits>

#include "../../include/ilqr_synth.h"
#include "ilqr_host.h"
#include "ilqr_kernel.cuh"
#include "params.h"

using namespace ilqr;

namespace {

/* per-trajectory scalars out of the state records, one thread per trajectory */
template <typename S>
__global__ void ilqr_gather_kernel(const TrajState<S> *st, long long B, int field, void *out) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= B) return;
  const TrajState<S> &s = st[b];
  S *fo = (S *)out;
  int32_t *io = (int32_t *)out;
  switch (field) {
    case ILQR_F_COST: fo[b] = s.cost; break;
    case ILQR_F_DV: fo[2 * b] = s.dV0; fo[2 * b + 1] = s.dV1; break;
    case ILQR_F_LAMBDA: fo[b] = s.lam; break;
    case ILQR_F_DLAMBDA: fo[b] = s.dlam; break;
    case ILQR_F_GNORM: fo[b] = s.gnorm; break;
    case ILQR_F_ITERS: io[b] = s.trips; break;
    case ILQR_F_STATUS: io[b] = s.status; break;
    case ILQR_F_ALPHA_INDEX: io[b] = s.alpha_index; break;
    case ILQR_F_N_ACCEPT: io[b] = s.n_accept; break;
    case ILQR_F_N_REJECT: io[b] = s.n_reject; break;
    case ILQR_F_N_BACKWARD: io[b] = s.n_backward; break;
    case ILQR_F_DIVERGE: io[b] = s.diverge; break;
    default: break;
  }
}

/* iLQR::generate_trajectory() re-entered (src/ilqr_core.cpp:88-102): iter = 0, flgChange = true, running again */
template <typename S>
__global__ void ilqr_resume_kernel(TrajState<S> *st, long long B) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= B) return;
  st[b].iter = 0;
  st[b].flg_change = 1;
  st[b].status = kRunning;
}

/* fp64 issue-rate probe (ilqr_measure_fp64): 8 independent chains per thread, enough warps to fill every scheduler */
template <int OP>
__global__ void ilqr_fp64_probe_kernel(double a, double b, int n, double *sink) {
  double x[8];
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = a + threadIdx.x + k;
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (OP == 0) x[k] = fma(x[k], b, a);
      if (OP == 1) x[k] = x[k] * b;
      if (OP == 2) x[k] = x[k] + b;
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += x[k];
  if (s == 12345.678) sink[0] = s; /* never true: keeps the chains alive */
}

thread_local std::string g_create_error;

}  // namespace

int ilqr_fail(ilqr_handle *h, int code, const std::string &msg) {
  if (h) h->err = msg;
  else g_create_error = msg;
  return code;
}

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

int fail(ilqr_handle *h, int code, const std::string &msg) { return ilqr_fail(h, code, msg); }

}  // namespace
#include "ilqr_rtc.inc"
namespace {

int launch(ilqr_handle *h, int op, int n_iters, double scalar) {
  if (h->desc.model_id >= ILQR_MODEL_USER_BASE)
    return h->desc.dtype == ILQR_F32 ? launch_user<float>(h, op, n_iters, scalar) : launch_user<double>(h, op, n_iters, scalar);
  if (h->desc.flags & ILQR_FLAG_FAST_FMA)
    return h->desc.model_id == ILQR_MODEL_ACROBOT ? ilqr_launch_acrobot_fma(h, op, n_iters, scalar)
                                                  : ilqr_launch_double_integrator_fma(h, op, n_iters, scalar);
  if (h->desc.model_id == ILQR_MODEL_ACROBOT) return ilqr_launch_acrobot(h, op, n_iters, scalar);
  return ilqr_launch_double_integrator(h, op, n_iters, scalar);
}

size_t state_size(const ilqr_handle *h) {
  return h->desc.dtype == ILQR_F32 ? sizeof(TrajState<float>) : sizeof(TrajState<double>);
}

}  // namespace

extern "C" {

int ilqr_default_params(ilqr_params *p) {
  if (!p) return ILQR_E_INVALID;
  default_params(p);
  return ILQR_OK;
}

int ilqr_model_info(int32_t model_id, int32_t *n, int32_t *m, double *u_min, double *u_max) {
  int nn, mm;
  if (model_info(model_id, &nn, &mm, u_min, u_max) != 0) return ILQR_E_INVALID;
  if (n) *n = nn;
  if (m) *m = mm;
  return ILQR_OK;
}

int ilqr_register_model(const char *struct_name, const char *cuda_source, int32_t n, int32_t m, const double *u_min,
                        const double *u_max, int32_t *model_id) {
  if (!struct_name || !cuda_source || !u_min || !u_max || !model_id) return fail(nullptr, ILQR_E_INVALID, "null argument");
  if (n < 1 || n > ILQR_MAX_N || m < 1 || m > ILQR_MAX_M) return fail(nullptr, ILQR_E_INVALID, "n or m out of range");
  std::lock_guard<std::mutex> lock(g_rtc_mutex);
  UserModel um;
  um.name = struct_name;
  um.source = cuda_source;
  um.n = n;
  um.m = m;
  for (int j = 0; j < m; j++) {
    um.u_min[j] = u_min[j];
    um.u_max[j] = u_max[j];
  }
  g_user_models.push_back(um);
  g_user_model_info = [](int id, int *pn, int *pm, double *lo, double *hi) -> int {
    std::lock_guard<std::mutex> guard(g_rtc_mutex); /* never called with the lock held */
    const UserModel *u = user_model(id);
    if (!u) return -1;
    *pn = u->n;
    *pm = u->m;
    for (int j = 0; j < u->m; j++) {
      if (lo) lo[j] = u->u_min[j];
      if (hi) hi[j] = u->u_max[j];
    }
    return 0;
  };
  *model_id = ILQR_MODEL_USER_BASE + (int)g_user_models.size() - 1;
  return ILQR_OK;
}

int ilqr_compile_model(int32_t model_id, int32_t dtype, int32_t cost_deriv, char *log, size_t log_bytes) {
  UserModel um;
  {
    std::lock_guard<std::mutex> lock(g_rtc_mutex);
    const UserModel *u = user_model(model_id);
    if (!u) return fail(nullptr, ILQR_E_INVALID, "unknown model_id");
    um = *u;
  }
  std::vector<char> cubin;
  std::string lowered, text;
  const int rc = rtc_compile(um, dtype, cost_deriv, &cubin, &lowered, &text);
  if (log && log_bytes) {
    const size_t k = text.size() < log_bytes - 1 ? text.size() : log_bytes - 1;
    memcpy(log, text.data(), k);
    log[k] = '\0';
  }
  if (rc != ILQR_OK) fail(nullptr, rc, "user model did not compile");
  return rc;
}

const char *ilqr_last_error(const ilqr_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ilqr_destroy(ilqr_handle *h) {
  if (!h) return ILQR_OK;
  {
    DeviceGuard g(h->desc.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void *bufs[] = {h->x0, h->xs, h->us, h->K, h->k, h->Vx0, h->Vxx0, h->st, h->tmp, h->queue,
                    h->slotF, h->slotC, h->slotCandX, h->slotCandU,
                    h->phF, h->phC, h->phCandX, h->phCandU, h->phNewcost, h->phGterm, h->phAct, h->phNact};
    for (void *b : bufs)
      if (b) cudaFree(b);
    if (h->phHostCount) cudaFreeHost(h->phHostCount);
    for (cudaEvent_t e : h->phEvent)
      if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
  }
  delete h;
  return ILQR_OK;
}

int ilqr_create(const ilqr_desc *desc, ilqr_handle **out) {
  if (!desc || !out) return fail(nullptr, ILQR_E_INVALID, "null argument");
  *out = nullptr;
  int n, m;
  if (model_info(desc->model_id, &n, &m, nullptr, nullptr) != 0) return fail(nullptr, ILQR_E_INVALID, "unknown model_id");
  if (desc->dtype != ILQR_F64 && desc->dtype != ILQR_F32) return fail(nullptr, ILQR_E_INVALID, "unknown dtype");
  if (desc->cost_deriv != ILQR_COST_FD && desc->cost_deriv != ILQR_COST_ANALYTIC)
    return fail(nullptr, ILQR_E_INVALID, "unknown cost_deriv");
  if (desc->B < 1 || desc->T < 1) return fail(nullptr, ILQR_E_INVALID, "B and T must be positive");
  if (!(desc->dt > 0)) return fail(nullptr, ILQR_E_INVALID, "dt must be positive");
  if ((desc->flags & ~ILQR_FLAG_ALL) != 0 || desc->reserved1 != 0) return fail(nullptr, ILQR_E_INVALID, "unknown flags");
  if ((desc->flags & ILQR_FLAG_ANALYTIC_DYN) && desc->model_id >= ILQR_MODEL_USER_BASE)
    return fail(nullptr, ILQR_E_INVALID, "ILQR_FLAG_ANALYTIC_DYN: built-in models only");
  if ((desc->flags & ILQR_FLAG_FAST_FMA) && desc->model_id >= ILQR_MODEL_USER_BASE)
    return fail(nullptr, ILQR_E_INVALID, "ILQR_FLAG_FAST_FMA: built-in models only (user models are compiled without contraction)");
  {
    SolveParams<double> chk;
    if (make_solve_params<double>(*desc, &chk) != 0) return fail(nullptr, ILQR_E_INVALID, "bad solver parameters");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(nullptr, ILQR_E_CUDA, "no CUDA device: this library has no CPU path");
  if (desc->device < 0 || desc->device >= ndev) return fail(nullptr, ILQR_E_INVALID, "device ordinal out of range");
  ilqr_handle *h = new (std::nothrow) ilqr_handle;
  if (!h) return fail(nullptr, ILQR_E_NOMEM, "out of host memory");
  h->desc = *desc;
  if (const char *e = getenv("ILQR_B200_LANES")) h->lanes = atoi(e) == 16 ? 16 : (atoi(e) == 32 ? 32 : 0);
  h->engine_warp = (desc->flags & ILQR_FLAG_ENGINE_WARP) != 0;
  if (const char *e = getenv("ILQR_B200_ENGINE")) h->engine_warp = strcmp(e, "warp") == 0 || (h->engine_warp && strcmp(e, "phase") != 0);
  if (h->lanes != 0) h->engine_warp = true; /* a forced lane decomposition is a property of the warp kernel */
  h->n = n;
  h->m = m;
  h->ssize = desc->dtype == ILQR_F32 ? 4 : 8;
  DeviceGuard g(desc->device);
  const size_t B = (size_t)desc->B, T = (size_t)desc->T, s = h->ssize;
  struct {
    void **p;
    size_t bytes;
  } allocs[] = {{&h->x0, B * n * s},          {&h->xs, B * (T + 1) * n * s}, {&h->us, B * T * m * s},
                {&h->K, B * T * m * n * s},   {&h->k, B * T * m * s},        {&h->Vx0, B * n * s},
                {&h->Vxx0, B * n * n * s},    {&h->st, B * state_size(h)},   {&h->tmp, B * 2 * 8},
                {(void **)&h->queue, sizeof(unsigned long long)}};
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  for (auto &al : allocs) {
    if (e != cudaSuccess) break;
    e = cudaMalloc(al.p, al.bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(*al.p, 0, al.bytes, h->stream);
  }
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, desc->device);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    const int code = e == cudaErrorMemoryAllocation ? ILQR_E_NOMEM : ILQR_E_CUDA;
    fail(nullptr, code, std::string("ilqr_create: ") + cudaGetErrorString(e));
    ilqr_destroy(h);
    return code;
  }
  *out = h;
  return ILQR_OK;
}

int ilqr_set_initial(ilqr_handle *h, const void *x0, const void *u0, int on_device) {
  if (!h || !x0 || !u0) return fail(h, ILQR_E_INVALID, "null argument");
  DeviceGuard g(h->desc.device);
  const size_t B = (size_t)h->desc.B, T = (size_t)h->desc.T, s = h->ssize;
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CU(h, cudaMemcpyAsync(h->x0, x0, B * h->n * s, kind, h->stream));
  CU(h, cudaMemcpyAsync(h->us, u0, B * T * h->m * s, kind, h->stream));
  CU(h, cudaMemsetAsync(h->K, 0, B * T * h->m * h->n * s, h->stream)); /* src/ilqr_core.cpp:44-48 */
  CU(h, cudaMemsetAsync(h->k, 0, B * T * h->m * s, h->stream));
  const int rc = launch(h, kOpInit, 0, 0.0);
  if (rc == ILQR_OK) h->initialised = true;
  return rc;
}

int ilqr_warm_start(ilqr_handle *h, const void *x0, int on_device) {
  if (!h || !x0) return fail(h, ILQR_E_INVALID, "null argument");
  if (!h->initialised) return fail(h, ILQR_E_STATE, "ilqr_warm_start before ilqr_set_initial");
  DeviceGuard g(h->desc.device);
  CU(h, cudaMemcpyAsync(h->x0, x0, (size_t)h->desc.B * h->n * h->ssize,
                        on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  return launch(h, kOpWarm, 0, 0.0);
}

int ilqr_iterate(ilqr_handle *h, int n_iters) {
  if (!h) return ILQR_E_INVALID;
  if (!h->initialised) return fail(h, ILQR_E_STATE, "ilqr_iterate before ilqr_set_initial");
  if (n_iters < 0) return fail(h, ILQR_E_INVALID, "n_iters must be >= 0");
  DeviceGuard g(h->desc.device);
  if (!h->engine_warp && h->desc.model_id < ILQR_MODEL_USER_BASE) { /* the batch-lockstep phase kernels (ilqr_phases.cuh) */
    if (h->desc.flags & ILQR_FLAG_FAST_FMA)
      return h->desc.model_id == ILQR_MODEL_ACROBOT ? ilqr_phase_iterate_acrobot_fma(h, n_iters)
                                                    : ilqr_phase_iterate_double_integrator_fma(h, n_iters);
    return h->desc.model_id == ILQR_MODEL_ACROBOT ? ilqr_phase_iterate_acrobot(h, n_iters)
                                                  : ilqr_phase_iterate_double_integrator(h, n_iters);
  }
  return launch(h, kOpIterate, n_iters, 0.0);
}

int ilqr_resume(ilqr_handle *h) {
  if (!h) return ILQR_E_INVALID;
  if (!h->initialised) return fail(h, ILQR_E_STATE, "ilqr_resume before ilqr_set_initial");
  DeviceGuard g(h->desc.device);
  const int threads = 256;
  const unsigned blocks = (unsigned)((h->desc.B + threads - 1) / threads);
  if (h->desc.dtype == ILQR_F32) ilqr_resume_kernel<float><<<blocks, threads, 0, h->stream>>>((TrajState<float> *)h->st, h->desc.B);
  else ilqr_resume_kernel<double><<<blocks, threads, 0, h->stream>>>((TrajState<double> *)h->st, h->desc.B);
  CU(h, cudaGetLastError());
  h->launches++;
  return ILQR_OK;
}

int ilqr_solve(ilqr_handle *h) {
  if (!h) return ILQR_E_INVALID;
  return ilqr_iterate(h, h->desc.params.max_iter + 1);
}

int ilqr_backward_once(ilqr_handle *h, double lambda) {
  if (!h) return ILQR_E_INVALID;
  if (!h->initialised) return fail(h, ILQR_E_STATE, "ilqr_backward_once before ilqr_set_initial");
  DeviceGuard g(h->desc.device);
  return launch(h, kOpBackwardOnce, 0, lambda);
}

int ilqr_rollout_once(ilqr_handle *h, double alpha) {
  if (!h) return ILQR_E_INVALID;
  if (!h->initialised) return fail(h, ILQR_E_STATE, "ilqr_rollout_once before ilqr_set_initial");
  DeviceGuard g(h->desc.device);
  return launch(h, kOpRolloutOnce, 0, alpha);
}

int ilqr_get(ilqr_handle *h, int field, void *dst, int on_device) {
  if (!h || !dst) return fail(h, ILQR_E_INVALID, "null argument");
  if (!h->initialised) return fail(h, ILQR_E_STATE, "ilqr_get before ilqr_set_initial");
  DeviceGuard g(h->desc.device);
  const size_t B = (size_t)h->desc.B, T = (size_t)h->desc.T, s = h->ssize, n = h->n, m = h->m;
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  const void *src = nullptr;
  size_t bytes = 0;
  switch (field) {
    case ILQR_F_XS: src = h->xs; bytes = B * (T + 1) * n * s; break;
    case ILQR_F_US: src = h->us; bytes = B * T * m * s; break;
    case ILQR_F_K: src = h->K; bytes = B * T * m * n * s; break;
    case ILQR_F_KFF: src = h->k; bytes = B * T * m * s; break;
    case ILQR_F_VX0: src = h->Vx0; bytes = B * n * s; break;
    case ILQR_F_VXX0: src = h->Vxx0; bytes = B * n * n * s; break;
    case ILQR_F_COST: case ILQR_F_LAMBDA: case ILQR_F_DLAMBDA: case ILQR_F_GNORM: bytes = B * s; break;
    case ILQR_F_DV: bytes = B * 2 * s; break;
    case ILQR_F_ITERS: case ILQR_F_STATUS: case ILQR_F_ALPHA_INDEX: case ILQR_F_N_ACCEPT: case ILQR_F_N_REJECT:
    case ILQR_F_N_BACKWARD: case ILQR_F_DIVERGE: bytes = B * 4; break;
    default: return fail(h, ILQR_E_INVALID, "unknown field");
  }
  if (!src) { /* a column of the state records: gather on the device, then one dense copy */
    void *target = on_device ? dst : h->tmp;
    const int threads = 256;
    const unsigned blocks = (unsigned)((B + threads - 1) / threads);
    if (h->desc.dtype == ILQR_F32)
      ilqr_gather_kernel<float><<<blocks, threads, 0, h->stream>>>((const TrajState<float> *)h->st, (long long)B, field, target);
    else
      ilqr_gather_kernel<double><<<blocks, threads, 0, h->stream>>>((const TrajState<double> *)h->st, (long long)B, field, target);
    CU(h, cudaGetLastError());
    h->launches++;
    if (on_device) return ILQR_OK;
    src = h->tmp;
  }
  CU(h, cudaMemcpyAsync(dst, src, bytes, kind, h->stream));
  if (!on_device) CU(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

int ilqr_sync(ilqr_handle *h) {
  if (!h) return ILQR_E_INVALID;
  DeviceGuard g(h->desc.device);
  CU(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

void *ilqr_stream(ilqr_handle *h) { return h ? (void *)h->stream : nullptr; }

int64_t ilqr_launch_count(const ilqr_handle *h) { return h ? h->launches : 0; }

int ilqr_make_inputs(uint64_t seed, int64_t B, int32_t T, int32_t n, int32_t m, double x_scale, double u_scale,
                     int canonical_first, double *x0, double *u0) {
  if (B < 0 || T < 1 || n < 1 || m < 1 || !x0 || !u0) return ILQR_E_INVALID;
  ilqr_synth_fill(seed, (size_t)B, T, n, m, x_scale, u_scale, canonical_first, x0, u0);
  return ILQR_OK;
}

int ilqr_measure_fp64(int32_t device, double *fma_per_s, double *mul_per_s, double *add_per_s) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return ILQR_E_CUDA;
  DeviceGuard g(device);
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return ILQR_E_CUDA;
  double *sink = nullptr;
  if (cudaMalloc(&sink, 8) != cudaSuccess) return ILQR_E_NOMEM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 8, threads = 256, n = 1 << 15;
  double out[3] = {0, 0, 0};
  for (int op = 0; op < 3; op++) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      if (op == 0) ilqr_fp64_probe_kernel<0><<<blocks, threads>>>(1.5, 1.0000001, n, sink);
      if (op == 1) ilqr_fp64_probe_kernel<1><<<blocks, threads>>>(1.5, 1.0000001, n, sink);
      if (op == 2) ilqr_fp64_probe_kernel<2><<<blocks, threads>>>(1.5, 1.0000001, n, sink);
      cudaEventRecord(e1);
      if (cudaEventSynchronize(e1) != cudaSuccess) {
        cudaFree(sink);
        return ILQR_E_CUDA;
      }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    out[op] = (double)blocks * threads * 8.0 * n / (best * 1e-3);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  if (fma_per_s) *fma_per_s = out[0];
  if (mul_per_s) *mul_per_s = out[1];
  if (add_per_s) *add_per_s = out[2];
  return ILQR_OK;
}

const char *ilqr_version(void) { return "ilqr_b200 0.2 sm_100a batch-lockstep phase kernels + warp-per-trajectory kernel, f64/f32"; }

}  // extern "C"
