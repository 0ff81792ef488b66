// batch_solver.cpp — see batch_solver.h.  Host orchestration only: shards, threads, pinned staging, one NCCL gather.
#include "batch_solver.h"

#include <cuda_runtime.h>
#include <nccl.h>
#include <string.h>

#include <stdexcept>
#include <thread>

namespace {
void cu(cudaError_t e, const char *what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
void nc(ncclResult_t r, const char *what) {
  if (r != ncclSuccess) throw std::runtime_error(std::string(what) + ": " + ncclGetErrorString(r));
}
void ab(int rc, ilqr_handle *h, const char *what) {
  if (rc != ILQR_OK) throw std::runtime_error(std::string(what) + ": " + ilqr_last_error(h));
}
}  // namespace

struct BatchSolver::Shard {
  int device = 0;
  long lo = 0, hi = 0, padded = 0; /* [lo, hi) of the batch; every shard gathers `padded` entries */
  ilqr_handle *h = nullptr;
  double *stage_x0 = nullptr, *stage_u0 = nullptr; /* pinned, used when the caller's memory is pageable */
  double *cost_d = nullptr, *all_cost_d = nullptr; /* [padded], [world * padded] */
  int32_t *iters_d = nullptr, *all_iters_d = nullptr;
  cudaEvent_t g0 = nullptr, g1 = nullptr;
  std::string error;
};

void BatchSolver::shard_bounds(long total, int world, int rank, long *lo, long *hi) {
  const long base = total / world, extra = total % world;
  *lo = rank * base + (rank < extra ? rank : extra);
  *hi = *lo + base + (rank < extra ? 1 : 0);
}

double *BatchSolver::alloc_pinned(size_t doubles) {
  void *p = nullptr;
  cu(cudaHostAlloc(&p, doubles * sizeof(double), cudaHostAllocPortable), "cudaHostAlloc");
  return (double *)p;
}
void BatchSolver::free_pinned(double *p) {
  if (p) cudaFreeHost(p);
}

BatchSolver::BatchSolver(const ilqr_desc &desc, std::vector<int> devices) : desc_(desc), dev_(devices) {
  int32_t n = 0, m = 0;
  if (ilqr_model_info(desc.model_id, &n, &m, nullptr, nullptr) != ILQR_OK) throw std::runtime_error("BatchSolver: unknown model_id");
  n_ = n;
  m_ = m;
  if (desc.dtype != ILQR_F64) throw std::runtime_error("BatchSolver: f64 handles only");
  if (dev_.empty()) {
    int nd = 0;
    cu(cudaGetDeviceCount(&nd), "cudaGetDeviceCount");
    if (nd < 1) throw std::runtime_error("BatchSolver: no CUDA device (there is no CPU path)");
    for (int d = 0; d < nd; d++) dev_.push_back(d);
  }
  ncclComm_t *comms = new ncclComm_t[dev_.size()];
  nc(ncclCommInitAll(comms, (int)dev_.size(), dev_.data()), "ncclCommInitAll"); /* one process, all devices */
  comms_ = comms;
  for (size_t r = 0; r < dev_.size(); r++) {
    Shard *s = new Shard;
    s->device = dev_[r];
    shard_.push_back(s);
  }
}

BatchSolver::~BatchSolver() {
  for (size_t r = 0; r < shard_.size(); r++) {
    Shard *s = shard_[r];
    cudaSetDevice(s->device);
    ilqr_destroy(s->h);
    if (s->stage_x0) cudaFreeHost(s->stage_x0);
    if (s->stage_u0) cudaFreeHost(s->stage_u0);
    for (void *p : {(void *)s->cost_d, (void *)s->all_cost_d, (void *)s->iters_d, (void *)s->all_iters_d})
      if (p) cudaFree(p);
    if (s->g0) cudaEventDestroy(s->g0);
    if (s->g1) cudaEventDestroy(s->g1);
    delete s;
  }
  ncclComm_t *comms = (ncclComm_t *)comms_;
  if (comms) {
    for (size_t r = 0; r < dev_.size(); r++) ncclCommDestroy(comms[r]);
    delete[] comms;
  }
}

void BatchSolver::prepare(long B) {
  if (B == B_) return;
  const int world = (int)dev_.size();
  long padded = 0;
  for (int r = 0; r < world; r++) {
    shard_bounds(B, world, r, &shard_[r]->lo, &shard_[r]->hi);
    if (shard_[r]->hi - shard_[r]->lo > padded) padded = shard_[r]->hi - shard_[r]->lo;
  }
  for (int r = 0; r < world; r++) {
    Shard *s = shard_[r];
    cu(cudaSetDevice(s->device), "cudaSetDevice");
    ilqr_destroy(s->h);
    s->h = nullptr;
    for (void **p : {(void **)&s->cost_d, (void **)&s->all_cost_d, (void **)&s->iters_d, (void **)&s->all_iters_d}) {
      if (*p) cudaFree(*p);
      *p = nullptr;
    }
    if (s->stage_x0) cudaFreeHost(s->stage_x0);
    if (s->stage_u0) cudaFreeHost(s->stage_u0);
    s->stage_x0 = s->stage_u0 = nullptr;
    s->padded = padded;
    const long Bs = s->hi - s->lo;
    if (Bs > 0) {
      ilqr_desc d = desc_;
      d.B = Bs;
      d.device = s->device;
      ab(ilqr_create(&d, &s->h), nullptr, "ilqr_create");
    }
    cu(cudaMalloc((void **)&s->cost_d, padded * sizeof(double)), "cudaMalloc");
    cu(cudaMalloc((void **)&s->all_cost_d, (size_t)world * padded * sizeof(double)), "cudaMalloc");
    cu(cudaMalloc((void **)&s->iters_d, padded * sizeof(int32_t)), "cudaMalloc");
    cu(cudaMalloc((void **)&s->all_iters_d, (size_t)world * padded * sizeof(int32_t)), "cudaMalloc");
    cu(cudaMemset(s->cost_d, 0, padded * sizeof(double)), "cudaMemset");
    cu(cudaMemset(s->iters_d, 0, padded * sizeof(int32_t)), "cudaMemset");
    if (!s->g0) {
      cu(cudaEventCreate(&s->g0), "cudaEventCreate");
      cu(cudaEventCreate(&s->g1), "cudaEventCreate");
    }
  }
  B_ = B;
}

void BatchSolver::solve(const double *x0, const double *u0, long B, double *cost, int32_t *iters) {
  if (B < 1 || !x0 || !u0) throw std::runtime_error("BatchSolver::solve: bad arguments");
  prepare(B);
  const int world = (int)dev_.size();
  const size_t T = (size_t)desc_.T;
  cudaPointerAttributes at;
  const bool pinned = cudaPointerGetAttributes(&at, x0) == cudaSuccess && at.type == cudaMemoryTypeHost &&
                      cudaPointerGetAttributes(&at, u0) == cudaSuccess && at.type == cudaMemoryTypeHost;
  cudaGetLastError();
  /* one host thread per device: set_initial (H2D on the handle's stream), the solve, the per-shard results */
  std::vector<std::thread> workers;
  for (int r = 0; r < world; r++) {
    workers.emplace_back([&, r]() {
      Shard *s = shard_[r];
      try {
        const long Bs = s->hi - s->lo;
        if (Bs <= 0) return;
        cu(cudaSetDevice(s->device), "cudaSetDevice");
        const double *px = x0 + (size_t)s->lo * n_, *pu = u0 + (size_t)s->lo * T * m_;
        if (!pinned) { /* stage through pinned memory of this shard's own */
          if (!s->stage_x0) {
            s->stage_x0 = alloc_pinned((size_t)Bs * n_);
            s->stage_u0 = alloc_pinned((size_t)Bs * T * m_);
          }
          memcpy(s->stage_x0, px, (size_t)Bs * n_ * sizeof(double));
          memcpy(s->stage_u0, pu, (size_t)Bs * T * m_ * sizeof(double));
          px = s->stage_x0;
          pu = s->stage_u0;
        }
        ab(ilqr_set_initial(s->h, px, pu, 0), s->h, "ilqr_set_initial");
        ab(ilqr_solve(s->h), s->h, "ilqr_solve");
        ab(ilqr_get(s->h, ILQR_F_COST, s->cost_d, 1), s->h, "ilqr_get");
        ab(ilqr_get(s->h, ILQR_F_ITERS, s->iters_d, 1), s->h, "ilqr_get");
      } catch (const std::exception &e) {
        s->error = e.what();
      }
    });
  }
  for (std::thread &t : workers) t.join();
  for (int r = 0; r < world; r++)
    if (!shard_[r]->error.empty()) throw std::runtime_error("BatchSolver shard " + std::to_string(r) + ": " + shard_[r]->error);
  /* the one collective of the job: every device receives every shard's final costs (and trip counts) */
  ncclComm_t *comms = (ncclComm_t *)comms_;
  nc(ncclGroupStart(), "ncclGroupStart");
  for (int r = 0; r < world; r++) {
    Shard *s = shard_[r];
    cu(cudaSetDevice(s->device), "cudaSetDevice");
    cudaStream_t st = s->h ? (cudaStream_t)ilqr_stream(s->h) : 0;
    cu(cudaEventRecord(s->g0, st), "cudaEventRecord");
    nc(ncclAllGather(s->cost_d, s->all_cost_d, (size_t)s->padded, ncclDouble, comms[r], st), "ncclAllGather");
    nc(ncclAllGather(s->iters_d, s->all_iters_d, (size_t)s->padded, ncclInt32, comms[r], st), "ncclAllGather");
    cu(cudaEventRecord(s->g1, st), "cudaEventRecord");
  }
  nc(ncclGroupEnd(), "ncclGroupEnd");
  /* results leave through device 0 */
  Shard *s0 = shard_[0];
  cu(cudaSetDevice(s0->device), "cudaSetDevice");
  std::vector<double> all((size_t)world * s0->padded);
  std::vector<int32_t> alli((size_t)world * s0->padded);
  cudaStream_t st0 = s0->h ? (cudaStream_t)ilqr_stream(s0->h) : 0;
  cu(cudaMemcpyAsync(all.data(), s0->all_cost_d, all.size() * sizeof(double), cudaMemcpyDeviceToHost, st0), "cudaMemcpyAsync");
  cu(cudaMemcpyAsync(alli.data(), s0->all_iters_d, alli.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st0), "cudaMemcpyAsync");
  cu(cudaStreamSynchronize(st0), "cudaStreamSynchronize");
  gather_ms = 0;
  total_trips = 0;
  for (int r = 0; r < world; r++) {
    Shard *s = shard_[r];
    cu(cudaSetDevice(s->device), "cudaSetDevice");
    cu(cudaEventSynchronize(s->g1), "cudaEventSynchronize");
    float ms = 0;
    cudaEventElapsedTime(&ms, s->g0, s->g1);
    if (ms > gather_ms) gather_ms = ms;
    for (long b = s->lo; b < s->hi; b++) {
      if (cost) cost[b] = all[(size_t)r * s->padded + (b - s->lo)];
      const int32_t it = alli[(size_t)r * s->padded + (b - s->lo)];
      if (iters) iters[b] = it;
      total_trips += it;
    }
  }
}

void BatchSolver::get(int field, void *dst) {
  size_t per = 0; /* bytes per instance */
  const size_t T = (size_t)desc_.T, n = n_, m = m_;
  switch (field) {
    case ILQR_F_XS: per = (T + 1) * n * 8; break;
    case ILQR_F_US: per = T * m * 8; break;
    case ILQR_F_K: per = T * m * n * 8; break;
    case ILQR_F_KFF: per = T * m * 8; break;
    case ILQR_F_VX0: per = n * 8; break;
    case ILQR_F_VXX0: per = n * n * 8; break;
    case ILQR_F_DV: per = 16; break;
    case ILQR_F_COST: case ILQR_F_LAMBDA: case ILQR_F_DLAMBDA: case ILQR_F_GNORM: per = 8; break;
    default: per = 4; break;
  }
  for (Shard *s : shard_) {
    if (!s->h) continue;
    cu(cudaSetDevice(s->device), "cudaSetDevice");
    ab(ilqr_get(s->h, field, (char *)dst + (size_t)s->lo * per, 0), s->h, "ilqr_get");
  }
}
