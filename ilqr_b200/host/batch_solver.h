/*
 * batch_solver.h — a batch of independent iLQR problems on EVERY GPU of the box from ONE host process.
 *
 * Problem instances share nothing (no cross-instance term anywhere in src/ilqr_core.cpp), so the batch is cut into
 * contiguous shards, one per device, each solved by its own handle of the C ABI (include/ilqr_b200.h) driven by its own
 * host thread; nothing is exchanged during the solve.  The ONLY collective is one ncclAllGather of the final costs
 * (plus a 4-byte-per-instance one of the trip counts) over NVLink, on communicators made with ncclCommInitAll — no
 * launcher, no second process.  Inputs are read from pinned host memory (alloc_pinned), so the H2D copies of the
 * shards run concurrently at full PCIe rate and no element-by-element marshalling happens on the way in.
 *
 * Plain arrays in the ABI's layouts: x0 [B][n], u0 [B][T][m], row-major, f64.
 */
#ifndef ILQR_BATCH_SOLVER_H_
#define ILQR_BATCH_SOLVER_H_

#include <stdint.h>

#include <string>
#include <vector>

#include "ilqr_b200.h"

class BatchSolver {
 public:
  /* desc: model, T, dt, cost_deriv, limits, params, flags (B and device are filled in per shard); devices: CUDA
   * ordinals, empty = every visible device */
  explicit BatchSolver(const ilqr_desc &desc, std::vector<int> devices = std::vector<int>());
  ~BatchSolver();
  BatchSolver(const BatchSolver &) = delete;
  BatchSolver &operator=(const BatchSolver &) = delete;

  /* init_traj + generate_trajectory for B instances; cost [B] and iters [B] (may be null) are written on return.
   * x0 / u0 may be pageable; pinned memory (alloc_pinned) avoids the staging copy. */
  void solve(const double *x0, const double *u0, long B, double *cost, int32_t *iters);
  /* any per-instance field of ilqr_get after a solve, gathered shard by shard into dst (host, the ABI's layout) */
  void get(int field, void *dst);

  static double *alloc_pinned(size_t doubles);
  static void free_pinned(double *p);
  static void shard_bounds(long total, int world, int rank, long *lo, long *hi); /* blocks differ by at most one */

  int num_devices() const { return (int)dev_.size(); }
  int n() const { return n_; }
  int m() const { return m_; }
  long total_trips = 0;   /* of the last solve */
  double gather_ms = 0;   /* device time of the collective of the last solve (max over devices) */

 private:
  struct Shard;
  void prepare(long B);
  ilqr_desc desc_;
  std::vector<int> dev_;
  std::vector<Shard *> shard_;
  void *comms_ = nullptr; /* ncclComm_t[num_devices] */
  long B_ = 0;
  int n_ = 0, m_ = 0;
};

#endif
