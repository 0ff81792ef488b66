// bench_batch.cpp — end-to-end rate of the C++ host path: host arrays in, final costs out, every visible GPU, ONE process.
//   bench_batch [B_total = 4096 per GPU] [T = 200] [steps = 5] [warmup = 3] [cost_deriv: analytic|fd]
// Synthetic acrobot instances (include/ilqr_synth.h, seed 12345: BASELINE configs[1] at B = 4096 on one GPU, configs[4] at
// B = 1048576 on eight).  Prints one JSON line; the timed region is BatchSolver::solve — pinned host inputs -> H2D ->
// init_traj + generate_trajectory on every shard -> one ncclAllGather of costs and trip counts -> host.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "../../include/ilqr_synth.h"
#include "batch_solver.h"

int main(int argc, char **argv) {
  if (argc > 1 && strcmp(argv[1], "--shards") == 0) { /* bench_batch --shards TOTAL WORLD: the partition, no GPU needed */
    const long total = argc > 2 ? atol(argv[2]) : 0;
    const int world = argc > 3 ? atoi(argv[3]) : 1;
    for (int r = 0; r < world; r++) {
      long lo, hi;
      BatchSolver::shard_bounds(total, world, r, &lo, &hi);
      printf("%ld %ld\n", lo, hi);
    }
    return 0;
  }
  try {
    ilqr_desc d;
    memset(&d, 0, sizeof(d));
    d.model_id = ILQR_MODEL_ACROBOT;
    d.dtype = ILQR_F64;
    d.T = argc > 2 ? atoi(argv[2]) : 200;
    d.dt = 0.02;
    d.cost_deriv = (argc > 5 && strcmp(argv[5], "fd") == 0) ? ILQR_COST_FD : ILQR_COST_ANALYTIC;
    ilqr_default_params(&d.params);
    BatchSolver solver(d);
    const int G = solver.num_devices();
    const long B = argc > 1 && atol(argv[1]) > 0 ? atol(argv[1]) : 4096L * G;
    const int steps = argc > 3 ? atoi(argv[3]) : 5, warmup = argc > 4 ? atoi(argv[4]) : 3;
    double *x0 = BatchSolver::alloc_pinned((size_t)B * 4), *u0 = BatchSolver::alloc_pinned((size_t)B * d.T);
    ilqr_synth_fill(12345, (size_t)B, d.T, 4, 1, 1.0, 0.5, 1, x0, u0);
    std::vector<double> cost(B);
    std::vector<int32_t> iters(B);
    for (int i = 0; i < warmup; i++) solver.solve(x0, u0, B, cost.data(), iters.data());
    double secs = 0, gather = 0;
    long trips = 0;
    for (int i = 0; i < steps; i++) {
      const auto t0 = std::chrono::steady_clock::now();
      solver.solve(x0, u0, B, cost.data(), iters.data());
      secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      trips += solver.total_trips;
      gather += solver.gather_ms;
    }
    double csum = 0;
    for (long b = 0; b < B; b++) csum += cost[b];
    printf("{\"host\": \"c++ BatchSolver (one process, ncclCommInitAll)\", \"n_gpus\": %d, \"batch_total\": %ld, \"T\": %d, "
           "\"steps\": %d, \"e2e_iterations_per_s\": %.1f, \"ms_per_step\": %.3f, \"trips_per_step\": %.1f, "
           "\"gather_ms\": %.3f, \"cost0\": %.12g, \"cost_checksum\": %.12g}\n",
           G, B, d.T, steps, trips / secs, 1e3 * secs / steps, (double)trips / steps, gather / steps, cost[0], csum);
    BatchSolver::free_pinned(x0);
    BatchSolver::free_pinned(u0);
    return 0;
  } catch (const std::exception &e) {
    fprintf(stderr, "bench_batch: %s\n", e.what());
    return 1;
  }
}
