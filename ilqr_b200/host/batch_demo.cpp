// batch_demo.cpp — the reference's models, unchanged, driving a batch through the host shim:
//   batch_demo [B] [T]   -> solves B random acrobot instances (include/ilqr_synth.h, seed 12345) and prints
//   per-instance cost / iterations for the first few, plus the single-trajectory API on instance 0.
//   batch_demo B T PREFIX  additionally writes PREFIX.bin (whole batch, iLQR::export_batch) and PREFIX_b3.csv
//   (trajectory 3 in the reference's CSV format).
#include <stdio.h>
#include <stdlib.h>

#include "acrobot.h"
#include "ilqr.h"

#include "../../include/ilqr_synth.h"

int main(int argc, char **argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 64, T = argc > 2 ? atoi(argv[2]) : 200;
  std::vector<double> x0((size_t)B * 4), u0((size_t)B * T);
  ilqr_synth_fill(12345, B, T, 4, 1, 1.0, 0.5, 1, x0.data(), u0.data());
  std::vector<VectorXd> X0(B, VectorXd::Zero(4));
  std::vector<VecOfVecXd> U0(B, VecOfVecXd(T, VectorXd::Zero(1)));
  for (int b = 0; b < B; b++) {
    for (int i = 0; i < 4; i++) X0[b](i) = x0[(size_t)b * 4 + i];
    for (int t = 0; t < T; t++) U0[b][t](0) = u0[(size_t)b * T + t];
  }
  iLQR solver(new Acrobot(), 0.02);
  solver.quiet = true;
  if (const char *e = getenv("ILQR_DEMO_DEVICES")) /* e.g. "0,1": shard the batch over these GPUs (iLQR::devices) */
    for (const char *p = e; *p;) {
      solver.devices.push_back(atoi(p));
      while (*p && *p != ',') p++;
      if (*p == ',') p++;
    }
  std::vector<double> cost = solver.solve_batch(X0, U0);
  for (int b = 0; b < B && b < 6; b++) printf("batch %d cost %.12f iterations %d\n", b, cost[b], solver.batch_iterations(b));
  if (argc > 3) {
    solver.export_batch(std::string(argv[3]) + ".bin");
    solver.output_to_csv(std::string(argv[3]) + "_b3.csv", B > 3 ? 3 : 0);
  }
  solver.generate_trajectory(X0[0], U0[0]);
  printf("single 0 cost %.12f iterations %d status %d xT %.9f %.9f %.9f %.9f\n", solver.get_cost(), solver.get_iterations(),
         solver.get_status(), solver.get_xs()[T](0), solver.get_xs()[T](1), solver.get_xs()[T](2), solver.get_xs()[T](3));
  return 0;
}
