// user_model_demo.cpp — the reference's plugin surface with a model of the user's own:
//   user_model_demo <twin.cu> [T]
// `class Pendulum : public Model` (pendulum_model.h) is what a user of the reference writes (include/model.h:6-21: dynamics, cost,
// final_cost, u_min / u_max, x_dims / u_dims).  Its device twin — the same three functions as a CUDA struct, read
// here from the file given on the command line — is registered once; `new iLQR(new Pendulum, dt)` then runs on the
// GPU, after checking the twin against the host object.
#include <stdio.h>
#include <stdlib.h>

#include <fstream>
#include <sstream>

#include "common.h"
#include "model.h"
#include "ilqr.h"

#include "pendulum_model.h"

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: user_model_demo <twin.cu> [T]\n");
    return 2;
  }
  std::ifstream f(argv[1]);
  std::stringstream ss;
  ss << f.rdbuf();
  const int T = argc > 2 ? atoi(argv[2]) : 150;
  try {
    iLQR::register_device_twin<Pendulum>("Pendulum", ss.str().c_str(), {3.141592653589793});
    iLQR solver(new Pendulum, 0.05);
    solver.quiet = true;
    solver.cost_deriv = 1; /* closed-form cost derivatives of the twin */
    VectorXd x0(2);
    x0 << 0.2, -0.1;
    VecOfVecXd u0(T, VectorXd::Zero(1));
    for (int t = 0; t < T; t++) u0[t](0) = 0.1 * sin(0.3 * t);
    const double c0 = solver.init_traj(x0, u0);
    solver.generate_trajectory(x0, u0);
    printf("pendulum initial cost %.12f final cost %.12f iterations %d status %d xT %.9f %.9f\n", c0, solver.get_cost(),
           solver.get_iterations(), solver.get_status(), solver.get_xs()[T](0), solver.get_xs()[T](1));
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
