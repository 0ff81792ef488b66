/*
 * ilqr.h — drop-in replacement for the reference's include/ilqr.h: the same `class iLQR`
 * (constructor taking ownership of a `Model*`, the three `generate_trajectory` overloads,
 * `init_traj`, `output_to_csv`; reference include/ilqr.h:28-54), implemented on libilqr_b200.so
 * through the C ABI (include/ilqr_b200.h) instead of the Eigen loops of src/ilqr_core.cpp.
 *
 * Put this directory BEFORE the reference's include/ on the include path: the reference's own
 * model.h, common.h, acrobot.h, double_integrator.h and its src/run_ilqr.cpp then compile
 * unchanged against it (ilqr_b200/host/Makefile does exactly that).  On top of the reference's
 * surface it adds what the reference keeps private (read access to xs, us, K, k, cost) and a
 * batched entry point, since one trajectory cannot fill a B200.
 *
 * A Model subclass runs on the GPU through a hand-written device twin (ilqr_b200/csrc/models.cuh).
 * The constructor recognises the reference's Acrobot and DoubleIntegrator by RTTI, checks the twin
 * against the host object's own dynamics / cost / final_cost on random probes, and throws
 * std::runtime_error for any other subclass or on a mismatch: there is no CPU fallback.
 */
#ifndef _ILQR_H_
#define _ILQR_H_

#include <stdio.h>

#include <memory>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#include "common.h"
#include "model.h"

#include "ilqr_b200.h"

/* one trajectory (xs [T+1][n], us [T][m], row-major) in the reference's result format, src/ilqr_core.cpp:414-431 */
void ilqr_write_csv(FILE *f, int T, int n, int m, const double *xs, const double *us);

class iLQR {
 public:
  iLQR(Model *p_dyn, double timeDelta);
  ~iLQR();
  iLQR(const iLQR &) = delete;
  iLQR &operator=(const iLQR &) = delete;

  /* the reference's public methods (include/ilqr.h:49-54) */
  void generate_trajectory();                                          /* continue the current solve      */
  void generate_trajectory(const VectorXd &x_0);                       /* warm start, src/ilqr_core.cpp:65-76 */
  void generate_trajectory(const VectorXd &x_0, const VecOfVecXd &u0); /* fresh solve, :59-62             */
  void output_to_csv(const std::string filename);                      /* :414-431, byte for byte          */
  double init_traj(const VectorXd &x_0, const VecOfVecXd &u_0);        /* :11-56, returns the initial cost */

  /* what the reference keeps private (include/ilqr.h:57-85) */
  const VecOfVecXd &get_xs() const { return xs; }
  const VecOfVecXd &get_us() const { return us; }
  const VecOfVecXd &get_k() const { return k; }
  const VecOfMatXd &get_K() const { return K; }
  double get_cost() const { return cost_s; }
  int get_iterations() const { return iterations; }
  int get_status() const { return status; } /* ILQR_EXIT_* of include/ilqr_b200.h */

  /* A Model subclass of the user's own.  The host object (dynamics / cost / final_cost virtuals, include/model.h:6-21)
   * cannot run in a kernel, so its device twin is registered once as CUDA source — a struct with the static interface
   * of ilqr_b200/csrc/models.cuh, see include/ilqr_b200.h: ilqr_register_model — and `new iLQR(new MyModel, dt)` then
   * finds it by the dynamic type of the object.  model_params reach the twin's functions as `mp`.  The solver checks
   * the twin against the host virtuals at construction: one Euler step and both costs on a few probe points through
   * the GPU must agree with p_dyn->integrate_dynamics / cost / final_cost to 1e-9. */
  template <class UserModel>
  static void register_device_twin(const char *struct_name, const char *cuda_source,
                                   const std::vector<double> &model_params = std::vector<double>()) {
    register_device_twin(typeid(UserModel), struct_name, cuda_source, model_params);
  }
  static void register_device_twin(const std::type_info &type, const char *struct_name, const char *cuda_source,
                                   const std::vector<double> &model_params);

  /* B independent problems at once: X0 is B x n, U0[b] the T initial controls of problem b.
   * Returns the final costs; the trajectory of problem b is then available through batch_xs(b) etc. */
  std::vector<double> solve_batch(const std::vector<VectorXd> &X0, const std::vector<VecOfVecXd> &U0);
  VecOfVecXd batch_xs(int b) const;
  VecOfVecXd batch_us(int b) const;
  int batch_iterations(int b) const { return batch_iters.at(b); }
  int batch_exit_status(int b) const { return batch_status.at(b); }
  /* results of the last solve_batch on disk: one trajectory in the reference's CSV format (what plot_results.py
   * reads), or the whole batch in one binary file (layout in ilqr_host.cpp; reader: ilqr_b200/export.py) */
  void output_to_csv(const std::string filename, int b) const;
  void export_batch(const std::string filename) const;

  std::shared_ptr<Model> model;
  double dt;
  int T = 0;
  /* solver settings the reference hard-codes at file scope (include/ilqr.h:14-22); defaults are the reference's */
  int maxIter = 100;
  bool quiet = false; /* the reference prints a progress table; here only the banner lines survive */
  int cost_deriv = 0; /* ILQR_COST_FD (the reference's behaviour) | ILQR_COST_ANALYTIC */
  int flags = 0;      /* ILQR_FLAG_* of include/ilqr_b200.h; 0 = the reference's behaviour */
  /* solve_batch: CUDA ordinals to shard the batch over (contiguous blocks, one handle and one host thread per device,
   * one ncclAllGather of the final costs: batch_solver.h).  Empty = device 0 alone through the plain handle. */
  std::vector<int> devices;

 private:
  void create(long B, int T_);
  void fetch_single();
  ilqr_handle *h = nullptr;
  ilqr_desc hdesc; /* what h was created with */
  long hB = 0;
  int hT = 0;
  int model_id = -1;
  double model_params[16] = {0};
  VecOfVecXd xs, us, k;
  VecOfMatXd K;
  double cost_s = 0;
  int iterations = 0, status = 0;
  std::vector<double> bxs, bus;
  std::vector<int> batch_iters, batch_status;
  std::vector<double> batch_cost;
  class BatchSolver *multi = nullptr; /* the sharded path of solve_batch */
  std::vector<int> multi_devices;
  long multi_B = 0;
  int multi_T = 0;
};

#endif
