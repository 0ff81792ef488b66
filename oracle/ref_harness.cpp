// oracle/ref_harness.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" probe around the UNMODIFIED reference (kazuotani14/iLQR),
// compiled in place from /root/reference by oracle/Makefile into
// oracle/_ref/libref_oracle.so and oracle/_ref/ref_bench.  No reference source
// is copied into this repository: the three library translation units are
// pulled in by path below, as one unity TU, because
//   * include/finite_diff.h defines non-inline functions (one TU only), and
//   * `lambda` / `dlambda` are mutable TU-level statics (include/ilqr.h:17-18)
//     that the probe must save, reset and restore per problem instance.
// Private solver state is reached through the FRIEND_TEST hook the reference
// declares (include/ilqr.h:103-106) together with oracle/stub/gtest/gtest_prod.h.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` leg may load what this file builds.
#include "src/ilqr_core.cpp"
#include "src/derivatives.cpp"
#include "src/boxqp.cpp"
#include "acrobot.h"
#include "double_integrator.h"
// a Model subclass of the user's own (the plugin surface, include/model.h:6-21): the reference solves it like any other
#include "../ilqr_b200/host/pendulum_model.h"

#include <fcntl.h>
#include <random>
#include <unistd.h>

namespace {

// The reference prints progress unconditionally (SHOWPROGRESS/TIMESTUFF are
// hard-defined, src/ilqr_core.cpp:1-2); keep the probe quiet.
struct Silence {
  int saved;
  Silence() {
    fflush(stdout);
    std::cout.flush();
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    close(nul);
  }
  ~Silence() {
    fflush(stdout);
    std::cout.flush();
    dup2(saved, 1);
    close(saved);
  }
};

}  // namespace

// Exit reasons reported by the probe (the reference only prints them,
// src/ilqr_core.cpp:156,259,278).
enum { REF_RUNNING = 0, REF_GRAD = 1, REF_TOLFUN = 2, REF_LAMBDA_MAX = 3, REF_MAXITER = 4 };

struct RefState {
  iLQR *solver = nullptr;
  Model *model = nullptr;  // owned by solver->model
  int n = 0, m = 0;
  // how the instance was made, so ref_init() can start from a FRESH iLQR object every time:
  // init_traj() on a reused object rolls out with the previous solve's K (K.size()>0 at
  // src/ilqr_core.cpp:316 before :48 zeroes it), which is not the "fresh process" semantics.
  int model_id = 0;
  double dt = 0;
  double goal[4] = {0, 0, 0, 0};
  bool have_limits = false;
  double u_min[4], u_max[4];
  // per-instance copies of the reference's TU statics + loop-carried locals
  double lam = 1, dlam = 1;
  bool flgChange = true;
  int iter = 0;          // loop counter of src/ilqr_core.cpp:103
  int loop_trips = 0;    // loop bodies entered (the "iteration" unit of the metric)
  int status = REF_RUNNING;
  // last-iteration diagnostics
  double gnorm = 0, dcost = 0, expected = 0, alpha = 0, new_cost = 0;
  int alpha_index = -1;
  // counters
  long n_accept = 0, n_reject = 0, n_rollouts = 0, n_backward = 0, n_deriv = 0;
};

// The friend named by FRIEND_TEST(ILQRSetup, ForwardPassTest) (include/ilqr.h:106).
class ILQRSetup_ForwardPassTest_Test {
 public:
  static double init(RefState *s, const VectorXd &x0, const VecOfVecXd &u0) {
    Silence q;
    s->lam = 1;
    s->dlam = 1;
    s->flgChange = true;
    s->iter = 0;
    s->loop_trips = 0;
    s->status = REF_RUNNING;
    s->alpha_index = -1;
    s->n_accept = s->n_reject = s->n_rollouts = s->n_backward = s->n_deriv = 0;
    return s->solver->init_traj(x0, u0);
  }

  // Replica of the loop body of iLQR::generate_trajectory()
  // (src/ilqr_core.cpp:103-288): same statements in the same order, calling the
  // reference's own private methods; the prints, timers and CSV dump are left
  // out and the loop-carried locals live in RefState so that iterate(1) x N is
  // the same as iterate(N).  Validated against the real generate_trajectory()
  // by tests/test_oracle_ref.py.
  static int iterate(RefState *s, int n_iters) {
    iLQR &q = *s->solver;
    lambda = s->lam;
    dlambda = s->dlam;
    VecOfVecXd x_old, u_old;
    int done_here = 0;
    for (; s->iter < maxIter && done_here < n_iters && s->status == REF_RUNNING; s->iter++) {
      done_here++;
      s->loop_trips++;
      x_old = q.xs;
      u_old = q.us;

      if (s->flgChange) {
        q.get_dynamics_derivatives(q.xs, q.us, q.fx, q.fu);
        q.get_cost_derivatives(q.xs, q.us, q.cx, q.cu);
        q.get_cost_2nd_derivatives(q.xs, q.us, q.cxx, q.cxu, q.cuu);
        s->flgChange = false;
        s->n_deriv++;
      }

      bool backPassDone = false;
      while (!backPassDone) {
        int diverge = q.backward_pass();
        s->n_backward++;
        if (diverge != 0) {
          dlambda = std::max(dlambda * lambdaFactor, lambdaFactor);
          lambda = std::max(lambda * dlambda, lambdaMin);
          if (lambda > lambdaMax) break;
          continue;
        }
        backPassDone = true;
      }

      s->gnorm = q.get_gradient_norm(q.k, q.us);
      if (s->gnorm < tolGrad && lambda < 1e-5) {
        s->status = REF_GRAD;
        break;
      }

      bool fwdPassDone = false;
      double alpha = 0;
      s->alpha_index = -1;
      if (backPassDone) {
        for (int i = 0; i < Alpha.size(); i++) {
          alpha = Alpha(i);
          VecOfVecXd u_plus_feedforward = q.us;
          for (unsigned int j = 0; j < q.us.size(); j++) u_plus_feedforward[j] += q.k[j] * alpha;

          s->new_cost = q.forward_pass(q.x0, u_plus_feedforward);
          s->n_rollouts++;
          s->dcost = q.cost_s - s->new_cost;
          s->expected = -alpha * (q.dV(0) + alpha * q.dV(1));
          double z;
          if (s->expected > 0) {
            z = s->dcost / s->expected;
          } else {
            z = sgn(s->dcost);
          }
          if (z > zMin) {
            fwdPassDone = true;
            s->alpha_index = i;
            break;
          }
          q.xs = x_old;
          q.us = u_old;
        }
        if (!fwdPassDone) alpha = 0.0;
      }
      s->alpha = alpha;

      if (fwdPassDone) {
        dlambda = std::min(dlambda / lambdaFactor, 1 / lambdaFactor);
        lambda = lambda * dlambda * (lambda > lambdaMin);
        q.cost_s = s->new_cost;
        s->flgChange = true;
        s->n_accept++;
        if (s->dcost < tolFun) {
          s->status = REF_TOLFUN;
          break;
        }
      } else {
        dlambda = std::max(dlambda * lambdaFactor, lambdaFactor);
        lambda = std::max(lambda * dlambda, lambdaMin);
        s->n_reject++;
        if (lambda > lambdaMax) {
          s->status = REF_LAMBDA_MAX;
          break;
        }
      }
    }
    if (s->status == REF_RUNNING && s->iter >= maxIter) s->status = REF_MAXITER;
    s->lam = lambda;
    s->dlam = dlambda;
    return done_here;
  }

  // One derivative sweep + one backward pass at a given lambda, no loop.
  static int backward_once(RefState *s, double lam, int recompute_derivs) {
    iLQR &q = *s->solver;
    if (recompute_derivs) {
      q.get_dynamics_derivatives(q.xs, q.us, q.fx, q.fu);
      q.get_cost_derivatives(q.xs, q.us, q.cx, q.cu);
      q.get_cost_2nd_derivatives(q.xs, q.us, q.cxx, q.cxu, q.cuu);
    }
    lambda = lam;
    int d = q.backward_pass();
    s->gnorm = q.get_gradient_norm(q.k, q.us);
    return d;
  }

  // The closed-loop rollout of the line search for one alpha
  // (src/ilqr_core.cpp:188-197): overwrites xs/us like the reference does.
  static double rollout_once(RefState *s, double alpha) {
    iLQR &q = *s->solver;
    VecOfVecXd u_plus = q.us;
    for (unsigned int j = 0; j < q.us.size(); j++) u_plus[j] += q.k[j] * alpha;
    return q.forward_pass(q.x0, u_plus);
  }

  // The real thing, start to finish (prints silenced; CSV lands in cwd).
  static void solve_native(RefState *s, const VectorXd &x0, const VecOfVecXd &u0) {
    Silence quiet;
    lambda = 1;
    dlambda = 1;
    s->solver->generate_trajectory(x0, u0);
    s->lam = lambda;
    s->dlam = dlambda;
  }

  // Warm start.  Replica of iLQR::generate_trajectory(const VectorXd&) up to the loop (src/ilqr_core.cpp:65-76):
  // the same two statements on the reference's own members, then the loop-carried locals the reference
  // re-initialises on entry to generate_trajectory() (:88-95, :102: flgChange = true, iter = 0).  lambda / dlambda are
  // NOT touched: they are TU statics (include/ilqr.h:17-18) and carry over from the previous solve.
  static double warm_start(RefState *s, const VectorXd &x_0) {
    Silence quiet;
    iLQR &q = *s->solver;
    q.x0 = x_0;
    const double cost_i = q.forward_pass(x_0, q.us);
    q.cost_s = cost_i;
    s->flgChange = true;
    s->iter = 0;
    s->status = REF_RUNNING;
    return cost_i;
  }
  // The real warm start, start to finish: the TU statics are set to what this instance's previous solve left.
  static void warm_native(RefState *s, const VectorXd &x_0) {
    Silence quiet;
    lambda = s->lam;
    dlambda = s->dlam;
    s->solver->generate_trajectory(x_0);
    s->lam = lambda;
    s->dlam = dlambda;
  }
  // "Continue": iLQR::generate_trajectory() called again on a finished solve re-enters the loop with iter = 0 and
  // flgChange = true (:88-102), lambda / dlambda carried over.
  static void resume(RefState *s) {
    s->flgChange = true;
    s->iter = 0;
    s->status = REF_RUNNING;
  }
  static void resume_native(RefState *s) {
    Silence quiet;
    lambda = s->lam;
    dlambda = s->dlam;
    s->solver->generate_trajectory();
    s->lam = lambda;
    s->dlam = dlambda;
  }

  static int T(RefState *s) { return s->solver->T; }

  // field ids shared with tests/refharness.py
  static int get(RefState *s, int field, double *dst) {
    iLQR &q = *s->solver;
    const int T = q.T, n = s->n, m = s->m;
    auto vecs = [&](const VecOfVecXd &v, int count, int len) {
      for (int t = 0; t < count; t++)
        for (int i = 0; i < len; i++) dst[t * len + i] = v[t](i);
      return count * len;
    };
    auto mats = [&](const VecOfMatXd &v, int count, int r, int c) {  // row-major out
      for (int t = 0; t < count; t++)
        for (int i = 0; i < r; i++)
          for (int j = 0; j < c; j++) dst[(t * r + i) * c + j] = v[t](i, j);
      return count * r * c;
    };
    switch (field) {
      case 0: return vecs(q.xs, T + 1, n);
      case 1: return vecs(q.us, T, m);
      case 2: return mats(q.K, T, m, n);
      case 3: return vecs(q.k, T, m);
      case 4: dst[0] = q.cost_s; return 1;
      case 5: dst[0] = q.dV(0); dst[1] = q.dV(1); return 2;
      case 6: return vecs(q.Vx, T + 1, n);
      case 7: return mats(q.Vxx, T + 1, n, n);
      case 8: return mats(q.fx, T + 1, n, n);
      case 9: return mats(q.fu, T + 1, n, m);
      case 10: return vecs(q.cx, T + 1, n);
      case 11: return vecs(q.cu, T + 1, m);
      case 12: return mats(q.cxx, T + 1, n, n);
      case 13: return mats(q.cxu, T + 1, n, m);
      case 14: return mats(q.cuu, T + 1, m, m);
      default: return -1;
    }
  }
};
typedef ILQRSetup_ForwardPassTest_Test Probe;

extern "C" {

// model_id: 0 = Acrobot (include/acrobot.h), 2 = the user-model example Pendulum(goal[0]), 1 = DoubleIntegrator(goal)
// (include/double_integrator.h).  u_min/u_max may be NULL (keep the model's own).
static void rebuild(RefState *s) {
  delete s->solver;  // also deletes the model it owns
  if (s->model_id == 0) {
    s->model = new Acrobot();
  } else if (s->model_id == 2) {
    s->model = new Pendulum(s->goal[0]);
  } else {
    VectorXd g(4);
    for (int i = 0; i < 4; i++) g(i) = s->goal[i];
    s->model = new DoubleIntegrator(g);
  }
  s->n = s->model->x_dims;
  s->m = s->model->u_dims;
  if (s->have_limits)
    for (int j = 0; j < s->m; j++) {
      s->model->u_min(j) = s->u_min[j];
      s->model->u_max(j) = s->u_max[j];
    }
  s->solver = new iLQR(s->model, s->dt);  // takes ownership (include/ilqr.h:30-31)
}

void *ref_new(int model_id, const double *goal, double dt, const double *u_min, const double *u_max) {
  RefState *s = new RefState;
  s->model_id = model_id;
  s->dt = dt;
  if (goal) for (int i = 0; i < 4; i++) s->goal[i] = goal[i];
  const int m = model_id == 1 ? 2 : 1;
  if (u_min && u_max) {
    s->have_limits = true;
    for (int j = 0; j < m; j++) {
      s->u_min[j] = u_min[j];
      s->u_max[j] = u_max[j];
    }
  }
  rebuild(s);
  return s;
}

void ref_free(void *h) {
  RefState *s = (RefState *)h;
  delete s->solver;
  delete s;
}

void ref_dims(void *h, int *n, int *m) {
  RefState *s = (RefState *)h;
  *n = s->n;
  *m = s->m;
}

static void unpack(RefState *s, const double *x0, const double *u0, int T, VectorXd &x, VecOfVecXd &u) {
  x.resize(s->n);
  for (int i = 0; i < s->n; i++) x(i) = x0[i];
  u.clear();
  for (int t = 0; t < T; t++) {
    VectorXd ut(s->m);
    for (int j = 0; j < s->m; j++) ut(j) = u0[t * s->m + j];
    u.push_back(ut);
  }
}

double ref_init(void *h, const double *x0, const double *u0, int T) {
  RefState *s = (RefState *)h;
  VectorXd x;
  VecOfVecXd u;
  rebuild(s);
  unpack(s, x0, u0, T, x, u);
  return Probe::init(s, x, u);
}

int ref_iterate(void *h, int n_iters) { return Probe::iterate((RefState *)h, n_iters); }
int ref_backward_once(void *h, double lam, int recompute) { return Probe::backward_once((RefState *)h, lam, recompute); }
double ref_rollout_once(void *h, double alpha) { return Probe::rollout_once((RefState *)h, alpha); }

void ref_solve_native(void *h, const double *x0, const double *u0, int T) {
  RefState *s = (RefState *)h;
  VectorXd x;
  VecOfVecXd u;
  rebuild(s);
  unpack(s, x0, u0, T, x, u);
  Probe::solve_native(s, x, u);
}

// generate_trajectory(x_0): replica entry (continue with ref_iterate) and the native call
double ref_warm_start(void *h, const double *x0) {
  RefState *s = (RefState *)h;
  VectorXd x(s->n);
  for (int i = 0; i < s->n; i++) x(i) = x0[i];
  return Probe::warm_start(s, x);
}
void ref_warm_native(void *h, const double *x0) {
  RefState *s = (RefState *)h;
  VectorXd x(s->n);
  for (int i = 0; i < s->n; i++) x(i) = x0[i];
  Probe::warm_native(s, x);
}
void ref_resume(void *h) { Probe::resume((RefState *)h); }
void ref_resume_native(void *h) { Probe::resume_native((RefState *)h); }

int ref_get(void *h, int field, double *dst) { return Probe::get((RefState *)h, field, dst); }

// scalars: 0 lambda, 1 dlambda, 2 gnorm, 3 dcost, 4 expected, 5 alpha, 6 new_cost
double ref_scalar(void *h, int which) {
  RefState *s = (RefState *)h;
  switch (which) {
    case 0: return s->lam;
    case 1: return s->dlam;
    case 2: return s->gnorm;
    case 3: return s->dcost;
    case 4: return s->expected;
    case 5: return s->alpha;
    case 6: return s->new_cost;
    default: return 0;
  }
}

// ints: 0 iter, 1 loop_trips, 2 status, 3 alpha_index, 4 accepts, 5 rejects, 6 rollouts,
// 7 backward passes, 8 derivative sweeps, 9 T
long ref_int(void *h, int which) {
  RefState *s = (RefState *)h;
  switch (which) {
    case 0: return s->iter;
    case 1: return s->loop_trips;
    case 2: return s->status;
    case 3: return s->alpha_index;
    case 4: return s->n_accept;
    case 5: return s->n_reject;
    case 6: return s->n_rollouts;
    case 7: return s->n_backward;
    case 8: return s->n_deriv;
    case 9: return Probe::T(s);
    default: return -1;
  }
}

// --- direct probes of the leaf functions ------------------------------------------------

void ref_dynamics(void *h, const double *x, const double *u, double *dx) {
  RefState *s = (RefState *)h;
  VectorXd xv = Eigen::Map<const VectorXd>(x, s->n), uv = Eigen::Map<const VectorXd>(u, s->m);
  VectorXd r = s->model->dynamics(xv, uv);
  for (int i = 0; i < s->n; i++) dx[i] = r(i);
}
void ref_integrate(void *h, const double *x, const double *u, double dt, double *x1) {
  RefState *s = (RefState *)h;
  VectorXd xv = Eigen::Map<const VectorXd>(x, s->n), uv = Eigen::Map<const VectorXd>(u, s->m);
  VectorXd r = s->model->integrate_dynamics(xv, uv, dt);
  for (int i = 0; i < s->n; i++) x1[i] = r(i);
}
double ref_cost(void *h, const double *x, const double *u) {
  RefState *s = (RefState *)h;
  VectorXd xv = Eigen::Map<const VectorXd>(x, s->n), uv = Eigen::Map<const VectorXd>(u, s->m);
  return s->model->cost(xv, uv);
}
double ref_final_cost(void *h, const double *x) {
  RefState *s = (RefState *)h;
  VectorXd xv = Eigen::Map<const VectorXd>(x, s->n);
  return s->model->final_cost(xv);
}

// boxQP (src/boxqp.cpp:26-139).  Q row-major m x m.  R_free is written row-major into an
// m x m buffer using its own leading dimension *r_dim.  Returns res.result.
int ref_boxqp(int m, const double *Q, const double *c, const double *x0, const double *lo, const double *hi,
              double *x_opt, int *v_free, double *R_free, int *r_dim) {
  MatrixXd Qm(m, m);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < m; j++) Qm(i, j) = Q[i * m + j];
  VectorXd cv = Eigen::Map<const VectorXd>(c, m), xv = Eigen::Map<const VectorXd>(x0, m);
  VectorXd lv = Eigen::Map<const VectorXd>(lo, m), hv = Eigen::Map<const VectorXd>(hi, m);
  boxQPResult res = boxQP(Qm, cv, xv, lv, hv);
  for (int i = 0; i < m; i++) {
    x_opt[i] = res.x_opt(i);
    v_free[i] = res.v_free(i);
  }
  *r_dim = (int)res.R_free.rows();
  for (int i = 0; i < res.R_free.rows(); i++)
    for (int j = 0; j < res.R_free.cols(); j++) R_free[i * res.R_free.cols() + j] = res.R_free(i, j);
  return res.result;
}

// quadclamp_line_search (src/boxqp.cpp:143-178).  Returns failed flag.
int ref_quadclamp(int m, const double *x0, const double *dir, const double *Q, const double *c, const double *lo,
                  const double *hi, double *x_opt, double *v_opt, int *n_steps) {
  MatrixXd Qm(m, m);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < m; j++) Qm(i, j) = Q[i * m + j];
  VectorXd cv = Eigen::Map<const VectorXd>(c, m), xv = Eigen::Map<const VectorXd>(x0, m);
  VectorXd dv = Eigen::Map<const VectorXd>(dir, m);
  VectorXd lv = Eigen::Map<const VectorXd>(lo, m), hv = Eigen::Map<const VectorXd>(hi, m);
  lineSearchResult r = quadclamp_line_search(xv, dv, Qm, cv, lv, hv);
  for (int i = 0; i < m; i++) x_opt[i] = r.x_opt(i);
  *v_opt = r.v_opt;
  *n_steps = r.n_steps;
  return r.failed ? 1 : 0;
}

double ref_quadcost(int m, const double *Q, const double *c, const double *x) {
  MatrixXd Qm(m, m);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < m; j++) Qm(i, j) = Q[i * m + j];
  return quadCost(Qm, Eigen::Map<const VectorXd>(c, m), Eigen::Map<const VectorXd>(x, m));
}

// Finite-difference stencils (include/finite_diff.h) applied to the model's own functions
// at an arbitrary point: which = 0 Jacobian of integrate_dynamics wrt x (n x n), 1 wrt u (n x m),
// 2 gradient of cost wrt x, 3 wrt u, 4 gradient of final_cost, 5 Hessian of cost wrt x,
// 6 Hessian wrt u, 7 Hessian of final_cost.  Row-major out.  Returns element count.
int ref_fd(void *h, int which, const double *x, const double *u, double dt, double *out) {
  RefState *s = (RefState *)h;
  Model *M = s->model;
  const int n = s->n, m = s->m;
  VectorXd xv = Eigen::Map<const VectorXd>(x, n), uv = Eigen::Map<const VectorXd>(u, m);
  auto put = [&](const MatrixXd &A) {
    for (int i = 0; i < A.rows(); i++)
      for (int j = 0; j < A.cols(); j++) out[i * A.cols() + j] = A(i, j);
    return (int)A.size();
  };
  switch (which) {
    case 0: return put(finite_diff_jacobian([&](VectorXd a) { return M->integrate_dynamics(a, uv, dt); }, xv, n));
    case 1: return put(finite_diff_jacobian([&](VectorXd a) { return M->integrate_dynamics(xv, a, dt); }, uv, n));
    case 2: return put(finite_diff_gradient(std::function<double(VectorXd)>([&](VectorXd a) { return M->cost(a, uv); }), xv));
    case 3: return put(finite_diff_gradient(std::function<double(VectorXd)>([&](VectorXd a) { return M->cost(xv, a); }), uv));
    case 4: return put(finite_diff_gradient(std::function<double(VectorXd)>([&](VectorXd a) { return M->final_cost(a); }), xv));
    case 5: { MatrixXd H(n, n); finite_diff_hessian([&](VectorXd a) { return M->cost(a, uv); }, xv, H); return put(H); }
    case 6: { MatrixXd H(m, m); finite_diff_hessian([&](VectorXd a) { return M->cost(xv, a); }, uv, H); return put(H); }
    case 7: { MatrixXd H(n, n); finite_diff_hessian([&](VectorXd a) { return M->final_cost(a); }, xv, H); return put(H); }
    default: return -1;
  }
}

// Solve instances [b0, b1) of a batch one after another on the calling thread with the reference's
// own code (fresh iLQR object and lambda = dlambda = 1 per instance): the CPU baseline of bench.py,
// which forks one worker per core (the TU statics rule out threads).  x0[B][n], u0[B][T][m].
// max_trips < 0: run to termination.  Returns the number of loop trips executed.
long ref_solve_range(void *h, long b0, long b1, const double *x0, const double *u0, int T, int max_trips,
                     double *cost, int *iters, int *status) {
  RefState *s = (RefState *)h;
  long total = 0;
  for (long b = b0; b < b1; b++) {
    VectorXd x;
    VecOfVecXd u;
    rebuild(s);
    unpack(s, x0 + b * s->n, u0 + b * (long)T * s->m, T, x, u);
    Probe::init(s, x, u);
    {
      Silence q;  // src/ilqr_core.cpp:207 prints a warning from inside the line search
      Probe::iterate(s, max_trips < 0 ? maxIter + 1 : max_trips);
    }
    total += s->loop_trips;
    if (cost) Probe::get(s, 4, cost + (b - b0));
    if (iters) iters[b - b0] = s->loop_trips;
    if (status) status[b - b0] = s->status;
  }
  return total;
}

// The real libstdc++ generator, to pin include/ilqr_synth.h.
void ref_std_uniform(unsigned long long seed, int count, double *out) {
  std::mt19937_64 g(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  for (int i = 0; i < count; i++) out[i] = U(g);
}

}  // extern "C"
