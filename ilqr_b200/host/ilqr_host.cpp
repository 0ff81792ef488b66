// ilqr_host.cpp — `class iLQR` of ilqr_b200/host/ilqr.h on top of the C ABI (include/ilqr_b200.h).
// Host orchestration only: marshal Eigen containers to the ABI's row-major arrays, call the library,
// marshal back.  No solver arithmetic happens here.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <iostream>

#include <random>
#include <typeinfo>

// The reference's model headers keep their parameters private (acrobot.h:112-118,
// double_integrator.h:50-53).  The device twin of DoubleIntegrator needs the goal the object was
// built with; reading it here keeps those headers unchanged.
#include "common.h"  // everything the model headers pull in (std, Eigen) is included before the next line
#include "model.h"
#define private public
#include "acrobot.h"
#include "double_integrator.h"
#undef private

#include "../../include/ilqr_b200.h"
#include "../csrc/models.cuh"  // host instantiation of the device twins, for the construction-time check
#include "batch_solver.h"
#include "ilqr.h"

#include <map>
#include <typeindex>

namespace {

void check(int rc, ilqr_handle *h, const char *what) {
  if (rc != ILQR_OK) throw std::runtime_error(std::string(what) + ": " + ilqr_last_error(h));
}

template <class Twin>
void verify_twin(Model &m, const double *mp, double dt) {
  std::mt19937_64 g(7);
  std::uniform_real_distribution<double> U(-3.0, 3.0);
  for (int probe = 0; probe < 16; probe++) {
    VectorXd x(Twin::N), u(Twin::M);
    double xa[Twin::N], ua[Twin::M], x1[Twin::N];
    for (int i = 0; i < Twin::N; i++) x(i) = xa[i] = U(g);
    for (int i = 0; i < Twin::M; i++) u(i) = ua[i] = U(g);
    const VectorXd ref = m.integrate_dynamics(x, u, dt);
    ilqr::integrate<Twin, double>(xa, ua, mp, dt, x1);
    double err = 0;
    for (int i = 0; i < Twin::N; i++) err = std::max(err, std::fabs(ref(i) - x1[i]));
    err = std::max(err, std::fabs(m.cost(x, u) - Twin::cost(xa, ua, mp)) / (1 + std::fabs(m.cost(x, u))));
    err = std::max(err, std::fabs(m.final_cost(x) - Twin::final_cost(xa, mp)) / (1 + std::fabs(m.final_cost(x))));
    if (!(err < 1e-11))
      throw std::runtime_error("iLQR: the device twin of this Model disagrees with the host object (edited model header?)");
  }
}

struct UserTwin {
  int model_id = -1;
  std::vector<double> params;
  std::string pending_name, pending_source;
};
std::map<std::type_index, UserTwin> &user_twins() {
  static std::map<std::type_index, UserTwin> m;
  return m;
}

/* the device twin of a user model against the host object, through the GPU: one open-loop Euler step and the costs
 * (ilqr_set_initial = iLQR::init_traj on a horizon of one) on a few probe points */
void verify_user_twin(Model &m, int model_id, const double *mp, double dt) {
  const int n = m.x_dims, mu = m.u_dims, B = 4;
  ilqr_desc d;
  memset(&d, 0, sizeof(d));
  d.model_id = model_id;
  d.dtype = ILQR_F64;
  d.cost_deriv = ILQR_COST_FD;
  d.T = 1;
  d.B = B;
  d.dt = dt;
  for (int i = 0; i < 16; i++) d.model_params[i] = mp[i];
  ilqr_default_params(&d.params);
  ilqr_handle *h = nullptr;
  if (ilqr_create(&d, &h) != ILQR_OK) throw std::runtime_error(std::string("iLQR: ") + ilqr_last_error(nullptr));
  std::vector<double> x0((size_t)B * n), u0((size_t)B * mu), xs((size_t)B * 2 * n), cost(B);
  for (int b = 0; b < B; b++) {
    for (int i = 0; i < n; i++) x0[(size_t)b * n + i] = 0.3 * std::sin(1.0 + 2.1 * b + 0.7 * i);
    for (int j = 0; j < mu; j++) u0[(size_t)b * mu + j] = 0.2 * std::cos(0.5 + 1.3 * b + j);
  }
  int rc = ilqr_set_initial(h, x0.data(), u0.data(), 0);
  if (rc == ILQR_OK) rc = ilqr_get(h, ILQR_F_XS, xs.data(), 0);
  if (rc == ILQR_OK) rc = ilqr_get(h, ILQR_F_COST, cost.data(), 0);
  const std::string err = rc == ILQR_OK ? "" : ilqr_last_error(h);
  ilqr_destroy(h);
  if (rc != ILQR_OK) throw std::runtime_error("iLQR: the device twin did not run: " + err);
  double worst = 0;
  for (int b = 0; b < B; b++) {
    VectorXd x(n), u(mu);
    for (int i = 0; i < n; i++) x(i) = x0[(size_t)b * n + i];
    for (int j = 0; j < mu; j++) u(j) = u0[(size_t)b * mu + j];
    const VectorXd x1 = m.integrate_dynamics(x, u, dt);
    for (int i = 0; i < n; i++) worst = std::max(worst, std::fabs(x1(i) - xs[((size_t)b * 2 + 1) * n + i]));
    const double c = m.cost(x, u) + m.final_cost(x1);
    worst = std::max(worst, std::fabs(c - cost[b]) / (1 + std::fabs(c)));
  }
  if (!(worst < 1e-9))
    throw std::runtime_error("iLQR: the registered device twin disagrees with the host Model object (max error " + std::to_string(worst) + ")");
}

}  // namespace

iLQR::iLQR(Model *p_dyn, double timeDelta) : dt(timeDelta) {
  model.reset(p_dyn);  // takes ownership, like the reference (include/ilqr.h:30-31)
  if (Acrobot *a = dynamic_cast<Acrobot *>(p_dyn)) {
    model_id = ILQR_MODEL_ACROBOT;
    for (int i = 0; i < 4; i++) model_params[i] = a->goal(i);
    verify_twin<ilqr::Acrobot>(*p_dyn, model_params, dt);
  } else if (DoubleIntegrator *d = dynamic_cast<DoubleIntegrator *>(p_dyn)) {
    model_id = ILQR_MODEL_DOUBLE_INTEGRATOR;
    for (int i = 0; i < 4; i++) model_params[i] = d->goal(i);
    verify_twin<ilqr::DoubleIntegrator>(*p_dyn, model_params, dt);
  } else {
    const auto it = user_twins().find(std::type_index(typeid(*p_dyn)));
    if (it == user_twins().end())
      throw std::runtime_error(std::string("iLQR: no device twin for Model subclass ") + typeid(*p_dyn).name() +
                               " (built in: ilqr_b200/csrc/models.cuh; your own: iLQR::register_device_twin); there is no CPU fallback");
    if (it->second.model_id < 0) { /* first object of this type: n, m and the limits are the object's (model.h:17-20) */
      double lo[ILQR_MAX_M] = {0}, hi[ILQR_MAX_M] = {0};
      for (int j = 0; j < p_dyn->u_dims && j < ILQR_MAX_M; j++) {
        lo[j] = p_dyn->u_min(j);
        hi[j] = p_dyn->u_max(j);
      }
      int32_t id = -1;
      if (ilqr_register_model(it->second.pending_name.c_str(), it->second.pending_source.c_str(), p_dyn->x_dims, p_dyn->u_dims,
                              lo, hi, &id) != ILQR_OK)
        throw std::runtime_error(std::string("iLQR: ") + ilqr_last_error(nullptr));
      it->second.model_id = id;
    }
    model_id = it->second.model_id;
    for (size_t i = 0; i < it->second.params.size() && i < 16; i++) model_params[i] = it->second.params[i];
    verify_user_twin(*p_dyn, model_id, model_params, dt);
  }
}

void iLQR::register_device_twin(const std::type_info &type, const char *struct_name, const char *cuda_source,
                                const std::vector<double> &params) {
  UserTwin t;
  t.params = params;
  t.pending_name = struct_name;
  t.pending_source = cuda_source;
  user_twins()[std::type_index(type)] = t; /* n, m and the limits are the object's: registration with the library happens at first construction */
}

iLQR::~iLQR() {
  ilqr_destroy(h);
  delete multi;
}

void iLQR::create(long B, int T_) {
  ilqr_desc d;
  memset(&d, 0, sizeof(d));
  d.model_id = model_id;
  d.dtype = ILQR_F64;
  d.cost_deriv = cost_deriv;
  d.device = 0;
  d.T = T_;
  d.B = B;
  d.dt = dt;
  d.override_limits = 1;  // Model::u_min / u_max are public data the caller may have changed (model.h:17)
  for (int j = 0; j < model->u_dims; j++) {
    d.u_min[j] = model->u_min(j);
    d.u_max[j] = model->u_max(j);
  }
  for (int i = 0; i < 16; i++) d.model_params[i] = model_params[i];
  ilqr_default_params(&d.params);
  d.params.max_iter = maxIter;
  d.flags = flags;
  // everything the handle was created with is in the descriptor: a changed limit (the reference reads model->u_min /
  // u_max on every backward pass, src/ilqr_core.cpp:369), maxIter, cost_deriv or flag makes a new handle
  if (h && memcmp(&d, &hdesc, sizeof(d)) == 0) return;
  ilqr_destroy(h);
  h = nullptr;
  check(ilqr_create(&d, &h), nullptr, "ilqr_create");
  hdesc = d;
  hB = B;
  hT = T_;
}

double iLQR::init_traj(const VectorXd &x_0, const VecOfVecXd &u_0) {
  T = (int)u_0.size();
  const int n = model->x_dims, m = model->u_dims;
  create(1, T);
  std::vector<double> x0(n), u0((size_t)T * m);
  for (int i = 0; i < n; i++) x0[i] = x_0(i);
  for (int t = 0; t < T; t++)
    for (int j = 0; j < m; j++) u0[(size_t)t * m + j] = u_0[t](j);
  check(ilqr_set_initial(h, x0.data(), u0.data(), 0), h, "ilqr_set_initial");
  fetch_single();
  if (!quiet) printf("Initial cost: %f\n", cost_s);
  return cost_s;
}

void iLQR::fetch_single() {
  const int n = model->x_dims, m = model->u_dims;
  std::vector<double> bx((size_t)(T + 1) * n), bu((size_t)T * m), bk((size_t)T * m), bK((size_t)T * m * n);
  check(ilqr_get(h, ILQR_F_XS, bx.data(), 0), h, "ilqr_get");
  check(ilqr_get(h, ILQR_F_US, bu.data(), 0), h, "ilqr_get");
  check(ilqr_get(h, ILQR_F_KFF, bk.data(), 0), h, "ilqr_get");
  check(ilqr_get(h, ILQR_F_K, bK.data(), 0), h, "ilqr_get");
  check(ilqr_get(h, ILQR_F_COST, &cost_s, 0), h, "ilqr_get");
  int32_t it = 0, st = 0;
  check(ilqr_get(h, ILQR_F_ITERS, &it, 0), h, "ilqr_get");
  check(ilqr_get(h, ILQR_F_STATUS, &st, 0), h, "ilqr_get");
  iterations = it;
  status = st;
  xs.assign(T + 1, VectorXd::Zero(n));
  us.assign(T, VectorXd::Zero(m));
  k.assign(T, VectorXd::Zero(m));
  K.assign(T, MatrixXd::Zero(m, n));
  for (int t = 0; t <= T; t++)
    for (int i = 0; i < n; i++) xs[t](i) = bx[(size_t)t * n + i];
  for (int t = 0; t < T; t++) {
    for (int j = 0; j < m; j++) {
      us[t](j) = bu[(size_t)t * m + j];
      k[t](j) = bk[(size_t)t * m + j];
      for (int i = 0; i < n; i++) K[t](j, i) = bK[((size_t)t * m + j) * n + i];
    }
  }
}

void iLQR::generate_trajectory(const VectorXd &x_0, const VecOfVecXd &u0) {
  init_traj(x_0, u0);
  generate_trajectory();
}

void iLQR::generate_trajectory(const VectorXd &x_0) {
  if (!h) throw std::runtime_error("iLQR::generate_trajectory(x_0): no previous solve to warm-start from");
  std::vector<double> x0(model->x_dims);
  for (int i = 0; i < model->x_dims; i++) x0[i] = x_0(i);
  check(ilqr_warm_start(h, x0.data(), 0), h, "ilqr_warm_start");
  generate_trajectory();
}

void iLQR::generate_trajectory() {
  if (!h) throw std::runtime_error("iLQR::generate_trajectory(): call init_traj first");
  // every call re-enters the loop with iter = 0 and fresh derivatives, like the reference (src/ilqr_core.cpp:88-102);
  // after init_traj / a warm start this changes nothing
  check(ilqr_resume(h), h, "ilqr_resume");
  check(ilqr_solve(h), h, "ilqr_solve");
  fetch_single();
  if (!quiet) {
    static const char *why[] = {"running", "gradient norm < tolGrad", "cost change < tolFun", "lambda > lambdaMax", "max iterations"};
    printf("\nEXIT: %s after %d iterations, cost %.12g\n", why[status], iterations, cost_s);
  }
  output_to_csv("ilqr_result.csv");  // the reference does this at the end of every solve (src/ilqr_core.cpp:300)
}

// src/ilqr_core.cpp:414-431, byte for byte: header "x1, ..., xn, u0, ..., um" (the reference names one control column
// too many, :418-419), one row per knot with the last control written "%f\n" (:421-425), and the terminal row —
// states only — ending in ", " with NO newline (:427); plot_results.py:15 tells the terminal row by that blank.
void ilqr_write_csv(FILE *f, int T, int n, int m, const double *xs, const double *us) {
  for (int i = 1; i <= n; i++) fprintf(f, "x%d, ", i);
  for (int j = 0; j < m; j++) fprintf(f, "u%d, ", j);
  fprintf(f, "u%d\n", m);
  for (int t = 0; t < T; t++) {
    for (int i = 0; i < n; i++) fprintf(f, "%f, ", xs[(size_t)t * n + i]);
    for (int j = 0; j + 1 < m; j++) fprintf(f, "%f, ", us[(size_t)t * m + j]);
    fprintf(f, "%f\n", us[(size_t)t * m + m - 1]);
  }
  for (int i = 0; i < n; i++) fprintf(f, "%f, ", xs[(size_t)T * n + i]);
}

void iLQR::output_to_csv(const std::string filename) {
  FILE *f = fopen(filename.c_str(), "w");
  if (!f) return;
  const int n = model->x_dims, m = model->u_dims;
  std::vector<double> bx((size_t)(T + 1) * n), bu((size_t)T * m);
  for (int t = 0; t <= T; t++)
    for (int i = 0; i < n; i++) bx[(size_t)t * n + i] = xs[t](i);
  for (int t = 0; t < T; t++)
    for (int j = 0; j < m; j++) bu[(size_t)t * m + j] = us[t](j);
  ilqr_write_csv(f, T, n, m, bx.data(), bu.data());
  fclose(f);
  if (!quiet) std::cout << "Saved iLQR result to " << filename << std::endl;  // :430
}

// One trajectory of the last solve_batch in the reference's CSV format.
void iLQR::output_to_csv(const std::string filename, int b) const {
  const int n = model->x_dims, m = model->u_dims;
  if (b < 0 || (size_t)(b + 1) * (T + 1) * n > bxs.size()) throw std::runtime_error("iLQR::output_to_csv: no such batch entry");
  FILE *f = fopen(filename.c_str(), "w");
  if (!f) return;
  ilqr_write_csv(f, T, n, m, bxs.data() + (size_t)b * (T + 1) * n, bus.data() + (size_t)b * T * m);
  fclose(f);
}

// The whole batch in one binary file (little endian): a 64-byte header
//   char magic[8] = "ILQRB200"; uint32 version = 1, dtype (0 = f64), n, m, T, reserved; uint64 B; 24 bytes of zeros
// then xs [B][T+1][n], us [B][T][m], cost [B] as f64 and iterations [B], status [B] as int32 — the ABI's own layouts,
// so the file is what ilqr_get returned.  ilqr_b200/export.py reads it.
void iLQR::export_batch(const std::string filename) const {
  const int n = model->x_dims, m = model->u_dims;
  const uint64_t B = batch_cost.size();
  if (B == 0) throw std::runtime_error("iLQR::export_batch: no batch has been solved");
  FILE *f = fopen(filename.c_str(), "wb");
  if (!f) throw std::runtime_error("iLQR::export_batch: cannot open " + filename);
  unsigned char head[64] = {0};
  memcpy(head, "ILQRB200", 8);
  const uint32_t w[6] = {1u, 0u, (uint32_t)n, (uint32_t)m, (uint32_t)T, 0u};
  memcpy(head + 8, w, sizeof(w));
  memcpy(head + 32, &B, sizeof(B));
  bool ok = fwrite(head, 1, 64, f) == 64;
  ok = ok && fwrite(bxs.data(), sizeof(double), bxs.size(), f) == bxs.size();
  ok = ok && fwrite(bus.data(), sizeof(double), bus.size(), f) == bus.size();
  ok = ok && fwrite(batch_cost.data(), sizeof(double), B, f) == B;
  ok = ok && fwrite(batch_iters.data(), sizeof(int), B, f) == B;
  ok = ok && fwrite(batch_status.data(), sizeof(int), B, f) == B;
  fclose(f);
  if (!ok) throw std::runtime_error("iLQR::export_batch: short write to " + filename);
}

std::vector<double> iLQR::solve_batch(const std::vector<VectorXd> &X0, const std::vector<VecOfVecXd> &U0) {
  const long B = (long)X0.size();
  if (B == 0 || U0.size() != X0.size()) throw std::runtime_error("iLQR::solve_batch: X0 and U0 must have the same non-zero length");
  const int n = model->x_dims, m = model->u_dims;
  T = (int)U0[0].size();
  std::vector<double> x0((size_t)B * n), u0((size_t)B * T * m);
  for (long b = 0; b < B; b++) {
    if ((int)U0[b].size() != T) throw std::runtime_error("iLQR::solve_batch: ragged horizons");
    for (int i = 0; i < n; i++) x0[(size_t)b * n + i] = X0[b](i);
    for (int t = 0; t < T; t++)
      for (int j = 0; j < m; j++) u0[((size_t)b * T + t) * m + j] = U0[b][t](j);
  }
  if (!devices.empty()) { /* sharded over several GPUs from this one process (batch_solver.h) */
    ilqr_desc d;
    memset(&d, 0, sizeof(d));
    d.model_id = model_id;
    d.dtype = ILQR_F64;
    d.cost_deriv = cost_deriv;
    d.T = T;
    d.dt = dt;
    d.override_limits = 1;
    for (int j = 0; j < m; j++) {
      d.u_min[j] = model->u_min(j);
      d.u_max[j] = model->u_max(j);
    }
    for (int i = 0; i < 16; i++) d.model_params[i] = model_params[i];
    ilqr_default_params(&d.params);
    d.params.max_iter = maxIter;
    d.flags = flags;
    if (!multi || multi_devices != devices || multi_T != T || memcmp(&d, &hdesc, sizeof(d)) != 0) {
      delete multi;
      multi = new BatchSolver(d, devices);
      multi_devices = devices;
      multi_T = T;
      hdesc = d;
    }
    std::vector<double> cost(B);
    std::vector<int32_t> iters(B), stat(B);
    multi->solve(x0.data(), u0.data(), B, cost.data(), iters.data());
    bxs.resize((size_t)B * (T + 1) * n);
    bus.resize((size_t)B * T * m);
    multi->get(ILQR_F_XS, bxs.data());
    multi->get(ILQR_F_US, bus.data());
    multi->get(ILQR_F_STATUS, stat.data());
    batch_iters.assign(iters.begin(), iters.end());
    batch_status.assign(stat.begin(), stat.end());
    batch_cost = cost;
    return cost;
  }
  create(B, T);
  check(ilqr_set_initial(h, x0.data(), u0.data(), 0), h, "ilqr_set_initial");
  check(ilqr_solve(h), h, "ilqr_solve");
  std::vector<double> cost(B);
  std::vector<int32_t> iters(B);
  bxs.resize((size_t)B * (T + 1) * n);
  bus.resize((size_t)B * T * m);
  check(ilqr_get(h, ILQR_F_COST, cost.data(), 0), h, "ilqr_get");
  check(ilqr_get(h, ILQR_F_ITERS, iters.data(), 0), h, "ilqr_get");
  check(ilqr_get(h, ILQR_F_XS, bxs.data(), 0), h, "ilqr_get");
  check(ilqr_get(h, ILQR_F_US, bus.data(), 0), h, "ilqr_get");
  std::vector<int32_t> stat(B);
  check(ilqr_get(h, ILQR_F_STATUS, stat.data(), 0), h, "ilqr_get");
  batch_iters.assign(iters.begin(), iters.end());
  batch_status.assign(stat.begin(), stat.end());
  batch_cost = cost;
  return cost;
}

VecOfVecXd iLQR::batch_xs(int b) const {
  const int n = model->x_dims;
  VecOfVecXd out(T + 1, VectorXd::Zero(n));
  for (int t = 0; t <= T; t++)
    for (int i = 0; i < n; i++) out[t](i) = bxs[((size_t)b * (T + 1) + t) * n + i];
  return out;
}
VecOfVecXd iLQR::batch_us(int b) const {
  const int m = model->u_dims;
  VecOfVecXd out(T, VectorXd::Zero(m));
  for (int t = 0; t < T; t++)
    for (int j = 0; j < m; j++) out[t](j) = bus[((size_t)b * T + t) * m + j];
  return out;
}
