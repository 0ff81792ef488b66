/*
 * params.h — host-side translation of the C-ABI descriptor (include/ilqr_b200.h) into the
 * scalar-typed constant block the kernels take by value.  The defaults are the reference's
 * compile-time constants: include/ilqr.h:14-25, include/boxqp.h:19-24,61-64,
 * include/finite_diff.h:9, src/derivatives.cpp:10, src/ilqr_core.cpp:154.
 */
#ifndef ILQR_PARAMS_H_
#define ILQR_PARAMS_H_

#include <string.h>

#include "../../include/ilqr_b200.h"
#include "ilqr_core.cuh"

namespace ilqr {

inline void default_params(ilqr_params *p) {
  static const double A[11] = {1.0000, 0.5012, 0.2512, 0.1259, 0.0631, 0.0316, 0.0158, 0.0079, 0.0040, 0.0020, 0.0010};
  memset(p, 0, sizeof(*p));
  p->max_iter = 100;
  p->n_alpha = 11;
  p->tol_fun = 1e-6;
  p->tol_grad = 1e-6;
  p->lambda_init = 1;
  p->dlambda_init = 1;
  p->lambda_factor = 1.6;
  p->lambda_max = 1e11;
  p->lambda_min = 1e-8;
  p->z_min = 0;
  p->grad_lambda_gate = 1e-5;
  for (int i = 0; i < 11; i++) p->alpha[i] = A[i];
  p->qp_max_iter = 100;
  p->qp_min_grad = 1e-8;
  p->qp_min_rel_improve = 1e-8;
  p->qp_step_dec = 0.6;
  p->qp_min_step = 1e-22;
  p->qp_armijo = 0.1;
  p->qp_clamp_tol = 1e-4;
  p->fd_eps = 1e-3;
}

/* models registered at run time (ilqr_register_model, ilqr_b200.cu) answer through this hook */
inline int (*g_user_model_info)(int model_id, int *n, int *m, double *u_min, double *u_max) = nullptr;

/* n, m and the model's own limits (Model::x_dims/u_dims/u_min/u_max) */
inline int model_info(int model_id, int *n, int *m, double *u_min, double *u_max) {
  if (model_id >= ILQR_MODEL_USER_BASE) return g_user_model_info ? g_user_model_info(model_id, n, m, u_min, u_max) : -1;
  if (model_id == ILQR_MODEL_ACROBOT) { /* include/acrobot.h:27-28,37 */
    *n = Acrobot::N;
    *m = Acrobot::M;
    if (u_min) u_min[0] = -5;
    if (u_max) u_max[0] = 5;
    return 0;
  }
  if (model_id == ILQR_MODEL_DOUBLE_INTEGRATOR) { /* include/double_integrator.h:16-17,25-26 */
    *n = DoubleIntegrator::N;
    *m = DoubleIntegrator::M;
    for (int j = 0; j < 2; j++) {
      if (u_min) u_min[j] = -0.5;
      if (u_max) u_max[j] = 0.5;
    }
    return 0;
  }
  return -1;
}

template <typename S>
inline int make_solve_params(const ilqr_desc &d, SolveParams<S> *out) {
  SolveParams<S> &P = *out;
  memset(&P, 0, sizeof(P));
  int n, m;
  double lo[ILQR_MAX_M] = {0}, hi[ILQR_MAX_M] = {0};
  if (model_info(d.model_id, &n, &m, lo, hi) != 0) return -1;
  const ilqr_params &p = d.params;
  if (p.n_alpha < 1 || p.n_alpha > kMaxAlpha || d.T < 1 || p.max_iter < 0 || p.qp_max_iter < 0) return -1;
  P.T = d.T;
  P.max_iter = p.max_iter;
  P.n_alpha = p.n_alpha;
  P.dt = S(d.dt);
  P.tol_fun = S(p.tol_fun);
  P.tol_grad = S(p.tol_grad);
  P.lambda_init = S(p.lambda_init);
  P.dlambda_init = S(p.dlambda_init);
  P.lambda_factor = S(p.lambda_factor);
  P.lambda_max = S(p.lambda_max);
  P.lambda_min = S(p.lambda_min);
  P.z_min = S(p.z_min);
  P.grad_lambda_gate = S(p.grad_lambda_gate);
  for (int i = 0; i < kMaxAlpha; i++) P.alpha[i] = S(p.alpha[i]);
  P.fd_eps = S(p.fd_eps);
  for (int j = 0; j < m; j++) {
    P.u_min[j] = S(d.override_limits ? d.u_min[j] : lo[j]);
    P.u_max[j] = S(d.override_limits ? d.u_max[j] : hi[j]);
  }
  if (d.model_id == ILQR_MODEL_ACROBOT) {
    /* the goal of include/acrobot.h:20-21 (the literal 3.1415, not pi) unless the caller passes the goal of its own
     * Acrobot object (the C++ host layer does, so an edited acrobot.h cannot silently diverge from the device) */
    bool given = false;
    for (int i = 0; i < 4; i++) given = given || d.model_params[i] != 0.0;
    if (given) {
      for (int i = 0; i < 4; i++) P.mp[i] = S(d.model_params[i]);
    } else {
      P.mp[0] = S(3.1415);
    }
  } else if (d.model_id >= ILQR_MODEL_USER_BASE) {
    for (int i = 0; i < 16; i++) P.mp[i] = S(d.model_params[i]);
  } else {
    for (int i = 0; i < 4; i++) P.mp[i] = S(d.model_params[i]);
  }
  P.qp.max_iter = p.qp_max_iter;
  P.qp.min_grad = S(p.qp_min_grad);
  P.qp.min_rel_improve = S(p.qp_min_rel_improve);
  P.qp.step_dec = S(p.qp_step_dec);
  P.qp.min_step = S(p.qp_min_step);
  P.qp.armijo = S(p.qp_armijo);
  P.qp.clamp_tol = S(p.qp_clamp_tol);
  P.flags = d.flags & (kFlagClampRollout | kFlagAnalyticDyn);
  P.bulk_f = 0;
  return 0;
}

}  // namespace ilqr
#endif
