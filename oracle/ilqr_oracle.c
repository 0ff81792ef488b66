/*
 * ilqr_oracle.c — CPU ORACLE (plain C99), TEST INFRASTRUCTURE.  See ilqr_oracle.h.
 *
 * Restates, statement for statement, the solve path of kazuotani14/iLQR.  Citations are
 * `file:line` into the reference tree.  Arithmetic is written in the order the reference's
 * expressions evaluate (Eigen dynamic products accumulate k = 0..n-1 in order); the file is
 * compiled with -ffp-contract=off so no FMA sneaks in.  Where the reference has a quirk the
 * quirk is kept and labelled QUIRK — "fixing" any of them changes results beyond 1e-6.
 */
#include "ilqr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NX ILQR_MAX_N
#define NU ILQR_MAX_M

/* ------------------------------------------------------------------------------------------
 * defaults: include/ilqr.h:14-25, include/boxqp.h:19-24,61-64, include/finite_diff.h:9
 * ---------------------------------------------------------------------------------------- */
static void default_params(ilqr_params *p) {
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
void orc_default_params(ilqr_params *p) { default_params(p); }

struct orc_solver {
  ilqr_desc d;
  int n, m, T;
  double umin[NU], umax[NU];
  double goal[NX];
  double *x0, *xs, *us;
  double *fx, *fu, *cx, *cu, *cxx, *cxu, *cuu;
  double *Vx, *Vxx, *k, *K;
  double dV[2];
  double cost_s;
  /* the reference's TU statics (include/ilqr.h:17-18), one pair per instance here */
  double lam, dlam;
  int flgChange, iter, loop_trips, status, diverge;
  double gnorm, dcost, expected, alpha, new_cost;
  int alpha_index;
  long n_accept, n_reject, n_rollouts, n_backward, n_deriv;
};

/* ------------------------------------------------------------------------------------------
 * Models
 * ---------------------------------------------------------------------------------------- */

/* Acrobot::dynamics, include/acrobot.h:72-81 with H :43-51, C :53-61, G :63-70.
 * Parameters (:19,23-25): I1=I2=l1=l2=m1=m2=1, lc1=lc2=0.5, g=9.81.  The 2x2 inverse is
 * Eigen's fixed-size closed form (Eigen/src/LU/InverseImpl.h:76-94). */
static void acrobot_dynamics(const double *x, const double *u, double *dx) {
  const double I1 = 1, I2 = 1, l1 = 1, l2 = 1, m1 = 1, m2 = 1, g = 9.81;
  const double lc1 = 0.5 * l1, lc2 = 0.5 * l2;
  const double q0 = x[0], q1 = x[1], qd0 = x[2], qd1 = x[3];
  const double c2 = cos(q1);
  const double H00 = I1 + I2 + m2 * l1 * l1 + 2 * m2 * l1 * lc2 * c2;
  const double H01 = I2 + m2 * l1 * lc2 * c2;
  const double H10 = I2 + m2 * l1 * lc2 * c2;
  const double H11 = I2;
  const double s2 = sin(q1);
  const double C00 = -2 * m2 * l1 * lc2 * s2 * qd1;
  const double C01 = -m2 * l2 * lc2 * s2 * qd1;
  const double C10 = m2 * l1 * lc2 * s2 * qd0;
  const double C11 = 0;
  const double s1 = sin(q0);
  const double s1p2 = sin(q0 + q1);
  const double G0 = m1 * g * lc1 * s1 + m2 * g * (l1 * s1 + lc2 * s1p2);
  const double G1 = m2 * g * lc2 * s1p2;
  /* Vector2d(0,u) - C*qdot - G */
  const double r0 = (0.0 - (C00 * qd0 + C01 * qd1)) - G0;
  const double r1 = (u[0] - (C10 * qd0 + C11 * qd1)) - G1;
  const double det = H00 * H11 - H10 * H01;
  const double invdet = 1.0 / det;
  const double Hi00 = H11 * invdet, Hi10 = -H10 * invdet, Hi01 = -H01 * invdet, Hi11 = H00 * invdet;
  dx[0] = qd0;
  dx[1] = qd1;
  dx[2] = Hi00 * r0 + Hi01 * r1;
  dx[3] = Hi10 * r0 + Hi11 * r1;
}
/* Acrobot::cost :83-92 (Ks=Kd=0, Kr=0.1) and final_cost :94-100 (Ks=Kd=20); goal :20-21 is the
 * literal 3.1415, not pi. */
static double acrobot_cost(const double *goal, const double *x, const double *u) {
  const double e0 = goal[0] - x[0], e1 = goal[1] - x[1], e2 = goal[2] - x[2], e3 = goal[3] - x[3];
  const double Ks = 0.0, Kd = 0.0, Kr = 0.1;
  return Ks * Ks * (e0 * e0 + e1 * e1) + Kd * Kd * (e2 * e2 + e3 * e3) + Kr * Kr * (u[0] * u[0]);
}
static double acrobot_final_cost(const double *goal, const double *x) {
  const double e0 = goal[0] - x[0], e1 = goal[1] - x[1], e2 = goal[2] - x[2], e3 = goal[3] - x[3];
  const double Ks = 20.0, Kd = 20.0;
  return Ks * Ks * (e0 * e0 + e1 * e1) + Kd * Kd * (e2 * e2 + e3 * e3);
}

/* DoubleIntegrator, include/double_integrator.h:29-48; Hx = diag(1,1,.2,.2), Hu = I, mass 1. */
static void di_dynamics(const double *x, const double *u, double *dx) {
  const double mass = 1.0;
  dx[0] = x[2];
  dx[1] = x[3];
  dx[2] = u[0] / mass;
  dx[3] = u[1] / mass;
}
static const double DI_HX[4] = {1, 1, 0.2, 0.2};
static double di_quad(const double *goal, const double *x, double scale) {
  double acc = 0;
  for (int i = 0; i < 4; i++) {
    const double e = goal[i] - x[i];
    acc += (e * (scale * DI_HX[i])) * e; /* (e^T Hx) e, Hx diagonal: off-diagonal products are exact zeros */
  }
  return acc;
}
static double di_cost(const double *goal, const double *x, const double *u) {
  return di_quad(goal, x, 1.0) + (u[0] * u[0] + u[1] * u[1]);
}
static double di_final_cost(const double *goal, const double *x) { return di_quad(goal, x, 10.0); }

/* OPT-IN, not in the reference (ILQR_FLAG_ANALYTIC_DYN; the reference only has finite differences and lists analytic
 * Jacobians as future work, notes.md:15,45): closed-form Jacobian of Acrobot::dynamics, A[i][j] = d dx_i / d x_j (4 x 4
 * row-major), Bm[i] = d dx_i / d u.  With qdd = H^-1 r:  d qdd / d z = H^-1 (d r / d z - (d H / d z) qdd). */
static void acrobot_dynamics_jac(const double *x, const double *u, double *A, double *Bm) {
  const double I1 = 1, I2 = 1, l1 = 1, l2 = 1, m1 = 1, m2 = 1, g = 9.81;
  const double lc1 = 0.5 * l1, lc2 = 0.5 * l2;
  const double q0 = x[0], q1 = x[1], qd0 = x[2], qd1 = x[3];
  const double c2 = cos(q1), s2 = sin(q1), c1 = cos(q0), c12 = cos(q0 + q1);
  const double a = m2 * l1 * lc2, b = m2 * l2 * lc2;
  const double H00 = I1 + I2 + m2 * l1 * l1 + 2 * a * c2, H01 = I2 + a * c2, H11 = I2;
  const double det = H00 * H11 - H01 * H01;
  const double invdet = 1.0 / det;
  const double Hi00 = H11 * invdet, Hi01 = -H01 * invdet, Hi11 = H00 * invdet;
  double dx[4];
  acrobot_dynamics(x, u, dx);
  const double qdd0 = dx[2], qdd1 = dx[3];
  double dr0[4], dr1[4]; /* d r / d z - (d H / d z) qdd for z = q0, q1, qd0, qd1 */
  dr0[0] = -(m1 * g * lc1 * c1 + m2 * g * (l1 * c1 + lc2 * c12));
  dr1[0] = -(m2 * g * lc2 * c12);
  dr0[1] = ((2 * a * qd0 * qd1 + b * qd1 * qd1) * c2 - m2 * g * lc2 * c12) + s2 * (2 * a * qdd0 + a * qdd1);
  dr1[1] = (-(a * c2 * qd0 * qd0) - m2 * g * lc2 * c12) + s2 * (a * qdd0);
  dr0[2] = 2 * a * s2 * qd1;
  dr1[2] = -(2 * a * s2 * qd0);
  dr0[3] = 2 * a * s2 * qd0 + 2 * b * s2 * qd1;
  dr1[3] = 0.0;
  for (int i = 0; i < 16; i++) A[i] = 0.0;
  A[0 * 4 + 2] = 1.0;
  A[1 * 4 + 3] = 1.0;
  for (int j = 0; j < 4; j++) {
    A[2 * 4 + j] = Hi00 * dr0[j] + Hi01 * dr1[j];
    A[3 * 4 + j] = Hi01 * dr0[j] + Hi11 * dr1[j];
  }
  Bm[0] = 0.0;
  Bm[1] = 0.0;
  Bm[2] = Hi01;
  Bm[3] = Hi11;
}
static void di_dynamics_jac(double *A, double *Bm) {
  const double mass = 1.0;
  for (int i = 0; i < 16; i++) A[i] = 0.0;
  for (int i = 0; i < 8; i++) Bm[i] = 0.0;
  A[0 * 4 + 2] = 1.0;
  A[1 * 4 + 3] = 1.0;
  Bm[2 * 2 + 0] = 1.0 / mass;
  Bm[3 * 2 + 1] = 1.0 / mass;
}

static void model_dynamics(const orc_solver *s, const double *x, const double *u, double *dx) {
  if (s->d.model_id == ILQR_MODEL_ACROBOT) acrobot_dynamics(x, u, dx);
  else di_dynamics(x, u, dx);
}
static double model_cost(const orc_solver *s, const double *x, const double *u) {
  return s->d.model_id == ILQR_MODEL_ACROBOT ? acrobot_cost(s->goal, x, u) : di_cost(s->goal, x, u);
}
static double model_final_cost(const orc_solver *s, const double *x) {
  return s->d.model_id == ILQR_MODEL_ACROBOT ? acrobot_final_cost(s->goal, x) : di_final_cost(s->goal, x);
}
/* Model::integrate_dynamics, include/model.h:12-15: x + dynamics(x,u)*dt */
static void model_integrate(const orc_solver *s, const double *x, const double *u, double dt, double *x1) {
  double dx[NX];
  model_dynamics(s, x, u, dx);
  for (int i = 0; i < s->n; i++) x1[i] = x[i] + dx[i] * dt;
}

/* Closed-form cost derivatives of the two model twins (used when cost_deriv == ANALYTIC; the
 * reference has no analytic path, BASELINE configs 2/3/5 ask for one).  terminal != 0: final_cost. */
static void model_cost_derivs(const orc_solver *s, const double *x, const double *u, int terminal, double *cx,
                              double *cu, double *cxx, double *cxu, double *cuu) {
  const int n = s->n, m = s->m;
  memset(cx, 0, sizeof(double) * n);
  memset(cu, 0, sizeof(double) * m);
  memset(cxx, 0, sizeof(double) * n * n);
  memset(cxu, 0, sizeof(double) * n * m);
  memset(cuu, 0, sizeof(double) * m * m);
  if (s->d.model_id == ILQR_MODEL_ACROBOT) {
    if (terminal) {
      for (int i = 0; i < 4; i++) {
        cx[i] = -800.0 * (s->goal[i] - x[i]);
        cxx[i * 4 + i] = 800.0;
      }
    } else {
      const double w = 0.1 * 0.1;
      cu[0] = 2 * w * u[0];
    }
    cuu[0] = 2 * (0.1 * 0.1);
  } else {
    const double sc = terminal ? 10.0 : 1.0;
    for (int i = 0; i < 4; i++) {
      cx[i] = -2.0 * (sc * DI_HX[i]) * (s->goal[i] - x[i]);
      cxx[i * 4 + i] = 2.0 * (sc * DI_HX[i]);
    }
    if (!terminal) {
      cu[0] = 2 * u[0];
      cu[1] = 2 * u[1];
    }
    cuu[0] = cuu[3] = 2.0;
  }
}

/* ------------------------------------------------------------------------------------------
 * Finite differences: include/finite_diff.h
 * ---------------------------------------------------------------------------------------- */
typedef double (*scalar_fn)(const orc_solver *, const double *v, const void *ctx);

/* finite_diff_gradient, finite_diff.h:22-33 */
static void fd_gradient(const orc_solver *s, scalar_fn f, const void *ctx, const double *x, int nd, double *dx) {
  const double eps = s->d.params.fd_eps;
  double plus[NX], minus[NX];
  for (int i = 0; i < nd; i++) {
    memcpy(plus, x, sizeof(double) * nd);
    memcpy(minus, x, sizeof(double) * nd);
    plus[i] += eps;
    minus[i] -= eps;
    dx[i] = (f(s, plus, ctx) - f(s, minus, ctx)) / (2 * eps);
  }
}
/* finite_diff_hessian, finite_diff.h:67-86: upper triangle, mirrored; i == j perturbs twice */
static void fd_hessian(const orc_solver *s, scalar_fn f, const void *ctx, const double *x, int nd, double *out) {
  const double eps = s->d.params.fd_eps;
  double pp[NX], pm[NX], mp[NX], mm[NX];
  for (int i = 0; i < nd; i++) {
    for (int j = i; j < nd; j++) {
      memcpy(pp, x, sizeof(double) * nd);
      memcpy(pm, x, sizeof(double) * nd);
      memcpy(mp, x, sizeof(double) * nd);
      memcpy(mm, x, sizeof(double) * nd);
      pp[i] += eps; pp[j] += eps;
      pm[i] += eps; pm[j] -= eps;
      mp[i] -= eps; mp[j] += eps;
      mm[i] -= eps; mm[j] -= eps;
      const double v = (f(s, pp, ctx) - f(s, mp, ctx) - f(s, pm, ctx) + f(s, mm, ctx)) / (4 * eps * eps);
      out[i * nd + j] = out[j * nd + i] = v;
    }
  }
}
/* closures used by src/derivatives.cpp */
typedef struct { const double *other; } bind_ctx;
static double f_cost_x(const orc_solver *s, const double *v, const void *c) { return model_cost(s, v, ((const bind_ctx *)c)->other); }
static double f_cost_u(const orc_solver *s, const double *v, const void *c) { return model_cost(s, ((const bind_ctx *)c)->other, v); }
static double f_final(const orc_solver *s, const double *v, const void *c) { (void)c; return model_final_cost(s, v); }

/* finite_diff_jacobian (finite_diff.h:35-47) of integrate_dynamics wrt x (wrt_u = 0) or u (= 1),
 * as bound in get_dynamics_derivatives, src/derivatives.cpp:20-24.  out is n x nd row-major. */
static void fd_jacobian_dyn(const orc_solver *s, const double *x, const double *u, double dt, int wrt_u, double *out) {
  const double eps = s->d.params.fd_eps;
  const int n = s->n, nd = wrt_u ? s->m : s->n;
  double plus[NX], minus[NX], fp[NX], fm[NX];
  const double *base = wrt_u ? u : x;
  for (int i = 0; i < nd; i++) {
    memcpy(plus, base, sizeof(double) * nd);
    memcpy(minus, base, sizeof(double) * nd);
    plus[i] += eps;
    minus[i] -= eps;
    if (wrt_u) {
      model_integrate(s, x, plus, dt, fp);
      model_integrate(s, x, minus, dt, fm);
    } else {
      model_integrate(s, plus, u, dt, fp);
      model_integrate(s, minus, u, dt, fm);
    }
    for (int r = 0; r < n; r++) out[r * nd + i] = (fp[r] - fm[r]) / (2 * eps);
  }
}

/* get_dynamics_derivatives :15-26, get_cost_derivatives :29-54, get_cost_2nd_derivatives :57-144 */
static void compute_derivatives(orc_solver *s) {
  const int n = s->n, m = s->m, T = s->T;
  const double eps2 = s->d.params.fd_eps; /* src/derivatives.cpp:10 */
  double zero_u[NU] = {0};
  for (int t = 0; t < T; t++) { /* fx[T], fu[T] stay zero (src/ilqr_core.cpp:38-39) */
    if (s->d.flags & ILQR_FLAG_ANALYTIC_DYN) { /* opt-in: Jacobian of the Euler step x + dt f(x, u) in closed form */
      double A[NX * NX], Bm[NX * NU];
      if (s->d.model_id == ILQR_MODEL_ACROBOT) acrobot_dynamics_jac(s->xs + t * n, s->us + t * m, A, Bm);
      else di_dynamics_jac(A, Bm);
      for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) s->fx[(t * n + i) * n + j] = (i == j ? 1.0 : 0.0) + A[i * n + j] * s->d.dt;
        for (int j = 0; j < m; j++) s->fu[(t * n + i) * m + j] = Bm[i * m + j] * s->d.dt;
      }
      continue;
    }
    fd_jacobian_dyn(s, s->xs + t * n, s->us + t * m, s->d.dt, 0, s->fx + t * n * n);
    fd_jacobian_dyn(s, s->xs + t * n, s->us + t * m, s->d.dt, 1, s->fu + t * n * m);
  }
  for (int t = 0; t <= T; t++) {
    const double *xt = s->xs + t * n;
    const double *ut = t < T ? s->us + t * m : zero_u; /* derivatives.cpp:35-38,87,107,126 */
    double *cx = s->cx + t * n, *cu = s->cu + t * m, *cxx = s->cxx + t * n * n, *cxu = s->cxu + t * n * m,
           *cuu = s->cuu + t * m * m;
    if (s->d.cost_deriv == ILQR_COST_ANALYTIC) {
      model_cost_derivs(s, xt, ut, t == T, cx, cu, cxx, cxu, cuu);
      continue;
    }
    bind_ctx bu = {ut}, bx = {xt};
    if (t < T) {
      fd_gradient(s, f_cost_x, &bu, xt, n, cx); /* :45 */
      fd_gradient(s, f_cost_u, &bx, ut, m, cu); /* :46 */
      fd_hessian(s, f_cost_x, &bu, xt, n, cxx); /* :89-94 */
    } else {
      fd_gradient(s, f_final, NULL, xt, n, cx); /* :49 */
      memset(cu, 0, sizeof(double) * m);       /* :50-51 */
      fd_hessian(s, f_final, NULL, xt, n, cxx); /* :91-94 */
    }
    fd_hessian(s, f_cost_u, &bx, ut, m, cuu); /* :109-110 (t == T: u = 0; never read) */
    /* calculate_cxu :114-144, its own mixed stencil */
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < m; j++) {
        double px[NX], mx[NX], pu[NU], mu[NU];
        memcpy(px, xt, sizeof(double) * n);
        memcpy(mx, xt, sizeof(double) * n);
        memcpy(pu, ut, sizeof(double) * m);
        memcpy(mu, ut, sizeof(double) * m);
        px[i] += eps2; mx[i] -= eps2; pu[j] += eps2; mu[j] -= eps2;
        if (t < T)
          cxu[i * m + j] = (model_cost(s, px, pu) - model_cost(s, mx, pu) - model_cost(s, px, mu) + model_cost(s, mx, mu)) / (4 * (eps2 * eps2));
        else /* QUIRK :140 "this is wrong": algebraically zero, never read */
          cxu[i * m + j] = (model_final_cost(s, px) - model_final_cost(s, mx) - model_final_cost(s, px) + model_final_cost(s, mx)) / (4 * (eps2 * eps2));
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * boxQP: src/boxqp.cpp, include/boxqp.h
 * ---------------------------------------------------------------------------------------- */
static double clampd(double x, double lo, double hi) { /* clamp_to_limits boxqp.h:48-51: upper.cwiseMin(x.cwiseMax(lower)) */
  const double a = x < lo ? lo : x; /* Eigen max: (a < b) ? b : a */
  return hi < a ? hi : a;           /* Eigen min: (b < a) ? b : a */
}
static double quad_cost(int m, const double *Q, const double *c, const double *x) { /* boxqp.h:53-55 */
  double acc = 0; /* (0.5 x^T) Q x + x.c */
  double row[NU];
  for (int j = 0; j < m; j++) {
    double a = 0;
    for (int i = 0; i < m; i++) a += (0.5 * x[i]) * Q[i * m + j];
    row[j] = a;
  }
  for (int j = 0; j < m; j++) acc += row[j] * x[j];
  double d = 0;
  for (int j = 0; j < m; j++) d += x[j] * c[j];
  return acc + d;
}
/* Eigen::LLT on a dense r x r matrix, unblocked path (Eigen/src/Cholesky/LLT.h:302-325; size < 32).
 * Works in place on the lower triangle of A (row-major, ld r).  QUIRK: info() is never checked by
 * boxQP (src/boxqp.cpp:85-88); on a non-positive pivot Eigen returns early leaving that pivot
 * un-square-rooted and the trailing columns untouched — reproduced. */
static void eigen_llt_lower(int r, double *A) {
  for (int k = 0; k < r; k++) {
    double x = A[k * r + k];
    if (k > 0) {
      double sq = 0;
      for (int j = 0; j < k; j++) sq += A[k * r + j] * A[k * r + j];
      x -= sq;
    }
    if (x <= 0) return;
    A[k * r + k] = x = sqrt(x);
    for (int i = k + 1; i < r; i++) {
      if (k > 0) {
        double acc = 0;
        for (int j = 0; j < k; j++) acc += A[i * r + j] * A[k * r + j];
        A[i * r + k] -= acc;
      }
      A[i * r + k] /= x;
    }
  }
}
/* MatrixXd::inverse() for a dynamic matrix = PartialPivLU then solve(Identity)
 * (Eigen/src/LU/InverseImpl.h:23-28). */
static void lu_inverse(int r, const double *A, double *inv) {
  double lu[NU * NU];
  int perm[NU];
  memcpy(lu, A, sizeof(double) * r * r);
  for (int i = 0; i < r; i++) perm[i] = i;
  for (int k = 0; k < r; k++) {
    int piv = k;
    double best = fabs(lu[k * r + k]);
    for (int i = k + 1; i < r; i++)
      if (fabs(lu[i * r + k]) > best) { best = fabs(lu[i * r + k]); piv = i; }
    if (piv != k) {
      for (int j = 0; j < r; j++) { double t = lu[k * r + j]; lu[k * r + j] = lu[piv * r + j]; lu[piv * r + j] = t; }
      int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
    }
    for (int i = k + 1; i < r; i++) {
      lu[i * r + k] /= lu[k * r + k];
      for (int j = k + 1; j < r; j++) lu[i * r + j] -= lu[i * r + k] * lu[k * r + j];
    }
  }
  for (int c = 0; c < r; c++) {
    double y[NU];
    for (int i = 0; i < r; i++) y[i] = perm[i] == c ? 1.0 : 0.0;
    for (int i = 0; i < r; i++)
      for (int j = 0; j < i; j++) y[i] -= lu[i * r + j] * y[j];
    for (int i = r - 1; i >= 0; i--) {
      for (int j = i + 1; j < r; j++) y[i] -= lu[i * r + j] * y[j];
      y[i] /= lu[i * r + i];
    }
    for (int i = 0; i < r; i++) inv[i * r + c] = y[i];
  }
}
/* (R.inverse() * R.transpose().inverse()) — used at src/boxqp.cpp:105,110 and src/ilqr_core.cpp:379 */
static void rinv_rtinv(int r, const double *R, double *out) {
  double Rt[NU * NU], Ri[NU * NU], Rti[NU * NU];
  for (int i = 0; i < r; i++)
    for (int j = 0; j < r; j++) Rt[i * r + j] = R[j * r + i];
  lu_inverse(r, R, Ri);
  lu_inverse(r, Rt, Rti);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < r; j++) {
      double a = 0;
      for (int k = 0; k < r; k++) a += Ri[i * r + k] * Rti[k * r + j];
      out[i * r + j] = a;
    }
}

typedef struct { int failed, n_steps; double x_opt[NU]; double v_opt; } ls_result;

/* quadclamp_line_search, src/boxqp.cpp:143-178 */
static ls_result quadclamp_line_search(const ilqr_params *p, int m, const double *x0, const double *dir,
                                       const double *Q, const double *c, const double *lo, const double *hi) {
  ls_result res;
  memset(&res, 0, sizeof(res));
  double step = 1;
  double grad[NU];
  for (int i = 0; i < m; i++) {
    double a = 0;
    for (int j = 0; j < m; j++) a += Q[i * m + j] * x0[j];
    grad[i] = a + c[i];
  }
  double local_slope = 0;
  for (int i = 0; i < m; i++) local_slope += dir[i] * grad[i];
  if (local_slope >= 0) { /* :151 */
    res.failed = 1;
    return res; /* x_opt / v_opt left unset in the reference; zero here */
  }
  double xc[NU];
  for (int i = 0; i < m; i++) xc[i] = clampd(x0[i] + step * dir[i], lo[i], hi[i]);
  double v = quad_cost(m, Q, c, xc);
  const double old_v = quad_cost(m, Q, c, x0);
  while ((v - old_v) / (step * local_slope) < p->qp_armijo) { /* :161 */
    step *= p->qp_step_dec;
    res.n_steps++;
    for (int i = 0; i < m; i++) xc[i] = clampd(x0[i] + step * dir[i], lo[i], hi[i]);
    v = quad_cost(m, Q, c, xc);
    if (step < p->qp_min_step) { /* :169 */
      res.failed = 1;
      break;
    }
  }
  memcpy(res.x_opt, xc, sizeof(double) * m);
  res.v_opt = v;
  return res;
}

typedef struct { int result; double x_opt[NU]; int v_free[NU]; double R_free[NU * NU]; int r_dim; } qp_result;

/* boxQP, src/boxqp.cpp:26-139 */
static qp_result box_qp(const ilqr_params *p, int m, const double *Q, const double *c, const double *x0,
                        const double *lo, const double *hi) {
  qp_result res;
  memset(&res, 0, sizeof(res));
  res.r_dim = m; /* boxQPResult ctor sizes R_free m x m (boxqp.h:36-37); contents unspecified until factorised */
  double x[NU];
  for (int i = 0; i < m; i++) x[i] = clampd(x0[i], lo[i], hi[i]); /* :35 */
  /* QUIRK :36 — initial value has no 1/2: x^T Q x + x.c */
  double val;
  {
    double row[NU], acc = 0, d = 0;
    for (int j = 0; j < m; j++) {
      double a = 0;
      for (int i = 0; i < m; i++) a += x[i] * Q[i * m + j];
      row[j] = a;
    }
    for (int j = 0; j < m; j++) acc += row[j] * x[j];
    for (int j = 0; j < m; j++) d += x[j] * c[j];
    val = acc + d;
  }
  double oldvalue = 0;
  double clamped[NU], old_clamped[NU];
  for (int i = 0; i < m; i++) clamped[i] = 0; /* uninitialised in the reference; only read after iter 0 overwrote it */
  double grad[NU], grad_clamped[NU], search[NU];

  for (int iter = 0; iter <= p->qp_max_iter; iter++) { /* :50 */
    if (res.result != 0) break;
    if (iter > 0 && (oldvalue - val) < p->qp_min_rel_improve * fabs(oldvalue)) { /* :54 */
      res.result = 4;
      break;
    }
    for (int i = 0; i < m; i++) { /* :58 grad = Q x + c */
      double a = 0;
      for (int j = 0; j < m; j++) a += Q[i * m + j] * x[j];
      grad[i] = a + c[i];
    }
    oldvalue = val;

    memcpy(old_clamped, clamped, sizeof(clamped)); /* :62-71 */
    int all_clamped = 1, all_free = 1, nfree = 0;
    for (int i = 0; i < m; i++) {
      clamped[i] = 0;
      res.v_free[i] = 1;
      if ((fabs(x[i] - lo[i]) < p->qp_clamp_tol && grad[i] > 0) || (fabs(x[i] - hi[i]) < p->qp_clamp_tol && grad[i] < 0)) {
        clamped[i] = 1;
        res.v_free[i] = 0;
      }
      if (clamped[i] == 0) all_clamped = 0; else all_free = 0;
      nfree += res.v_free[i];
    }
    if (all_clamped) { /* :74 */
      res.result = 6;
      break;
    }
    /* QUIRK :80 — "changed" is detected by the SUM of flag differences, so a swap goes unnoticed */
    double dsum = 0;
    for (int i = 0; i < m; i++) dsum += old_clamped[i] - clamped[i];
    if (iter == 0 || dsum != 0) {
      double Qf[NU * NU];
      int idx[NU], r = 0;
      for (int i = 0; i < m; i++) if (res.v_free[i]) idx[r++] = i;
      for (int a = 0; a < r; a++)
        for (int b = 0; b < r; b++) Qf[a * r + b] = Q[idx[a] * m + idx[b]];
      eigen_llt_lower(r, Qf);
      /* R_free = matrixL().transpose(): upper triangular r x r */
      for (int a = 0; a < r; a++)
        for (int b = 0; b < r; b++) res.R_free[a * r + b] = (b >= a) ? Qf[b * r + a] : 0.0;
      res.r_dim = r;
    }
    double gn = 0; /* :93 */
    for (int i = 0; i < m; i++) if (res.v_free[i]) gn += grad[i] * grad[i];
    gn = sqrt(gn);
    if (gn < p->qp_min_grad) {
      res.result = 5;
      break;
    }
    for (int i = 0; i < m; i++) { /* :100 grad_clamped = Q (x .* clamped) + c */
      double a = 0;
      for (int j = 0; j < m; j++) a += Q[i * m + j] * (x[j] * clamped[j]);
      grad_clamped[i] = a + c[i];
    }
    { /* :103-119 */
      const int r = res.r_dim;
      double Hinv[NU * NU], gf[NU], xf[NU], sv[NU];
      rinv_rtinv(r, res.R_free, Hinv);
      int q = 0;
      for (int i = 0; i < m; i++) if (res.v_free[i]) { gf[q] = grad_clamped[i]; xf[q] = x[i]; q++; }
      for (int a = 0; a < r; a++) {
        double acc = 0;
        for (int b = 0; b < r; b++) acc += (-Hinv[a * r + b]) * gf[b];
        sv[a] = acc - xf[a];
      }
      (void)all_free;
      q = 0;
      for (int i = 0; i < m; i++) search[i] = res.v_free[i] ? sv[q++] : 0.0;
    }
    ls_result ls = quadclamp_line_search(p, m, x, search, Q, c, lo, hi); /* :121 */
    if (ls.failed) {
      res.result = 2;
      break;
    }
    memcpy(x, ls.x_opt, sizeof(double) * m); /* :133-134 */
    val = ls.v_opt;
  }
  memcpy(res.x_opt, x, sizeof(double) * m);
  return res;
}

/* ------------------------------------------------------------------------------------------
 * Solver: src/ilqr_core.cpp
 * ---------------------------------------------------------------------------------------- */

/* iLQR::forward_pass :305-337.  use_gains = (K.size() > 0) at :316.  Overwrites us in place (:323,
 * QUIRK: unclamped) and xs at the end (:334). */
static double forward_pass(orc_solver *s, const double *x0, const double *u, int use_gains) {
  const int n = s->n, m = s->m, T = s->T;
  double total = 0;
  double *x_new = (double *)malloc(sizeof(double) * (T + 1) * n);
  double xc[NX], uc[NU];
  memcpy(xc, x0, sizeof(double) * n);
  memcpy(x_new, x0, sizeof(double) * n);
  for (int t = 0; t < T; t++) {
    for (int j = 0; j < m; j++) uc[j] = u[t * m + j];
    if (use_gains) {
      for (int j = 0; j < m; j++) {
        double a = 0;
        for (int i = 0; i < n; i++) a += s->K[(t * m + j) * n + i] * (x_new[t * n + i] - s->xs[t * n + i]);
        uc[j] += a;
      }
    }
    if (s->d.flags & ILQR_FLAG_CLAMP_ROLLOUT) /* opt-in: "the right way" of :327-329, the applied control obeys the limits */
      for (int j = 0; j < m; j++) uc[j] = clampd(uc[j], s->umin[j], s->umax[j]);
    for (int j = 0; j < m; j++) s->us[t * m + j] = uc[j];
    total += model_cost(s, xc, uc);
    double x1[NX];
    model_integrate(s, xc, uc, s->d.dt, x1);
    memcpy(xc, x1, sizeof(double) * n);
    memcpy(x_new + (t + 1) * n, xc, sizeof(double) * n);
  }
  memcpy(s->xs, x_new, sizeof(double) * (T + 1) * n);
  free(x_new);
  total += model_final_cost(s, s->xs + T * n);
  return total;
}

/* iLQR::backward_pass :350-401 */
static int backward_pass(orc_solver *s) {
  const int n = s->n, m = s->m, T = s->T;
  const ilqr_params *p = &s->d.params;
  memcpy(s->Vx + T * n, s->cx + T * n, sizeof(double) * n);          /* :353 */
  memcpy(s->Vxx + T * n * n, s->cxx + T * n * n, sizeof(double) * n * n); /* :354 */
  s->dV[0] = s->dV[1] = 0;                                            /* :356 */
  for (int i = T - 1; i >= 0; i--) {
    const double *fx = s->fx + i * n * n, *fu = s->fu + i * n * m;
    const double *cx = s->cx + i * n, *cu = s->cu + i * m, *cxx = s->cxx + i * n * n, *cxu = s->cxu + i * n * m,
                 *cuu = s->cuu + i * m * m;
    const double *Vx1 = s->Vx + (i + 1) * n, *Vxx1 = s->Vxx + (i + 1) * n * n;
    double Qx[NX], Qu[NU], Qxx[NX * NX], Qux[NU * NX], Quu[NU * NU], QuuF[NU * NU];
    double fxtV[NX * NX], futV[NU * NX];
    for (int a = 0; a < n; a++) { /* :359 Qx = cx + fx^T Vx' */
      double acc = 0;
      for (int r = 0; r < n; r++) acc += fx[r * n + a] * Vx1[r];
      Qx[a] = cx[a] + acc;
    }
    for (int a = 0; a < m; a++) { /* :360 */
      double acc = 0;
      for (int r = 0; r < n; r++) acc += fu[r * m + a] * Vx1[r];
      Qu[a] = cu[a] + acc;
    }
    for (int a = 0; a < n; a++) /* fx^T Vxx' */
      for (int b = 0; b < n; b++) {
        double acc = 0;
        for (int r = 0; r < n; r++) acc += fx[r * n + a] * Vxx1[r * n + b];
        fxtV[a * n + b] = acc;
      }
    for (int a = 0; a < m; a++) /* fu^T Vxx' */
      for (int b = 0; b < n; b++) {
        double acc = 0;
        for (int r = 0; r < n; r++) acc += fu[r * m + a] * Vxx1[r * n + b];
        futV[a * n + b] = acc;
      }
    for (int a = 0; a < n; a++) /* :361 */
      for (int b = 0; b < n; b++) {
        double acc = 0;
        for (int r = 0; r < n; r++) acc += fxtV[a * n + r] * fx[r * n + b];
        Qxx[a * n + b] = cxx[a * n + b] + acc;
      }
    for (int a = 0; a < m; a++) /* :362 (and Qux_reg :366, the identical expression) */
      for (int b = 0; b < n; b++) {
        double acc = 0;
        for (int r = 0; r < n; r++) acc += futV[a * n + r] * fx[r * n + b];
        Qux[a * n + b] = cxu[b * m + a] + acc;
      }
    for (int a = 0; a < m; a++) /* :363 and :367 */
      for (int b = 0; b < m; b++) {
        double acc = 0;
        for (int r = 0; r < n; r++) acc += futV[a * n + r] * fu[r * m + b];
        Quu[a * m + b] = cuu[a * m + b] + acc;
        QuuF[a * m + b] = (cuu[a * m + b] + (a == b ? s->lam : 0.0)) + acc;
      }
    /* :369 — QUIRK: warm start k[min(i+1,T-1)]: for i = T-1 that is the previous pass's k[T-1] */
    const int iw = (i + 1 < T - 1) ? i + 1 : T - 1;
    double lo[NU], hi[NU];
    for (int j = 0; j < m; j++) {
      lo[j] = s->umin[j] - s->us[i * m + j];
      hi[j] = s->umax[j] - s->us[i * m + j];
    }
    qp_result res = box_qp(p, m, QuuF, Qu, s->k + iw * m, lo, hi);
    if (res.result < 1) return i; /* :371 — QUIRK: i == 0 is indistinguishable from success */

    double k_i[NU], K_i[NU * NX];
    memcpy(k_i, res.x_opt, sizeof(double) * m);
    memset(K_i, 0, sizeof(K_i));
    int any_free = 0;
    for (int j = 0; j < m; j++) any_free |= res.v_free[j];
    if (any_free) { /* :377-385 */
      const int r = res.r_dim;
      double Hinv[NU * NU];
      rinv_rtinv(r, res.R_free, Hinv);
      int rows[NU], q = 0;
      for (int j = 0; j < m; j++) if (res.v_free[j]) rows[q++] = j;
      for (int a = 0; a < r && a < q; a++)
        for (int b = 0; b < n; b++) {
          double acc = 0;
          for (int c = 0; c < r && c < q; c++) acc += (-Hinv[a * r + c]) * Qux[rows[c] * n + b];
          K_i[rows[a] * n + b] = acc;
        }
    }
    { /* :388-389 */
      double a0 = 0;
      for (int j = 0; j < m; j++) a0 += k_i[j] * Qu[j];
      s->dV[0] += a0;
      double row[NU], a1 = 0;
      for (int b = 0; b < m; b++) {
        double acc = 0;
        for (int a = 0; a < m; a++) acc += (0.5 * k_i[a]) * Quu[a * m + b];
        row[b] = acc;
      }
      for (int b = 0; b < m; b++) a1 += row[b] * k_i[b];
      s->dV[1] += a1;
    }
    double KtQuu[NX * NU]; /* K^T Quu  (n x m) */
    for (int a = 0; a < n; a++)
      for (int b = 0; b < m; b++) {
        double acc = 0;
        for (int c = 0; c < m; c++) acc += K_i[c * n + a] * Quu[c * m + b];
        KtQuu[a * m + b] = acc;
      }
    double *Vx = s->Vx + i * n, *Vxx = s->Vxx + i * n * n;
    for (int a = 0; a < n; a++) { /* :391 */
      double t1 = 0, t2 = 0, t3 = 0;
      for (int c = 0; c < m; c++) t1 += KtQuu[a * m + c] * k_i[c];
      for (int c = 0; c < m; c++) t2 += K_i[c * n + a] * Qu[c];
      for (int c = 0; c < m; c++) t3 += Qux[c * n + a] * k_i[c];
      Vx[a] = Qx[a] + t1 + t2 + t3;
    }
    double Vtmp[NX * NX];
    for (int a = 0; a < n; a++) /* :392 */
      for (int b = 0; b < n; b++) {
        double t1 = 0, t2 = 0, t3 = 0;
        for (int c = 0; c < m; c++) t1 += KtQuu[a * m + c] * K_i[c * n + b];
        for (int c = 0; c < m; c++) t2 += K_i[c * n + a] * Qux[c * n + b];
        for (int c = 0; c < m; c++) t3 += Qux[c * n + a] * K_i[c * n + b];
        Vtmp[a * n + b] = Qxx[a * n + b] + t1 + t2 + t3;
      }
    for (int a = 0; a < n; a++) /* :393 */
      for (int b = 0; b < n; b++) Vxx[a * n + b] = 0.5 * (Vtmp[a * n + b] + Vtmp[b * n + a]);
    memcpy(s->k + i * m, k_i, sizeof(double) * m);         /* :396 */
    memcpy(s->K + i * m * n, K_i, sizeof(double) * m * n); /* :397 */
  }
  return 0;
}

/* get_gradient_norm :405-412: mean_t max_j |k_tj| / (|u_tj| + 1) */
static double gradient_norm(const orc_solver *s) {
  double acc = 0.0;
  for (int t = 0; t < s->T; t++) {
    double mx = -INFINITY;
    for (int j = 0; j < s->m; j++) {
      const double v = fabs(s->k[t * s->m + j]) / (fabs(s->us[t * s->m + j]) + 1);
      if (v > mx) mx = v;
    }
    acc += mx;
  }
  return acc / s->T;
}

static int sgn(double v) { return (0 < v) - (v < 0); } /* include/common.h:43-44 */

/* the loop body of iLQR::generate_trajectory(), src/ilqr_core.cpp:103-288 */
static int iterate(orc_solver *s, int n_iters) {
  const ilqr_params *p = &s->d.params;
  const int n = s->n, m = s->m, T = s->T;
  double *x_old = (double *)malloc(sizeof(double) * (T + 1) * n);
  double *u_old = (double *)malloc(sizeof(double) * T * m);
  double *u_plus = (double *)malloc(sizeof(double) * T * m);
  int done_here = 0;
  for (; s->iter < p->max_iter && done_here < n_iters && s->status == ILQR_RUNNING; s->iter++) {
    done_here++;
    s->loop_trips++;
    memcpy(x_old, s->xs, sizeof(double) * (T + 1) * n); /* :104 */
    memcpy(u_old, s->us, sizeof(double) * T * m);
    if (s->flgChange) { /* :115-120 */
      compute_derivatives(s);
      s->flgChange = 0;
      s->n_deriv++;
    }
    int backPassDone = 0; /* :136-150 */
    while (!backPassDone) {
      s->diverge = backward_pass(s);
      s->n_backward++;
      if (s->diverge != 0) {
        s->dlam = fmax(s->dlam * p->lambda_factor, p->lambda_factor);
        s->lam = fmax(s->lam * s->dlam, p->lambda_min);
        if (s->lam > p->lambda_max) break;
        continue;
      }
      backPassDone = 1;
    }
    s->gnorm = gradient_norm(s); /* :153-159 */
    if (s->gnorm < p->tol_grad && s->lam < p->grad_lambda_gate) {
      s->status = ILQR_EXIT_GRAD;
      break;
    }
    int fwdPassDone = 0; /* :175-226 */
    double alpha = 0;
    s->alpha_index = -1;
    if (backPassDone) {
      for (int a = 0; a < p->n_alpha; a++) {
        alpha = p->alpha[a];
        for (int j = 0; j < T * m; j++) u_plus[j] = s->us[j] + s->k[j] * alpha; /* :188-190 */
        s->new_cost = forward_pass(s, s->x0, u_plus, 1);
        s->n_rollouts++;
        s->dcost = s->cost_s - s->new_cost;
        s->expected = -alpha * (s->dV[0] + alpha * s->dV[1]);
        double z;
        if (s->expected > 0) z = s->dcost / s->expected;
        else z = sgn(s->dcost); /* :206 */
        if (z > p->z_min) {
          fwdPassDone = 1;
          s->alpha_index = a;
          break;
        }
        memcpy(s->xs, x_old, sizeof(double) * (T + 1) * n); /* :218-219 */
        memcpy(s->us, u_old, sizeof(double) * T * m);
      }
      if (!fwdPassDone) alpha = 0.0;
    }
    s->alpha = alpha;
    if (fwdPassDone) { /* :242-263 */
      s->dlam = fmin(s->dlam / p->lambda_factor, 1 / p->lambda_factor);
      s->lam = s->lam * s->dlam * (s->lam > p->lambda_min); /* QUIRK :250: tests the OLD lambda; snaps to 0 */
      s->cost_s = s->new_cost;
      s->flgChange = 1;
      s->n_accept++;
      if (s->dcost < p->tol_fun) {
        s->status = ILQR_EXIT_TOLFUN;
        break;
      }
    } else { /* :264-282 */
      s->dlam = fmax(s->dlam * p->lambda_factor, p->lambda_factor);
      s->lam = fmax(s->lam * s->dlam, p->lambda_min);
      s->n_reject++;
      if (s->lam > p->lambda_max) {
        s->status = ILQR_EXIT_LAMBDA_MAX;
        break;
      }
    }
  }
  if (s->status == ILQR_RUNNING && s->iter >= p->max_iter) s->status = ILQR_EXIT_MAXITER;
  free(x_old);
  free(u_old);
  free(u_plus);
  return done_here;
}

/* ------------------------------------------------------------------------------------------
 * public probe API
 * ---------------------------------------------------------------------------------------- */
static void free_arrays(orc_solver *s) {
  free(s->x0); free(s->xs); free(s->us); free(s->fx); free(s->fu); free(s->cx); free(s->cu);
  free(s->cxx); free(s->cxu); free(s->cuu); free(s->Vx); free(s->Vxx); free(s->k); free(s->K);
  s->x0 = s->xs = s->us = s->fx = s->fu = s->cx = s->cu = s->cxx = s->cxu = s->cuu = s->Vx = s->Vxx = s->k = s->K = NULL;
}

orc_solver *orc_new(const ilqr_desc *desc) {
  orc_solver *s = (orc_solver *)calloc(1, sizeof(orc_solver));
  s->d = *desc;
  if (desc->model_id == ILQR_MODEL_ACROBOT) {
    s->n = 4; s->m = 1;
    s->umin[0] = -5; s->umax[0] = 5; /* acrobot.h:37 */
    s->goal[0] = 3.1415;              /* acrobot.h:20-21 */
    {                                 /* ... unless the caller hands over the goal of its own Acrobot object */
      int given = 0;
      for (int i = 0; i < 4; i++) given = given || desc->model_params[i] != 0.0;
      if (given)
        for (int i = 0; i < 4; i++) s->goal[i] = desc->model_params[i];
    }
  } else if (desc->model_id == ILQR_MODEL_DOUBLE_INTEGRATOR) {
    s->n = 4; s->m = 2;
    s->umin[0] = s->umin[1] = -0.5; s->umax[0] = s->umax[1] = 0.5; /* double_integrator.h:25-26 */
    for (int i = 0; i < 4; i++) s->goal[i] = desc->model_params[i];
  } else {
    free(s);
    return NULL;
  }
  if (desc->override_limits)
    for (int j = 0; j < s->m; j++) { s->umin[j] = desc->u_min[j]; s->umax[j] = desc->u_max[j]; }
  return s;
}
void orc_free(orc_solver *s) {
  if (!s) return;
  free_arrays(s);
  free(s);
}
void orc_dims(const orc_solver *s, int *n, int *m) { *n = s->n; *m = s->m; }

/* iLQR::init_traj :11-56 on a fresh object: open-loop rollout (K empty at :316), zero-fill */
double orc_init(orc_solver *s, const double *x0, const double *u0, int T) {
  const int n = s->n, m = s->m;
  free_arrays(s);
  s->T = T;
#define ZALLOC(cnt) ((double *)calloc((size_t)(cnt), sizeof(double)))
  s->x0 = ZALLOC(n); s->xs = ZALLOC((T + 1) * n); s->us = ZALLOC(T * m);
  s->fx = ZALLOC((T + 1) * n * n); s->fu = ZALLOC((T + 1) * n * m);
  s->cx = ZALLOC((T + 1) * n); s->cu = ZALLOC((T + 1) * m);
  s->cxx = ZALLOC((T + 1) * n * n); s->cxu = ZALLOC((T + 1) * n * m); s->cuu = ZALLOC((T + 1) * m * m);
  s->Vx = ZALLOC((T + 1) * n); s->Vxx = ZALLOC((T + 1) * n * n);
  s->k = ZALLOC(T * m); s->K = ZALLOC(T * m * n);
#undef ZALLOC
  memcpy(s->x0, x0, sizeof(double) * n);
  memcpy(s->xs, x0, sizeof(double) * n);
  double *u = (double *)malloc(sizeof(double) * T * m);
  memcpy(u, u0, sizeof(double) * T * m);
  s->cost_s = forward_pass(s, s->x0, u, 0);
  free(u);
  s->dV[0] = m; s->dV[1] = 1; /* :32 Vector2d(u_dims, 1) */
  s->lam = s->d.params.lambda_init;
  s->dlam = s->d.params.dlambda_init;
  s->flgChange = 1;
  s->iter = s->loop_trips = 0;
  s->status = ILQR_RUNNING;
  s->diverge = 0;
  s->alpha_index = -1;
  s->n_accept = s->n_reject = s->n_rollouts = s->n_backward = s->n_deriv = 0;
  return s->cost_s;
}

/* iLQR::generate_trajectory(x_0) :65-76 up to (not including) the loop: feedback rollout of the
 * kept us around the kept xs; lambda/dlambda carry over; loop-local state restarts (:87-92,103). */
double orc_warm_start(orc_solver *s, const double *x0) {
  memcpy(s->x0, x0, sizeof(double) * s->n);
  double *u = (double *)malloc(sizeof(double) * s->T * s->m);
  memcpy(u, s->us, sizeof(double) * s->T * s->m);
  s->cost_s = forward_pass(s, s->x0, u, 1);
  free(u);
  s->flgChange = 1;
  s->iter = 0;
  s->status = ILQR_RUNNING;
  return s->cost_s;
}

/* iLQR::generate_trajectory() called again on a finished solve (src/ilqr_core.cpp:78-102): the loop is re-entered
 * with iter = 0 and flgChange = true; lambda / dlambda (TU statics, include/ilqr.h:17-18) carry over. */
void orc_resume(orc_solver *s) {
  s->flgChange = 1;
  s->iter = 0;
  s->status = ILQR_RUNNING;
}

int orc_iterate(orc_solver *s, int n_iters) { return iterate(s, n_iters); }

int orc_backward_once(orc_solver *s, double lambda, int recompute) {
  if (recompute) compute_derivatives(s);
  s->lam = lambda;
  s->diverge = backward_pass(s);
  s->gnorm = gradient_norm(s);
  return s->diverge;
}
double orc_rollout_once(orc_solver *s, double alpha) {
  const int cnt = s->T * s->m;
  double *u_plus = (double *)malloc(sizeof(double) * cnt);
  for (int j = 0; j < cnt; j++) u_plus[j] = s->us[j] + s->k[j] * alpha;
  const double c = forward_pass(s, s->x0, u_plus, 1);
  free(u_plus);
  return c;
}

int orc_get(const orc_solver *s, int field, double *dst) {
  const int T = s->T, n = s->n, m = s->m;
  const double *src = NULL;
  int cnt = 0;
  switch (field) {
    case 0: src = s->xs; cnt = (T + 1) * n; break;
    case 1: src = s->us; cnt = T * m; break;
    case 2: src = s->K; cnt = T * m * n; break;
    case 3: src = s->k; cnt = T * m; break;
    case 4: dst[0] = s->cost_s; return 1;
    case 5: dst[0] = s->dV[0]; dst[1] = s->dV[1]; return 2;
    case 6: src = s->Vx; cnt = (T + 1) * n; break;
    case 7: src = s->Vxx; cnt = (T + 1) * n * n; break;
    case 8: src = s->fx; cnt = (T + 1) * n * n; break;
    case 9: src = s->fu; cnt = (T + 1) * n * m; break;
    case 10: src = s->cx; cnt = (T + 1) * n; break;
    case 11: src = s->cu; cnt = (T + 1) * m; break;
    case 12: src = s->cxx; cnt = (T + 1) * n * n; break;
    case 13: src = s->cxu; cnt = (T + 1) * n * m; break;
    case 14: src = s->cuu; cnt = (T + 1) * m * m; break;
    default: return -1;
  }
  memcpy(dst, src, sizeof(double) * cnt);
  return cnt;
}
double orc_scalar(const orc_solver *s, int which) {
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
long orc_int(const orc_solver *s, int which) {
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
    case 9: return s->T;
    case 10: return s->diverge;
    default: return -1;
  }
}

void orc_dynamics(const orc_solver *s, const double *x, const double *u, double *dx) { model_dynamics(s, x, u, dx); }
void orc_integrate(const orc_solver *s, const double *x, const double *u, double dt, double *x1) { model_integrate(s, x, u, dt, x1); }
double orc_cost(const orc_solver *s, const double *x, const double *u) { return model_cost(s, x, u); }
double orc_final_cost(const orc_solver *s, const double *x) { return model_final_cost(s, x); }

/* same `which` codes as ref_fd in oracle/ref_harness.cpp */
int orc_fd(const orc_solver *s, int which, const double *x, const double *u, double dt, double *out) {
  const int n = s->n, m = s->m;
  bind_ctx bu = {u}, bx = {x};
  switch (which) {
    case 0: fd_jacobian_dyn(s, x, u, dt, 0, out); return n * n;
    case 1: fd_jacobian_dyn(s, x, u, dt, 1, out); return n * m;
    case 2: fd_gradient(s, f_cost_x, &bu, x, n, out); return n;
    case 3: fd_gradient(s, f_cost_u, &bx, u, m, out); return m;
    case 4: fd_gradient(s, f_final, NULL, x, n, out); return n;
    case 5: fd_hessian(s, f_cost_x, &bu, x, n, out); return n * n;
    case 6: fd_hessian(s, f_cost_u, &bx, u, m, out); return m * m;
    case 7: fd_hessian(s, f_final, NULL, x, n, out); return n * n;
    default: return -1;
  }
}

int orc_boxqp(const ilqr_params *p, int m, const double *Q, const double *c, const double *x0, const double *lo,
              const double *hi, double *x_opt, int *v_free, double *R_free, int *r_dim) {
  ilqr_params dp;
  if (!p) { default_params(&dp); p = &dp; }
  qp_result r = box_qp(p, m, Q, c, x0, lo, hi);
  memcpy(x_opt, r.x_opt, sizeof(double) * m);
  memcpy(v_free, r.v_free, sizeof(int) * m);
  memcpy(R_free, r.R_free, sizeof(double) * r.r_dim * r.r_dim);
  *r_dim = r.r_dim;
  return r.result;
}
int orc_quadclamp(const ilqr_params *p, int m, const double *x0, const double *dir, const double *Q, const double *c,
                  const double *lo, const double *hi, double *x_opt, double *v_opt, int *n_steps) {
  ilqr_params dp;
  if (!p) { default_params(&dp); p = &dp; }
  ls_result r = quadclamp_line_search(p, m, x0, dir, Q, c, lo, hi);
  memcpy(x_opt, r.x_opt, sizeof(double) * m);
  *v_opt = r.v_opt;
  *n_steps = r.n_steps;
  return r.failed;
}
double orc_quadcost(int m, const double *Q, const double *c, const double *x) { return quad_cost(m, Q, c, x); }

long orc_solve_range(const ilqr_desc *desc, long b0, long b1, const double *x0, const double *u0, int max_trips,
                     double *cost, int *iters, int *status, long *n_accept, long *n_reject) {
  orc_solver *s = orc_new(desc);
  if (!s) return -1;
  const int T = desc->T;
  long total = 0;
  for (long b = b0; b < b1; b++) {
    orc_init(s, x0 + b * s->n, u0 + b * (long)T * s->m, T);
    iterate(s, max_trips < 0 ? desc->params.max_iter + 1 : max_trips);
    total += s->loop_trips;
    if (cost) cost[b - b0] = s->cost_s;
    if (iters) iters[b - b0] = s->loop_trips;
    if (status) status[b - b0] = s->status;
    if (n_accept) n_accept[b - b0] = s->n_accept;
    if (n_reject) n_reject[b - b0] = s->n_reject;
  }
  orc_free(s);
  return total;
}
