/*
 * models.cuh — device twins of the reference's Model subclasses.
 *
 * The reference's plugin surface is a host vtable over dynamic Eigen vectors
 * (include/model.h:6-21: dynamics / cost / final_cost, Euler integrate_dynamics :12-15).  Those
 * headers cannot pass through nvcc (vendored Eigen 3.3.4 vs CUDA 12.9), so every model that runs
 * on the GPU has a hand-written twin here: a struct with compile-time N (x_dims) and M (u_dims)
 * and static functions over plain arrays, templated on the scalar type.
 *
 * Expressions are written in the order the reference evaluates them so that, built without FMA
 * contraction, the f64 results are bit-identical to the reference model's up to the libm/libdevice
 * difference in sin/cos.  Model parameters (`mp`) come from ilqr_desc::model_params.
 */
#ifndef ILQR_MODELS_CUH_
#define ILQR_MODELS_CUH_

#include <math.h>

#include "trig.cuh"

#if defined(__CUDACC__)
#define ILQR_HDC __host__ __device__ constexpr
#else
#define ILQR_HDC constexpr
#endif

namespace ilqr {

ILQR_HD double t_sqrt(double v) { return ::sqrt(v); }
ILQR_HD double t_abs(double v) { return ::fabs(v); }
ILQR_HD float t_sqrt(float v) { return ::sqrtf(v); }
ILQR_HD float t_abs(float v) { return ::fabsf(v); }

/* ------------------------------------------------------------------------------------------
 * Acrobot — include/acrobot.h.  n = 4 (q1, q2, q1dot, q2dot), m = 1 (elbow torque) :27-28.
 * mp[0..3] = goal, set to (3.1415, 0, 0, 0) by the library (the literal of :20-21, not pi).
 * ---------------------------------------------------------------------------------------- */
struct Acrobot {
  static constexpr int N = 4;
  static constexpr int M = 1;
  /* Acrobot::dynamics :72-81 with H :43-51, C :53-61, G :63-70; parameters :19,23-25
   * (I1 = I2 = l1 = l2 = m1 = m2 = 1, lc = 0.5, g = 9.81).  The 2x2 inverse is the closed form
   * Eigen uses for fixed-size 2x2 (Eigen/src/LU/InverseImpl.h:76-94).  The three sincos
   * (of q2, q1, q1 + q2; trig.cuh) are independent, which the instruction scheduler exploits. */
  /* The part of the dynamics that depends on the configuration q = (x[0], x[1]) alone: the trigonometry, the
   * inverse of H and G.  The finite-difference sweep perturbs one variable at a time, and for the velocities
   * and the control this part is the same for every perturbed point of a timestep, so it is formed once
   * (Core::derivative_sweep); dynamics() below is configure() followed by dynamics_cfg(), statement for
   * statement what it was as one function. */
  static constexpr unsigned kConfigVars = 0x3; /* bit i set: the Config depends on x[i] */
  template <typename S>
  struct Config {
    S s2, Hi00, Hi01, Hi10, Hi11, G0, G1;
  };
  template <typename S>
  ILQR_HD static void configure(const S *x, const S * /*mp*/, Config<S> &cf) {
    const S I1 = 1, I2 = 1, l1 = 1, m1 = 1, m2 = 1, g = S(9.81);
    const S l2 = 1;
    const S lc1 = S(0.5) * l1, lc2 = S(0.5) * l2;
    const S q0 = x[0], q1 = x[1];
    S sn[3], cs[3];
    sincos_det3(q1, q0, q0 + q1, sn, cs);
    const S c2 = cs[0], s1 = sn[1], s1p2 = sn[2];
    cf.s2 = sn[0];
    const S H00 = I1 + I2 + m2 * l1 * l1 + 2 * m2 * l1 * lc2 * c2;
    const S H01 = I2 + m2 * l1 * lc2 * c2;
    const S H10 = I2 + m2 * l1 * lc2 * c2;
    const S H11 = I2;
    cf.G0 = m1 * g * lc1 * s1 + m2 * g * (l1 * s1 + lc2 * s1p2);
    cf.G1 = m2 * g * lc2 * s1p2;
    const S det = H00 * H11 - H10 * H01;
    const S invdet = S(1) / det;
    cf.Hi00 = H11 * invdet;
    cf.Hi10 = -H10 * invdet;
    cf.Hi01 = -H01 * invdet;
    cf.Hi11 = H00 * invdet;
  }
  template <typename S>
  ILQR_HD static void dynamics_cfg(const Config<S> &cf, const S *x, const S *u, const S * /*mp*/, S *dx) {
    const S l1 = 1, l2 = 1, m2 = 1;
    const S lc2 = S(0.5) * l2;
    const S qd0 = x[2], qd1 = x[3];
    const S s2 = cf.s2;
    const S C00 = -2 * m2 * l1 * lc2 * s2 * qd1;
    const S C01 = -m2 * l2 * lc2 * s2 * qd1;
    const S C10 = m2 * l1 * lc2 * s2 * qd0;
    const S C11 = 0;
    const S r0 = (S(0) - (C00 * qd0 + C01 * qd1)) - cf.G0; /* Vector2d(0,u) - C*qdot - G */
    const S r1 = (u[0] - (C10 * qd0 + C11 * qd1)) - cf.G1;
    dx[0] = qd0;
    dx[1] = qd1;
    dx[2] = cf.Hi00 * r0 + cf.Hi01 * r1;
    dx[3] = cf.Hi10 * r0 + cf.Hi11 * r1;
  }
  template <typename S>
  ILQR_HD static void dynamics(const S *x, const S *u, const S *mp, S *dx) {
    Config<S> cf;
    configure(x, mp, cf);
    dynamics_cfg(cf, x, u, mp, dx);
  }
  /* OPT-IN (ILQR_FLAG_ANALYTIC_DYN; not in the reference, which only has finite differences and lists analytic
   * Jacobians as future work, notes.md:15,45): closed-form Jacobian of dynamics(), A[i * N + j] = d dx_i / d x_j,
   * Bm[i * M + j] = d dx_i / d u_j.  With qdd = H^-1 r:  d qdd / d z = H^-1 (d r / d z - (d H / d z) qdd).  Same
   * expressions in the same order as oracle/ilqr_oracle.c: acrobot_dynamics_jac. */
  template <typename S>
  ILQR_HD static void dynamics_jac(const S *x, const S *u, const S *mp, S *A, S *Bm) {
    const S I1 = 1, I2 = 1, l1 = 1, l2 = 1, m1 = 1, m2 = 1, g = S(9.81);
    const S lc1 = S(0.5) * l1, lc2 = S(0.5) * l2;
    const S q0 = x[0], q1 = x[1], qd0 = x[2], qd1 = x[3];
    S sn[3], cs[3];
    sincos_det3(q1, q0, q0 + q1, sn, cs);
    const S c2 = cs[0], s2 = sn[0], c1 = cs[1], c12 = cs[2];
    const S a = m2 * l1 * lc2, b = m2 * l2 * lc2;
    const S H00 = I1 + I2 + m2 * l1 * l1 + 2 * a * c2, H01 = I2 + a * c2, H11 = I2;
    const S det = H00 * H11 - H01 * H01;
    const S invdet = S(1) / det;
    const S Hi00 = H11 * invdet, Hi01 = -H01 * invdet, Hi11 = H00 * invdet;
    S dx[N];
    dynamics(x, u, mp, dx);
    const S qdd0 = dx[2], qdd1 = dx[3];
    S dr0[4], dr1[4]; /* d r / d z - (d H / d z) qdd for z = q0, q1, qd0, qd1 */
    dr0[0] = -(m1 * g * lc1 * c1 + m2 * g * (l1 * c1 + lc2 * c12));
    dr1[0] = -(m2 * g * lc2 * c12);
    dr0[1] = ((2 * a * qd0 * qd1 + b * qd1 * qd1) * c2 - m2 * g * lc2 * c12) + s2 * (2 * a * qdd0 + a * qdd1);
    dr1[1] = (-(a * c2 * qd0 * qd0) - m2 * g * lc2 * c12) + s2 * (a * qdd0);
    dr0[2] = 2 * a * s2 * qd1;
    dr1[2] = -(2 * a * s2 * qd0);
    dr0[3] = 2 * a * s2 * qd0 + 2 * b * s2 * qd1;
    dr1[3] = S(0);
#pragma unroll
    for (int i = 0; i < 16; i++) A[i] = S(0);
    A[0 * 4 + 2] = S(1);
    A[1 * 4 + 3] = S(1);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      A[2 * 4 + j] = Hi00 * dr0[j] + Hi01 * dr1[j];
      A[3 * 4 + j] = Hi01 * dr0[j] + Hi11 * dr1[j];
    }
    Bm[0] = S(0);
    Bm[1] = S(0);
    Bm[2] = Hi01;
    Bm[3] = Hi11;
  }
  /* Acrobot::cost :83-92 — Ks = Kd = 0, Kr = 0.1: the reference still evaluates 0 * (e0^2 + e1^2) + 0 * (e2^2 + e3^2) +
   * Kr^2 u^2.  While every |x_i| <= 1e150 the two sums of squares are finite and non-negative, so the state terms
   * are exactly +0 and (+0 + +0) + r == r bit for bit: the 14 operations behind them are skipped (3 % of the
   * instructions of a solve).  Beyond that bound the squares may overflow and 0 * inf = NaN is what the reference
   * returns, so the full expression is evaluated. */
  template <typename S>
  ILQR_HD static S cost(const S *x, const S *u, const S *mp) {
    const S Ks = 0, Kd = 0, Kr = S(0.1);
    const S lim = sizeof(S) == 8 ? S(1e150) : S(1e18);
    const S run = Kr * Kr * (u[0] * u[0]);
    if (t_abs(x[0]) <= lim && t_abs(x[1]) <= lim && t_abs(x[2]) <= lim && t_abs(x[3]) <= lim) return run;
    const S e0 = mp[0] - x[0], e1 = mp[1] - x[1], e2 = mp[2] - x[2], e3 = mp[3] - x[3];
    return Ks * Ks * (e0 * e0 + e1 * e1) + Kd * Kd * (e2 * e2 + e3 * e3) + run;
  }
  /* Acrobot::final_cost :94-100 — Ks = Kd = 20 */
  template <typename S>
  ILQR_HD static S final_cost(const S *x, const S *mp) {
    const S e0 = mp[0] - x[0], e1 = mp[1] - x[1], e2 = mp[2] - x[2], e3 = mp[3] - x[3];
    const S Ks = 20, Kd = 20;
    return Ks * Ks * (e0 * e0 + e1 * e1) + Kd * Kd * (e2 * e2 + e3 * e3);
  }
  /* Closed-form cost derivatives (cost_deriv == ILQR_COST_ANALYTIC; the reference only has the
   * finite-difference path), one entry at a time over the stacked variable v = (x, u):
   * cost_d1(c) = d cost / d v_c, cost_d2(c, d) = d2 cost / d v_c d v_d.  The backward pass asks
   * for exactly the entry a lane needs, so nothing is staged through memory. */
  template <typename S>
  ILQR_HD static S cost_d1(int c, const S *x, const S *u, const S *mp, bool terminal) {
    if (terminal) return c < N ? S(-800.0) * (mp[c] - x[c]) : S(0);
    const S w = S(0.1) * S(0.1);
    const S cu = 2 * w * u[0]; /* unconditional: a select, not a branch around the load */
    return c == N ? cu : S(0);
  }
  template <typename S>
  ILQR_HD static S cost_d2(int c, int d, const S * /*x*/, const S * /*u*/, const S * /*mp*/, bool terminal) {
    const S diag = c == N ? 2 * (S(0.1) * S(0.1)) : (terminal ? S(800.0) : S(0));
    return c == d ? diag : S(0);
  }
};

/* ------------------------------------------------------------------------------------------
 * DoubleIntegrator — include/double_integrator.h.  n = 4 (x, y, vx, vy), m = 2 (Fx, Fy) :16-17.
 * mp[0..3] = goal (constructor argument :14).  Hx = diag(1, 1, .2, .2), Hu = I, mass 1 :19-24,51.
 * ---------------------------------------------------------------------------------------- */
struct DoubleIntegrator {
  static constexpr int N = 4;
  static constexpr int M = 2;
  template <typename S>
  ILQR_HD static void dynamics(const S *x, const S *u, const S * /*mp*/, S *dx) { /* :29-37 */
    const S mass = 1;
    dx[0] = x[2];
    dx[1] = x[3];
    dx[2] = u[0] / mass;
    dx[3] = u[1] / mass;
  }
  template <typename S>
  ILQR_HD static void dynamics_jac(const S *, const S *, const S *, S *A, S *Bm) { /* opt-in, see Acrobot::dynamics_jac */
    const S mass = 1;
#pragma unroll
    for (int i = 0; i < 16; i++) A[i] = S(0);
#pragma unroll
    for (int i = 0; i < 8; i++) Bm[i] = S(0);
    A[0 * 4 + 2] = S(1);
    A[1 * 4 + 3] = S(1);
    Bm[2 * 2 + 0] = S(1) / mass;
    Bm[3 * 2 + 1] = S(1) / mass;
  }
  static constexpr unsigned kConfigVars = 0; /* nothing to share between perturbed points */
  template <typename S>
  struct Config {};
  template <typename S>
  ILQR_HD static void configure(const S *, const S *, Config<S> &) {}
  template <typename S>
  ILQR_HD static void dynamics_cfg(const Config<S> &, const S *x, const S *u, const S *mp, S *dx) { dynamics(x, u, mp, dx); }
  template <typename S>
  ILQR_HD static S quad(const S *x, const S *mp, S scale) {
    const S hx[4] = {S(1), S(1), S(0.2), S(0.2)};
    S acc = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const S e = mp[i] - x[i];
      acc += (e * (scale * hx[i])) * e;
    }
    return acc;
  }
  template <typename S>
  ILQR_HD static S cost(const S *x, const S *u, const S *mp) { /* :39-43 */
    return quad(x, mp, S(1)) + (u[0] * u[0] + u[1] * u[1]);
  }
  template <typename S>
  ILQR_HD static S final_cost(const S *x, const S *mp) { /* :45-48 */
    return quad(x, mp, S(10));
  }
  template <typename S>
  ILQR_HD static S cost_d1(int c, const S *x, const S *u, const S *mp, bool terminal) {
    const S sc = terminal ? S(10) : S(1);
    if (c < N) {
      const S h = c < 2 ? S(1) : S(0.2);
      return S(-2.0) * (sc * h) * (mp[c] - x[c]);
    }
    return terminal ? S(0) : 2 * u[c - N];
  }
  template <typename S>
  ILQR_HD static S cost_d2(int c, int d, const S * /*x*/, const S * /*u*/, const S * /*mp*/, bool terminal) {
    if (c != d) return S(0);
    if (c >= N) return S(2);
    const S sc = terminal ? S(10) : S(1);
    const S h = c < 2 ? S(1) : S(0.2);
    return S(2.0) * (sc * h);
  }
};

/* Model::integrate_dynamics, include/model.h:12-15: x + dynamics(x, u) * dt */
#if defined(__CUDACC__) && defined(ILQR_NOINLINE_DYN)
#define ILQR_HD_DYN __host__ __device__ __noinline__
#else
#define ILQR_HD_DYN ILQR_HD
#endif
template <class Model, typename S>
ILQR_HD_DYN void integrate(const S *x, const S *u, const S *mp, S dt, S *x1) {
  S dx[Model::N];
  Model::dynamics(x, u, mp, dx);
#pragma unroll
  for (int i = 0; i < Model::N; i++) x1[i] = x[i] + dx[i] * dt;
}
/* the same with the configuration-dependent part already formed */
template <class Model, typename S>
ILQR_HD void integrate_cfg(const typename Model::template Config<S> &cf, const S *x, const S *u, const S *mp, S dt, S *x1) {
  S dx[Model::N];
  Model::dynamics_cfg(cf, x, u, mp, dx);
#pragma unroll
  for (int i = 0; i < Model::N; i++) x1[i] = x[i] + dx[i] * dt;
}
}  // namespace ilqr
#endif
