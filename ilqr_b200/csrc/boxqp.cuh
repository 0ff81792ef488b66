/*
 * boxqp.cuh — the box-constrained QP of the backward pass, one problem per call.
 *
 * Follows boxQP (src/boxqp.cpp:26-139) and quadclamp_line_search (:143-178) with the helpers of
 * include/boxqp.h (clamp_to_limits :48-51, quadCost :53-55, approx_eq :61-64) branch for branch:
 * the result codes, the un-halved initial value (:36), the 1e-4 clamp test, the
 * sum-of-flag-differences refactorisation test (:80) and the unchecked Cholesky (:85-88, Eigen
 * LLT.h:302-325 leaves a non-positive pivot un-square-rooted and stops) are all kept, because
 * the iLQR iterates depend on them.  `R^-1 R^-T` is formed the way the reference forms it: two
 * dense inverses by partially pivoted LU (Eigen/src/LU/InverseImpl.h:23-28) and a product.
 *
 * m > 1: one lane of the warp runs a problem; Q, c and the work arrays live in the warp's
 * shared-memory scratch, so the dynamic indexing of the free-set gathers costs nothing.
 * m == 1 (acrobot): a scalar version with the same arithmetic, entirely in registers.
 */
#ifndef ILQR_BOXQP_CUH_
#define ILQR_BOXQP_CUH_

#include "models.cuh"

namespace ilqr {

template <typename S>
struct QPParams {
  int max_iter;
  S min_grad, min_rel_improve, step_dec, min_step, armijo, clamp_tol;
};

/* problem + result + work space of one boxQP call */
template <int M, typename S>
struct QPWork {
  /* in */
  S Q[M * M], c[M], x0[M], lo[M], hi[M];
  /* out */
  S x[M];
  S R[M * M]; /* R_free, r_dim x r_dim row-major, upper triangular */
  S Hinv[M * M]; /* R^-1 R^-T of the final factor (what the gain computation needs, ilqr_core.cpp:379) */
  int v_free[M];
  int r_dim;
  int result;
  /* work */
  S grad[M], grad_clamped[M], search[M], clamped[M], old_clamped[M];
  S Qf[M * M], lu[M * M], Ri[M * M], Rti[M * M], Rt[M * M], tmp[M], xc[M], row[M], gf[M], xf[M], sv[M];
  int idx[M], perm[M];
};

template <typename S>
ILQR_HD S clampd(S x, S lo, S hi) { /* upper.cwiseMin(x.cwiseMax(lower)) */
  const S a = x < lo ? lo : x;
  return hi < a ? hi : a;
}

/* quadCost: (0.5 x^T) Q x + x.c */
template <int M, typename S>
ILQR_HD S quad_cost(const S *Q, const S *c, const S *x, S *row) {
  S acc = 0;
  for (int j = 0; j < M; j++) {
    S a = 0;
    for (int i = 0; i < M; i++) a += (S(0.5) * x[i]) * Q[i * M + j];
    row[j] = a;
  }
  for (int j = 0; j < M; j++) acc += row[j] * x[j];
  S d = 0;
  for (int j = 0; j < M; j++) d += x[j] * c[j];
  return acc + d;
}

/* Eigen::LLT unblocked, in place on the lower triangle of the r x r matrix A */
template <typename S>
ILQR_HD void llt_lower(int r, S *A) {
  for (int k = 0; k < r; k++) {
    S x = A[k * r + k];
    if (k > 0) {
      S sq = 0;
      for (int j = 0; j < k; j++) sq += A[k * r + j] * A[k * r + j];
      x -= sq;
    }
    if (x <= 0) return;
    A[k * r + k] = x = t_sqrt(x);
    for (int i = k + 1; i < r; i++) {
      if (k > 0) {
        S acc = 0;
        for (int j = 0; j < k; j++) acc += A[i * r + j] * A[k * r + j];
        A[i * r + k] -= acc;
      }
      A[i * r + k] /= x;
    }
  }
}

/* dense inverse by partially pivoted LU + solve against the identity */
template <typename S>
ILQR_HD void lu_inverse(int r, const S *A, S *inv, S *lu, int *perm, S *y) {
  for (int i = 0; i < r * r; i++) lu[i] = A[i];
  for (int i = 0; i < r; i++) perm[i] = i;
  for (int k = 0; k < r; k++) {
    int piv = k;
    S best = t_abs(lu[k * r + k]);
    for (int i = k + 1; i < r; i++)
      if (t_abs(lu[i * r + k]) > best) {
        best = t_abs(lu[i * r + k]);
        piv = i;
      }
    if (piv != k) {
      for (int j = 0; j < r; j++) {
        const S t = lu[k * r + j];
        lu[k * r + j] = lu[piv * r + j];
        lu[piv * r + j] = t;
      }
      const int t = perm[k];
      perm[k] = perm[piv];
      perm[piv] = t;
    }
    for (int i = k + 1; i < r; i++) {
      lu[i * r + k] /= lu[k * r + k];
      for (int j = k + 1; j < r; j++) lu[i * r + j] -= lu[i * r + k] * lu[k * r + j];
    }
  }
  for (int c = 0; c < r; c++) {
    for (int i = 0; i < r; i++) y[i] = perm[i] == c ? S(1) : S(0);
    for (int i = 0; i < r; i++)
      for (int j = 0; j < i; j++) y[i] -= lu[i * r + j] * y[j];
    for (int i = r - 1; i >= 0; i--) {
      for (int j = i + 1; j < r; j++) y[i] -= lu[i * r + j] * y[j];
      y[i] /= lu[i * r + i];
    }
    for (int i = 0; i < r; i++) inv[i * r + c] = y[i];
  }
}

/* R.inverse() * R.transpose().inverse()   (src/boxqp.cpp:105,110; src/ilqr_core.cpp:379) */
template <int M, typename S>
ILQR_HD void rinv_rtinv(QPWork<M, S> &w, int r, S *out) {
  for (int i = 0; i < r; i++)
    for (int j = 0; j < r; j++) w.Rt[i * r + j] = w.R[j * r + i];
  lu_inverse(r, w.R, w.Ri, w.lu, w.perm, w.tmp);
  lu_inverse(r, w.Rt, w.Rti, w.lu, w.perm, w.tmp);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < r; j++) {
      S a = 0;
      for (int k = 0; k < r; k++) a += w.Ri[i * r + k] * w.Rti[k * r + j];
      out[i * r + j] = a;
    }
}

/* quadclamp_line_search; returns failed, writes x_opt into w.xc and the value into *v_opt */
template <int M, typename S>
ILQR_HD bool quadclamp(const QPParams<S> &p, QPWork<M, S> &w, const S *x0, const S *dir, S *v_opt) {
  S step = 1;
  S slope = 0;
  for (int i = 0; i < M; i++) {
    S a = 0;
    for (int j = 0; j < M; j++) a += w.Q[i * M + j] * x0[j];
    w.tmp[i] = a + w.c[i];
  }
  for (int i = 0; i < M; i++) slope += dir[i] * w.tmp[i];
  if (slope >= 0) return true; /* :151 */
  for (int i = 0; i < M; i++) w.xc[i] = clampd(x0[i] + step * dir[i], w.lo[i], w.hi[i]);
  S v = quad_cost<M>(w.Q, w.c, w.xc, w.row);
  const S old_v = quad_cost<M>(w.Q, w.c, x0, w.row);
  bool failed = false;
  while ((v - old_v) / (step * slope) < p.armijo) { /* :161 */
    step *= p.step_dec;
    for (int i = 0; i < M; i++) w.xc[i] = clampd(x0[i] + step * dir[i], w.lo[i], w.hi[i]);
    v = quad_cost<M>(w.Q, w.c, w.xc, w.row);
    if (step < p.min_step) { /* :169 */
      failed = true;
      break;
    }
  }
  *v_opt = v;
  return failed;
}

/* generic m: the problem is in w (Q, c, x0, lo, hi); the result is left in w */
template <int M, typename S>
ILQR_HD void box_qp_generic(const QPParams<S> &p, QPWork<M, S> &w) {
  w.result = 0;
  w.r_dim = M;
  for (int i = 0; i < M * M; i++) w.R[i] = 0;
  for (int i = 0; i < M; i++) w.x[i] = clampd(w.x0[i], w.lo[i], w.hi[i]); /* :35 */
  S val; /* :36 — x^T Q x + x.c, no 1/2 */
  {
    S acc = 0, d = 0;
    for (int j = 0; j < M; j++) {
      S a = 0;
      for (int i = 0; i < M; i++) a += w.x[i] * w.Q[i * M + j];
      w.row[j] = a;
    }
    for (int j = 0; j < M; j++) acc += w.row[j] * w.x[j];
    for (int j = 0; j < M; j++) d += w.x[j] * w.c[j];
    val = acc + d;
  }
  S oldvalue = 0;
  for (int i = 0; i < M; i++) w.clamped[i] = 0;

  for (int iter = 0; iter <= p.max_iter; iter++) { /* :50 */
    if (iter > 0 && (oldvalue - val) < p.min_rel_improve * t_abs(oldvalue)) { /* :54 */
      w.result = 4;
      break;
    }
    for (int i = 0; i < M; i++) { /* :58 */
      S a = 0;
      for (int j = 0; j < M; j++) a += w.Q[i * M + j] * w.x[j];
      w.grad[i] = a + w.c[i];
    }
    oldvalue = val;
    bool all_clamped = true; /* :62-71 */
    for (int i = 0; i < M; i++) {
      w.old_clamped[i] = w.clamped[i];
      w.clamped[i] = 0;
      w.v_free[i] = 1;
      if ((t_abs(w.x[i] - w.lo[i]) < p.clamp_tol && w.grad[i] > 0) ||
          (t_abs(w.x[i] - w.hi[i]) < p.clamp_tol && w.grad[i] < 0)) {
        w.clamped[i] = 1;
        w.v_free[i] = 0;
      }
      if (w.clamped[i] == 0) all_clamped = false;
    }
    if (all_clamped) { /* :74 */
      w.result = 6;
      break;
    }
    S dsum = 0; /* :80 */
    for (int i = 0; i < M; i++) dsum += w.old_clamped[i] - w.clamped[i];
    if (iter == 0 || dsum != 0) {
      int r = 0;
      for (int i = 0; i < M; i++)
        if (w.v_free[i]) w.idx[r++] = i;
      for (int a = 0; a < r; a++)
        for (int b = 0; b < r; b++) w.Qf[a * r + b] = w.Q[w.idx[a] * M + w.idx[b]];
      llt_lower(r, w.Qf);
      for (int a = 0; a < r; a++)
        for (int b = 0; b < r; b++) w.R[a * r + b] = (b >= a) ? w.Qf[b * r + a] : S(0);
      w.r_dim = r;
    }
    S gn = 0; /* :93 */
    for (int i = 0; i < M; i++)
      if (w.v_free[i]) gn += w.grad[i] * w.grad[i];
    gn = t_sqrt(gn);
    if (gn < p.min_grad) {
      w.result = 5;
      break;
    }
    for (int i = 0; i < M; i++) { /* :100 */
      S a = 0;
      for (int j = 0; j < M; j++) a += w.Q[i * M + j] * (w.x[j] * w.clamped[j]);
      w.grad_clamped[i] = a + w.c[i];
    }
    { /* :103-119 */
      const int r = w.r_dim;
      rinv_rtinv(w, r, w.Hinv);
      int q = 0;
      for (int i = 0; i < M; i++)
        if (w.v_free[i]) {
          w.gf[q] = w.grad_clamped[i];
          w.xf[q] = w.x[i];
          q++;
        }
      for (int a = 0; a < r; a++) {
        S acc = 0;
        for (int b = 0; b < r; b++) acc += (-w.Hinv[a * r + b]) * w.gf[b];
        w.sv[a] = acc - w.xf[a];
      }
      q = 0;
      for (int i = 0; i < M; i++) w.search[i] = w.v_free[i] ? w.sv[q++] : S(0);
    }
    S v_opt;
    if (quadclamp<M>(p, w, w.x, w.search, &v_opt)) { /* :121-126 */
      w.result = 2;
      break;
    }
    for (int i = 0; i < M; i++) w.x[i] = w.xc[i]; /* :133-134 */
    val = v_opt;
  }
  /* the gain computation re-forms R^-1 R^-T from the returned factor (ilqr_core.cpp:378-379) */
  bool any_free = false;
  for (int i = 0; i < M; i++) any_free = any_free || (w.v_free[i] != 0);
  if (w.result >= 1 && any_free) rinv_rtinv(w, w.r_dim, w.Hinv);
}

/* m == 1: the same arithmetic on scalars held in registers; nothing touches memory, so every lane
 * of the warp can run it redundantly (ilqr_core.cuh, backward_step).  Values the general code
 * recomputes are formed once because they are bit-identical by construction: R is only ever
 * factorised at iteration 0 when there is a single variable (the flag difference of :80 is 0 - 0
 * afterwards: a clamped variable leaves the loop at once), 1/R is the same division for R^-1 and
 * R^-T, sqrt(grad^2) is |grad| (exact in IEEE-754 away from over/underflow, where both sides of
 * the `< minGrad` test agree anyway), and the R^-1 R^-T of the gain computation
 * (ilqr_core.cpp:379) is the one of the Newton steps.  sqrt(Q) -> 1/R -> 1/R^2 is the longest
 * dependency chain of a backward timestep, so it is issued first, unconditionally (the values
 * are only READ where the reference computes them), and overlaps the clamp / gradient tests.
 *
 * The Armijo test of quadclamp_line_search (:161), (v - old_v) / (step * slope) < armijo, costs a
 * double-precision division per trial.  With den = step * slope < 0 it is decided exactly by
 * comparing num with armijo * den whenever the two differ by more than a few dozen ulp (then the
 * correctly rounded quotient lies on the same side of `armijo` as the real one); only in the
 * remaining sliver, or when den is not a normal finite number, is the division carried out. */
template <typename S>
struct QPScalar {
  S x, R, Hinv;
  int v_free, result;
};

#if defined(ILQR_QP_STATS) && !defined(__CUDA_ARCH__)
#include <stdio.h>
#include <stdlib.h>
static long g_qp_calls, g_qp_fast[8], g_qp_slow_arm[3], g_qp_bt, g_qp_c0;
static void qp_stats_print() {
  fprintf(stderr, "qp calls %ld fast results [2]=%ld [4]=%ld [5]=%ld [6]=%ld; slow: armijo-fail %ld ambiguous %ld third-iteration %ld; backtracks %ld clamped-at-once %ld\n",
          g_qp_calls, g_qp_fast[2], g_qp_fast[4], g_qp_fast[5], g_qp_fast[6], g_qp_slow_arm[1], g_qp_slow_arm[2], g_qp_slow_arm[0], g_qp_bt, g_qp_c0);
}
static void qp_stats(int result, int arm) {
  if (g_qp_calls++ == 0) atexit(qp_stats_print);
  if (arm == 99) g_qp_c0++;
  if (result >= 0) g_qp_fast[result & 7]++;
  else g_qp_slow_arm[arm < 0 ? 2 : arm]++;
}
#endif

/* The Armijo test with den < 0: 1 = fails (backtrack), 0 = passes, -1 = too close to call without the division */
template <typename S>
ILQR_HD int armijo_quick(S num, S den, S armijo) {
  const S p = armijo * den;
  const S ap = t_abs(p);
  const S diff = num - p;
  const S tol = S(64) * (sizeof(S) == 8 ? S(2.220446049250313e-16) : S(1.1920929e-07));
  const bool definite = ap > S(1e-30) && ap < S(1e30) && t_abs(diff) > tol * ap;
  return definite ? (num > p ? 1 : 0) : -1; /* den < 0 flips the inequality */
}
template <typename S>
ILQR_HD bool armijo_fails(S num, S den, S armijo) {
  const int q = armijo_quick(num, den, armijo);
  return q >= 0 ? q != 0 : num / den < armijo;
}

template <typename S>
ILQR_HD bool qp_is_clamped(const QPParams<S> &p, S x, S grad, S lo, S hi) { /* :62-71 */
  return (t_abs(x - lo) < p.clamp_tol && grad > 0) || (t_abs(x - hi) < p.clamp_tol && grad < 0);
}

/* the loop of boxQP as written, from the start (every exit the reference has); rare, so out of line on the device */
template <typename S>
#if defined(__CUDACC__) && !defined(ILQR_QP_LOOP_INLINE)
__host__ __device__ __noinline__
#else
ILQR_HD
#endif
QPScalar<S> box_qp_scalar_loop(const QPParams<S> &p, S Q, S c, S x0, S lo, S hi, S R, S Hinv) {
  S x = clampd(x0, lo, hi);
  S val = (x * Q) * x + x * c; /* :36, no 1/2 */
  S oldvalue = 0;
  int result = 0, vfree = 1;
  bool factorised = false;
  for (int iter = 0; iter <= p.max_iter; iter++) {
    if (iter > 0 && (oldvalue - val) < p.min_rel_improve * t_abs(oldvalue)) {
      result = 4;
      break;
    }
    const S grad = Q * x + c;
    oldvalue = val;
    vfree = 1;
    if (qp_is_clamped(p, x, grad, lo, hi)) {
      vfree = 0;
      result = 6;
      break;
    }
    factorised = true; /* :80-90 */
    if (t_abs(grad) < p.min_grad) { /* sqrt(grad * grad) */
      result = 5;
      break;
    }
    const S grad_clamped = Q * (x * S(0)) + c;
    const S search = (-Hinv) * grad_clamped - x;
    /* quadclamp_line_search */
    const S slope = search * grad;
    if (slope >= 0) {
      result = 2;
      break;
    }
    S step = 1;
    S xc = clampd(x + step * search, lo, hi);
    S v = ((S(0.5) * xc) * Q) * xc + xc * c;
    const S old_v = ((S(0.5) * x) * Q) * x + x * c;
    bool failed = false;
    while (armijo_fails(v - old_v, step * slope, p.armijo)) {
#if defined(ILQR_QP_STATS) && !defined(__CUDA_ARCH__)
      g_qp_bt++;
#endif
      step *= p.step_dec;
      xc = clampd(x + step * search, lo, hi);
      v = ((S(0.5) * xc) * Q) * xc + xc * c;
      if (step < p.min_step) {
        failed = true;
        break;
      }
    }
    if (failed) {
      result = 2;
      break;
    }
    x = xc;
    val = v;
  }
  QPScalar<S> r;
  r.x = x;
  r.R = factorised ? R : S(0);
  r.Hinv = Hinv;
  r.v_free = vfree;
  r.result = result;
  return r;
}

/* Nearly every call ends in its first or second Newton iteration (clamped at once; or one step, then
 * "gradient small" / "all clamped" / "no improvement").  Those two iterations are written out as straight-line
 * code — every quantity either of them may need, then the reference's tests in the reference's order on the
 * finished values — so that the independent dependency chains (sqrt -> 1/R -> 1/R^2; x, value, gradient and
 * the clamp tests; the trial point with its value; the second gradient) overlap in the instruction stream
 * instead of queueing behind the loop's branches.  One call in ten has to back-track in the first line search
 * (a step cut short by a bound, measured on the BASELINE batch): that loop is the reference's, then the
 * second iteration's tests are redone on its result.  Anything else (a third iteration) runs the reference's
 * loop from the start.  Same operations on the same operands whichever way. */
template <typename S>
ILQR_HD QPScalar<S> box_qp_scalar(const QPParams<S> &p, S Q, S c, S x0, S lo, S hi) {
  /* iteration 0 */
  const S x = clampd(x0, lo, hi);
  const S grad0 = Q * x + c;
  const bool clamped0 = qp_is_clamped(p, x, grad0, lo, hi);
  if (clamped0) { /* "all clamped" at once (:74-77): 43 % of the calls of a solve — the warm start is the previous
                     timestep's k, which sat on the bound too — and none of what follows is needed (the gains of a
                     clamped control are zero, the factor is never formed) */
#if defined(ILQR_QP_STATS) && !defined(__CUDA_ARCH__)
    qp_stats(6, 99);
#endif
    QPScalar<S> r;
    r.x = x;
    r.R = S(0);
    r.Hinv = S(0);
    r.v_free = 0;
    r.result = 6;
    return r;
  }
  const S val0 = (x * Q) * x + x * c; /* :36, no 1/2 */
  const bool small0 = t_abs(grad0) < p.min_grad;
  const S gc0 = Q * (x * S(0)) + c;
  const S R = (Q <= 0) ? Q : t_sqrt(Q); /* Eigen LLT leaves a non-positive pivot as it is */
  const S Ri = S(1) / R;
  const S Hinv = Ri * Ri;
  const S search = (-Hinv) * gc0 - x;
  const S slope = search * grad0;
  const bool ascent = slope >= 0;
  S xc = clampd(x + search, lo, hi); /* step = 1 */
  S v = ((S(0.5) * xc) * Q) * xc + xc * c;
  const S old_v = ((S(0.5) * x) * Q) * x + x * c;
  const int arm = armijo_quick(v - old_v, slope, p.armijo);
  /* iteration 1, at (xc, v) */
  bool noimp1 = (val0 - v) < p.min_rel_improve * t_abs(val0);
  S grad1 = Q * xc + c;
  bool clamped1 = qp_is_clamped(p, xc, grad1, lo, hi);
  bool small1 = t_abs(grad1) < p.min_grad;
  /* the reference's order of tests */
  const bool stop0 = clamped0 || small0 || ascent;
  bool ls_failed = false;
  if (!stop0 && arm != 0) { /* quadclamp_line_search's loop (:161-173) */
    S step = 1;
    bool fails = arm > 0 ? true : (v - old_v) / slope < p.armijo;
    while (fails) {
      step *= p.step_dec;
      xc = clampd(x + step * search, lo, hi);
      v = ((S(0.5) * xc) * Q) * xc + xc * c;
      if (step < p.min_step) {
        ls_failed = true;
        break;
      }
      fails = armijo_fails(v - old_v, step * slope, p.armijo);
    }
    noimp1 = (val0 - v) < p.min_rel_improve * t_abs(val0);
    grad1 = Q * xc + c;
    clamped1 = qp_is_clamped(p, xc, grad1, lo, hi);
    small1 = t_abs(grad1) < p.min_grad;
  }
  const int res0 = clamped0 ? 6 : (small0 ? 5 : 2);
  const int res1 = ls_failed ? 2 : (noimp1 ? 4 : (clamped1 ? 6 : (small1 ? 5 : -1)));
  const int result = stop0 ? res0 : (p.max_iter >= 1 ? res1 : -1);
#if defined(ILQR_QP_STATS) && !defined(__CUDA_ARCH__)
  qp_stats(result, stop0 ? 0 : arm);
#endif
  if (__builtin_expect(result < 0, 0)) return box_qp_scalar_loop<S>(p, Q, c, x0, lo, hi, R, Hinv);
  const bool keep_x = stop0 || ls_failed; /* a failed line search leaves x where it was (:123-126) */
  QPScalar<S> r;
  r.x = keep_x ? x : xc;
  r.R = clamped0 ? S(0) : R;
  r.Hinv = Hinv;
  r.v_free = keep_x ? (clamped0 ? 0 : 1) : ((!noimp1 && clamped1) ? 0 : 1);
  r.result = result;
  return r;
}

/* the problem in w solved by the version the backward pass uses for this m (test hook) */
template <int M, typename S>
ILQR_HD void box_qp(const QPParams<S> &p, QPWork<M, S> &w) {
  if constexpr (M == 1) {
    const QPScalar<S> r = box_qp_scalar<S>(p, w.Q[0], w.c[0], w.x0[0], w.lo[0], w.hi[0]);
    w.x[0] = r.x;
    w.R[0] = r.R;
    w.Hinv[0] = r.Hinv;
    w.v_free[0] = r.v_free;
    w.r_dim = 1;
    w.result = r.result;
  } else {
    box_qp_generic<M, S>(p, w);
  }
}

}  // namespace ilqr
#endif
