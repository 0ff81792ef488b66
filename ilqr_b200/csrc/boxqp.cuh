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
 * One lane of the warp runs a problem; Q, c and the work arrays live in the warp's shared-memory
 * scratch, so the dynamic indexing of the free-set gathers costs nothing.  m == 1 (acrobot) has a
 * scalar fast path with the same arithmetic.
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

/* m == 1: the same arithmetic on scalars held in registers.  Three values the general code
 * recomputes are reused here because they are bit-identical by construction: 1/R is formed once
 * for R^-1 and R^-T (the same division), sqrt(grad^2) is |grad| (exact in IEEE-754 away from
 * over/underflow, where both sides of the `< minGrad` test agree anyway), and the R^-1 R^-T the
 * gain computation needs (ilqr_core.cpp:379) is the one of the last Newton step because R is only
 * ever factorised once when there is a single variable. */
template <typename S>
ILQR_HD void box_qp_scalar(const QPParams<S> &p, QPWork<1, S> &w) {
  const S Q = w.Q[0], c = w.c[0], lo = w.lo[0], hi = w.hi[0];
  S x = clampd(w.x0[0], lo, hi);
  S val = (x * Q) * x + x * c;
  S oldvalue = 0;
  S R = 0, Hinv = 0;
  bool have_hinv = false;
  int result = 0, vfree = 1;
  for (int iter = 0; iter <= p.max_iter; iter++) {
    if (iter > 0 && (oldvalue - val) < p.min_rel_improve * t_abs(oldvalue)) {
      result = 4;
      break;
    }
    const S grad = Q * x + c;
    oldvalue = val;
    vfree = 1;
    if ((t_abs(x - lo) < p.clamp_tol && grad > 0) || (t_abs(x - hi) < p.clamp_tol && grad < 0)) {
      vfree = 0;
      result = 6;
      break;
    }
    /* the flag difference of :80 is 0 - 0 here (a clamped variable has just left the loop): factorise at iter 0 only */
    if (iter == 0) R = (Q <= 0) ? Q : t_sqrt(Q);
    if (t_abs(grad) < p.min_grad) { /* sqrt(grad * grad) */
      result = 5;
      break;
    }
    const S grad_clamped = Q * (x * S(0)) + c;
    if (!have_hinv) {
      const S Ri = S(1) / R;
      Hinv = Ri * Ri;
      have_hinv = true;
    }
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
    while ((v - old_v) / (step * slope) < p.armijo) {
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
  w.x[0] = x;
  w.v_free[0] = vfree;
  w.R[0] = R;
  w.r_dim = 1;
  w.result = result;
  if (result >= 1 && vfree) {
    if (!have_hinv) {
      const S Ri = S(1) / R;
      Hinv = Ri * Ri;
    }
    w.Hinv[0] = Hinv;
  }
}

template <int M, typename S>
ILQR_HD void box_qp(const QPParams<S> &p, QPWork<M, S> &w) {
  if constexpr (M == 1) box_qp_scalar<S>(p, w);
  else box_qp_generic<M, S>(p, w);
}

}  // namespace ilqr
#endif
