/*
 * ilqr_core.cuh — one iLQR problem instance, solved by ONE WARP.
 *
 * This is the whole hot path of the reference for one trajectory, re-designed for a 32-lane warp
 * that owns the trajectory from the first rollout to termination.  One loop trip
 * (src/ilqr_core.cpp:103-288) is
 *
 *   1. derivative sweep — only when the trajectory changed (flgChange, :115-120).  The central
 *      differences of the Euler step (src/derivatives.cpp:15-26, finite_diff.h:35-47) have no
 *      dependence between timesteps, so they are NOT done inside the serial backward recursion:
 *      the T*(n+m) (timestep, variable) pairs are spread over the 32 lanes, each lane evaluating
 *      its +eps / -eps pair and writing one Jacobian column.  With finite-difference cost
 *      derivatives (src/derivatives.cpp:29-144) the T*20 stencil outputs are spread the same way.
 *      The columns go to a per-warp buffer that stays in L2 (acrobot, T = 200: 32 KB).
 *   2. backward pass (:350-401), serial in t, five warp phases per timestep: F^T Vxx' entries one
 *      per lane; the (n+m)^2 Q-function entries one per lane (:359-367); boxQP + gains on one lane
 *      (src/boxqp.cpp, :369-389); the n + n^2 entries of Vx, Vxx one per lane (:391-392);
 *      symmetrisation (:393) and the k / K / gradient-norm-term stores.  Vx, Vxx and every
 *      intermediate live in the warp's shared-memory scratch.
 *   3. line search (:184-226): the n_alpha candidate rollouts (:305-337) run concurrently, one per
 *      lane, each streaming its states/controls to a per-warp candidate buffer; the first accepted
 *      index is taken — result-identical to the reference's serial early-exit loop, because every
 *      candidate starts from the same (xs, us, K, k) — and the accepted candidate is committed by
 *      a coalesced copy (no re-roll).
 *   4. lambda schedule / termination (:136-159, :242-282) on per-trajectory scalars; lambda and
 *      dlambda are per trajectory here (process-wide statics in the reference, ilqr.h:17-18).
 *
 * Per-trajectory arrays are contiguous in t and move in tiles by coalesced warp-wide copies.
 *
 * Lanes communicate only through the scratch (never shuffles), in phases separated by a warp
 * barrier.  That lets tests/emu/ compile this same header with g++ and run the phases lane by
 * lane on the CPU (`HostExec`), so control flow and arithmetic order are checked bit for bit
 * against the oracle without a GPU.  The emulator is test infrastructure; the product only ever
 * instantiates `WarpExec`.
 *
 * Arithmetic order follows the reference statement by statement (sequential accumulations from
 * zero, no FMA contraction: the .cu is built with -fmad=false), so f64 results differ from the
 * reference only through sin/cos (trig.cuh).
 */
#ifndef ILQR_CORE_CUH_
#define ILQR_CORE_CUH_

#include "boxqp.cuh"
#include "models.cuh"

namespace ilqr {

constexpr int kTile = 32;      /* timesteps per staged tile in the rollouts */
constexpr int kTileB = 8;      /* timesteps per staged tile in the backward pass */
constexpr int kMaxAlpha = 16;  /* ILQR_MAX_ALPHA */

/* exit reasons / ops: numeric values equal the ILQR_* macros of include/ilqr_b200.h */
enum { kRunning = 0, kExitGrad = 1, kExitTolFun = 2, kExitLambdaMax = 3, kExitMaxIter = 4 };
enum { kCostFD = 0, kCostAnalytic = 1 };
enum { kRollOpen = 0, kRollWarm = 1, kRollClosed = 2 };

/* every constant of the solve, in the scalar type the path computes in */
template <typename S>
struct SolveParams {
  int T, max_iter, n_alpha;
  S dt;
  S tol_fun, tol_grad, lambda_init, dlambda_init, lambda_factor, lambda_max, lambda_min, z_min, grad_lambda_gate;
  S alpha[kMaxAlpha];
  S fd_eps;
  S u_min[4], u_max[4];
  S mp[16];
  QPParams<S> qp;
};

/* per-trajectory solver state that survives between launches (one slot per trajectory) */
template <typename S>
struct TrajState {
  S cost, lam, dlam, dV0, dV1, gnorm, dcost, expected, alpha, new_cost;
  int iter, trips, status, alpha_index, n_accept, n_reject, n_backward, diverge, flg_change, n_rollouts, n_deriv;
};

/* base pointers of one trajectory's arrays (contiguous in t) */
template <typename S>
struct TrajPtrs {
  const S *x0; /* [n]           */
  S *xs;       /* [T+1][n]      */
  S *us;       /* [T][m]        */
  S *K;        /* [T][m][n]     */
  S *k;        /* [T][m]        */
  S *Vx0;      /* [n]   Vx[0] of the last backward pass  */
  S *Vxx0;     /* [n][n]                                  */
  TrajState<S> *st;
};

/* per-WARP work buffers in global memory (one set per resident warp, reused across trajectories) */
template <typename S>
struct SlotPtrs {
  S *F;      /* [T][n+m][n]   Jacobian columns of the Euler step: F[t][j][r] = d x'_r / d (x|u)_j   */
  S *C;      /* [T][NC]       finite-difference cost derivatives (cx cu cxx cxu cuu), FD mode only  */
  S *cand_x; /* [n_alpha][T][n]  candidate states x_1..x_T of the line search                       */
  S *cand_u; /* [n_alpha][T][m]  candidate controls                                                  */
  S *gterm;  /* [T] per-timestep terms of the gradient norm (shared memory on the device)           */
};

/* the warp's working set */
template <int N, int M, typename S, int CD>
struct Scratch {
  static constexpr int NM = N + M;
  static constexpr int NC = N + M + N * N + N * M + M * M;
  /* staged tiles: rollouts use kTile timesteps, the backward pass the first kTileB of the same arrays */
  S xs[kTile * N], us[kTile * M], K[kTile * M * N], k[kTile * M];
  S Ft[kTileB * NM * N];                     /* backward: Jacobian columns of the tile      */
  S Ct[CD == kCostFD ? kTileB * NC : 1];     /* backward: FD cost derivatives of the tile   */
  /* one timestep */
  S x[N], u[M];
  S cx[N], cu[M], cxx[N * N], cxu[N * M], cuu[M * M]; /* terminal / analytic cost derivatives */
  S Vx[N], Vxx[N * N];  /* value function at i+1, overwritten with i at the end of the step */
  S W[NM * N];          /* F^T Vxx' */
  S Qx[N], Qu[M], Qxx[N * N], Qux[M * N], Quu[M * M];
  S Kc[M * N], kc[M], kprev[M];
  S Vtmp[N * N], Vxn[N];
  S newcost[kMaxAlpha];
  QPWork<M, S> qp;
  TrajState<S> st;
  int flag;
};

template <int N, int M, typename S>
struct LaneRegs {
  S x[N];
  S uc[M];
  S cost;
};

#if defined(__CUDACC__)
/* device: the calling thread is one lane; a phase ends with a warp barrier */
template <int N, int M, typename S>
struct WarpExec {
  LaneRegs<N, M, S> regs;
  int lane;
  template <class Fn>
  __device__ __forceinline__ void lanes(Fn fn) {
    fn(lane, regs);
    __syncwarp();
  }
};
#endif
/* host (tests only): run the phase for lane 0..31 in turn */
template <int N, int M, typename S>
struct HostExec {
  LaneRegs<N, M, S> regs[32];
  template <class Fn>
  void lanes(Fn fn) {
    for (int l = 0; l < 32; l++) fn(l, regs[l]);
  }
};

/* read that must see what another lane of this warp stored to global memory before the last barrier */
template <typename S>
ILQR_HD S ld_fresh(const S *p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(p);
#else
  return *p;
#endif
}

template <class Model, typename S, int CD, class Exec>
struct Core {
  static constexpr int N = Model::N, M = Model::M, NM = N + M;
  using Sc = Scratch<N, M, S, CD>;
  using Lane = LaneRegs<N, M, S>;
  static constexpr int NC = Sc::NC;

  const SolveParams<S> &P;
  Sc &sc;
  Exec &ex;
  TrajPtrs<S> tr;
  SlotPtrs<S> sl;

  ILQR_HD Core(const SolveParams<S> &p, Sc &s, Exec &e, const TrajPtrs<S> &t, const SlotPtrs<S> &w)
      : P(p), sc(s), ex(e), tr(t), sl(w) {}

  /* ---- finite differences ------------------------------------------------------------------ */

  /* perturbed copy: v[q] (+ da if q == a) (+ db if q == b), in that order */
  template <int D>
  ILQR_HD static void perturb(const S *v, int a, S da, int b, S db, S *out) {
#pragma unroll
    for (int q = 0; q < D; q++) {
      S t = v[q];
      if (q == a) t += da;
      if (q == b) t += db;
      out[q] = t;
    }
  }

  /* One finite-difference cost-stencil output (src/derivatives.cpp:29-144) at (x, u).  Output ids:
   * [0,N) cx, [N,N+M) cu, then cxx upper triangle (i <= j, row by row), cuu upper triangle,
   * then cxu (i, j).  terminal: cx and cxx of final_cost only. */
  static constexpr int kNxx = N * (N + 1) / 2, kNuu = M * (M + 1) / 2;
  static constexpr int kStencilStep = N + M + kNxx + kNuu + N * M;
  static constexpr int kStencilTerm = N + kNxx;

  ILQR_HD static void tri_index(int o, int D, int &i, int &j) {
    i = 0;
    int rowlen = D;
    while (o >= rowlen) {
      o -= rowlen;
      rowlen--;
      i++;
    }
    j = i + o;
  }

  ILQR_HD void cost_stencil(int o, bool terminal, const S *x, const S *u, S *cx, S *cu, S *cxx, S *cxu, S *cuu) {
    const S eps = P.fd_eps;
    const S *mp = P.mp;
    S xa[N], ua[M];
    auto fx = [&](const S *xv, const S *uv) -> S { return terminal ? Model::final_cost(xv, mp) : Model::cost(xv, uv, mp); };
    if (o < N) { /* finite_diff_gradient wrt x, finite_diff.h:22-33 */
      perturb<N>(x, o, eps, -1, S(0), xa);
      const S p = fx(xa, u);
      perturb<N>(x, o, -eps, -1, S(0), xa);
      const S m = fx(xa, u);
      cx[o] = (p - m) / (2 * eps);
      return;
    }
    o -= N;
    if (!terminal) {
      if (o < M) {
        perturb<M>(u, o, eps, -1, S(0), ua);
        const S p = Model::cost(x, ua, mp);
        perturb<M>(u, o, -eps, -1, S(0), ua);
        const S m = Model::cost(x, ua, mp);
        cu[o] = (p - m) / (2 * eps);
        return;
      }
      o -= M;
    }
    if (o < kNxx) { /* finite_diff_hessian wrt x, finite_diff.h:67-86 */
      int i, j;
      tri_index(o, N, i, j);
      perturb<N>(x, i, eps, j, eps, xa);
      const S pp = fx(xa, u);
      perturb<N>(x, i, -eps, j, eps, xa);
      const S mpv = fx(xa, u);
      perturb<N>(x, i, eps, j, -eps, xa);
      const S pm = fx(xa, u);
      perturb<N>(x, i, -eps, j, -eps, xa);
      const S mm = fx(xa, u);
      const S v = (pp - mpv - pm + mm) / (4 * eps * eps);
      cxx[i * N + j] = v;
      cxx[j * N + i] = v;
      return;
    }
    o -= kNxx;
    if (terminal) return;
    if (o < kNuu) {
      int i, j;
      tri_index(o, M, i, j);
      perturb<M>(u, i, eps, j, eps, ua);
      const S pp = Model::cost(x, ua, mp);
      perturb<M>(u, i, -eps, j, eps, ua);
      const S mpv = Model::cost(x, ua, mp);
      perturb<M>(u, i, eps, j, -eps, ua);
      const S pm = Model::cost(x, ua, mp);
      perturb<M>(u, i, -eps, j, -eps, ua);
      const S mm = Model::cost(x, ua, mp);
      const S v = (pp - mpv - pm + mm) / (4 * eps * eps);
      cuu[i * M + j] = v;
      cuu[j * M + i] = v;
      return;
    }
    o -= kNuu;
    if (o < N * M) { /* calculate_cxu, src/derivatives.cpp:114-144 (its own stencil) */
      const int i = o / M, j = o % M;
      S xp[N], xm[N], up[M], um[M];
      perturb<N>(x, i, eps, -1, S(0), xp);
      perturb<N>(x, i, -eps, -1, S(0), xm);
      perturb<M>(u, j, eps, -1, S(0), up);
      perturb<M>(u, j, -eps, -1, S(0), um);
      cxu[i * M + j] =
          (Model::cost(xp, up, mp) - Model::cost(xm, up, mp) - Model::cost(xp, um, mp) + Model::cost(xm, um, mp)) /
          (4 * (eps * eps));
    }
  }

  /* get_dynamics_derivatives (+ get_cost_derivatives / get_cost_2nd_derivatives in FD mode) for the
   * whole horizon, parallel over (timestep, variable): lane <- task, each task the +eps / -eps pair
   * of finite_diff_jacobian (finite_diff.h:35-47) for one column. */
  ILQR_HD void derivative_sweep() {
    const int T = P.T;
    const int n_dyn = T * NM;
    for (int base = 0; base < n_dyn; base += 32) {
      ex.lanes([&](int lane, Lane &) {
        const int task = base + lane;
        if (task >= n_dyn) return;
        const int t = task / NM, j = task - t * NM;
        S x[N], u[M], xa[N], ua[M], fp[N], fm[N];
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = tr.xs[t * N + i];
#pragma unroll
        for (int i = 0; i < M; i++) u[i] = tr.us[t * M + i];
        perturb<N>(x, j, P.fd_eps, -1, S(0), xa);
        perturb<M>(u, j - N, P.fd_eps, -1, S(0), ua);
        integrate<Model, S>(xa, ua, P.mp, P.dt, fp);
        perturb<N>(x, j, -P.fd_eps, -1, S(0), xa);
        perturb<M>(u, j - N, -P.fd_eps, -1, S(0), ua);
        integrate<Model, S>(xa, ua, P.mp, P.dt, fm);
#pragma unroll
        for (int r = 0; r < N; r++) sl.F[(size_t)task * N + r] = (fp[r] - fm[r]) / (2 * P.fd_eps);
      });
    }
    if constexpr (CD == kCostFD) {
      const int n_c = T * kStencilStep;
      for (int base = 0; base < n_c; base += 32) {
        ex.lanes([&](int lane, Lane &) {
          const int task = base + lane;
          if (task >= n_c) return;
          const int t = task / kStencilStep, o = task - t * kStencilStep;
          S x[N], u[M];
#pragma unroll
          for (int i = 0; i < N; i++) x[i] = tr.xs[t * N + i];
#pragma unroll
          for (int i = 0; i < M; i++) u[i] = tr.us[t * M + i];
          S *c = sl.C + (size_t)t * NC;
          cost_stencil(o, false, x, u, c, c + N, c + N + M, c + N + M + N * N, c + N + M + N * N + N * M);
        });
      }
    }
  }

  /* Vx[T] = cx[T], Vxx[T] = cxx[T]  (src/ilqr_core.cpp:353-354) from sc.x = xs[T] */
  ILQR_HD void phase_terminal() {
    ex.lanes([&](int lane, Lane &) {
      if (CD == kCostFD) {
        for (int o = lane; o < kStencilTerm; o += 32) cost_stencil(o, true, sc.x, sc.u, sc.cx, sc.cu, sc.cxx, sc.cxu, sc.cuu);
      } else if (lane == 0) {
        S cu[M], cxu[N * M], cuu[M * M];
        Model::cost_derivs(sc.x, sc.u, P.mp, true, sc.cx, cu, sc.cxx, cxu, cuu);
      }
    });
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < N * N + N; e += 32) {
        if (e < N * N) sc.Vxx[e] = sc.cxx[e];
        else sc.Vx[e - N * N] = sc.cx[e - N * N];
      }
    });
  }

  /* ---- backward pass -------------------------------------------------------------------- */

  /* One timestep of the backward recursion for tile entry tt; returns false when the boxQP reports
   * failure (result < 1, src/ilqr_core.cpp:371).  On success sc.kc / sc.Kc hold k_i / K_i and
   * sc.Vx / sc.Vxx the value function at i. */
  ILQR_HD bool backward_step(int tt, S lam) {
    const S *F = sc.Ft + tt * NM * N; /* F[j][r]: column j of [fx | fu] */
    const S *ut = sc.us + tt * M;
    const S *cx, *cu, *cxx, *cxu, *cuu;
    if constexpr (CD == kCostFD) {
      cx = sc.Ct + tt * NC;
      cu = cx + N;
      cxx = cu + M;
      cxu = cxx + N * N;
      cuu = cxu + N * M;
    } else {
      cx = sc.cx;
      cu = sc.cu;
      cxx = sc.cxx;
      cxu = sc.cxu;
      cuu = sc.cuu;
    }
    /* W = F^T Vxx'  (the inner product of :361-363); closed-form cost derivatives ride on the last lane */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < NM * N; e += 32) {
        const int c = e / N, b = e % N;
        S acc = 0;
#pragma unroll
        for (int r = 0; r < N; r++) acc += F[c * N + r] * sc.Vxx[r * N + b];
        sc.W[e] = acc;
      }
      if (CD == kCostAnalytic && lane == 31)
        Model::cost_derivs(sc.xs + tt * N, ut, P.mp, false, sc.cx, sc.cu, sc.cxx, sc.cxu, sc.cuu);
    });
    /* Qx, Qu (:359-360), Qxx, Qux, Quu and the regularised QuuF (:361-367); QuuF goes straight into the QP */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < N * N + M * N + M * M + NM; e += 32) {
        if (e < N * N) {
          const int a = e / N, b = e % N;
          S acc = 0;
#pragma unroll
          for (int r = 0; r < N; r++) acc += sc.W[a * N + r] * F[b * N + r];
          sc.Qxx[e] = cxx[e] + acc;
        } else if (e < N * N + M * N) {
          const int q = e - N * N, a = q / N, b = q % N;
          S acc = 0;
#pragma unroll
          for (int r = 0; r < N; r++) acc += sc.W[(N + a) * N + r] * F[b * N + r];
          sc.Qux[q] = cxu[b * M + a] + acc;
        } else if (e < N * N + M * N + M * M) {
          const int q = e - N * N - M * N, a = q / M, b = q % M;
          S acc = 0;
#pragma unroll
          for (int r = 0; r < N; r++) acc += sc.W[(N + a) * N + r] * F[(N + b) * N + r];
          sc.Quu[q] = cuu[q] + acc;
          sc.qp.Q[q] = (cuu[q] + (a == b ? lam : S(0))) + acc;
        } else {
          const int c = e - N * N - M * N - M * M;
          S acc = 0;
#pragma unroll
          for (int r = 0; r < N; r++) acc += F[c * N + r] * sc.Vx[r];
          if (c < N) sc.Qx[c] = cx[c] + acc;
          else sc.Qu[c - N] = cu[c - N] + acc;
        }
      }
    });
    /* boxQP, gains, dV (:369-389) — one lane */
    ex.lanes([&](int lane, Lane &) {
      if (lane != 0) return;
      QPWork<M, S> &w = sc.qp;
#pragma unroll
      for (int j = 0; j < M; j++) {
        w.c[j] = sc.Qu[j];
        w.x0[j] = sc.kprev[j];
        w.lo[j] = P.u_min[j] - ut[j];
        w.hi[j] = P.u_max[j] - ut[j];
      }
      box_qp<M, S>(P.qp, w);
      if (w.result < 1) return;
#pragma unroll
      for (int j = 0; j < M; j++) sc.kc[j] = w.x[j];
      for (int e = 0; e < M * N; e++) sc.Kc[e] = 0;
      if constexpr (M == 1) {
        if (w.v_free[0]) {
#pragma unroll
          for (int b = 0; b < N; b++) sc.Kc[b] = (-w.Hinv[0]) * sc.Qux[b];
        }
      } else {
        const int r = w.r_dim;
        int q = 0;
        for (int j = 0; j < M; j++)
          if (w.v_free[j]) w.idx[q++] = j;
        if (q > 0) {
          for (int a = 0; a < r && a < q; a++)
            for (int b = 0; b < N; b++) {
              S acc = 0;
              for (int c = 0; c < r && c < q; c++) acc += (-w.Hinv[a * r + c]) * sc.Qux[w.idx[c] * N + b];
              sc.Kc[w.idx[a] * N + b] = acc;
            }
        }
      }
      S a0 = 0; /* :388-389, unregularised Quu */
#pragma unroll
      for (int j = 0; j < M; j++) a0 += sc.kc[j] * sc.Qu[j];
      sc.st.dV0 += a0;
      S a1 = 0;
      S row[M];
#pragma unroll
      for (int b = 0; b < M; b++) {
        S acc = 0;
#pragma unroll
        for (int a = 0; a < M; a++) acc += (S(0.5) * sc.kc[a]) * sc.Quu[a * M + b];
        row[b] = acc;
      }
#pragma unroll
      for (int b = 0; b < M; b++) a1 += row[b] * sc.kc[b];
      sc.st.dV1 += a1;
    });
    if (sc.qp.result < 1) return false;
    /* Vx, Vxx before symmetrisation (:391-392) */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < N * N + N; e += 32) {
        const int a = (e < N * N) ? e / N : e - N * N;
        S ktq[M]; /* row a of K^T Quu */
#pragma unroll
        for (int b = 0; b < M; b++) {
          S acc = 0;
#pragma unroll
          for (int c = 0; c < M; c++) acc += sc.Kc[c * N + a] * sc.Quu[c * M + b];
          ktq[b] = acc;
        }
        if (e < N * N) {
          const int b = e % N;
          S t1 = 0, t2 = 0, t3 = 0;
#pragma unroll
          for (int c = 0; c < M; c++) t1 += ktq[c] * sc.Kc[c * N + b];
#pragma unroll
          for (int c = 0; c < M; c++) t2 += sc.Kc[c * N + a] * sc.Qux[c * N + b];
#pragma unroll
          for (int c = 0; c < M; c++) t3 += sc.Qux[c * N + a] * sc.Kc[c * N + b];
          sc.Vtmp[e] = sc.Qxx[e] + t1 + t2 + t3;
        } else {
          S t1 = 0, t2 = 0, t3 = 0;
#pragma unroll
          for (int c = 0; c < M; c++) t1 += ktq[c] * sc.kc[c];
#pragma unroll
          for (int c = 0; c < M; c++) t2 += sc.Kc[c * N + a] * sc.Qu[c];
#pragma unroll
          for (int c = 0; c < M; c++) t3 += sc.Qux[c * N + a] * sc.kc[c];
          sc.Vxn[a] = sc.Qx[a] + t1 + t2 + t3;
        }
      }
    });
    /* symmetrise (:393), roll the value function, remember k for the next warm start, stage k / K (:396-397) */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < N * N + N + M + M * N + M; e += 32) {
        if (e < N * N) {
          const int a = e / N, b = e % N;
          sc.Vxx[e] = S(0.5) * (sc.Vtmp[a * N + b] + sc.Vtmp[b * N + a]);
        } else if (e < N * N + N) {
          sc.Vx[e - N * N] = sc.Vxn[e - N * N];
        } else if (e < N * N + N + M) {
          sc.kprev[e - N * N - N] = sc.kc[e - N * N - N];
        } else if (e < N * N + N + M + M * N) {
          sc.K[tt * M * N + e - (N * N + N + M)] = sc.Kc[e - (N * N + N + M)];
        } else {
          sc.k[tt * M + e - (N * N + N + M + M * N)] = sc.kc[e - (N * N + N + M + M * N)];
        }
      }
    });
    return true;
  }

  /* iLQR::backward_pass.  Returns the failing timestep or 0 (:371,400). */
  ILQR_HD int backward_pass(S lam) {
    const int T = P.T;
    ex.lanes([&](int lane, Lane &) {
      if (lane < N) sc.x[lane] = tr.xs[T * N + lane];
      if (lane < M) {
        sc.u[lane] = 0;
        sc.kprev[lane] = tr.k[(T - 1) * M + lane]; /* :369 warm start of i = T-1: the previous pass's k[T-1] */
      }
      if (lane == 0) {
        sc.st.dV0 = 0; /* :356 */
        sc.st.dV1 = 0;
        sc.st.n_backward++;
      }
    });
    phase_terminal();
    int diverged_at = -1;
    for (int ti = (T - 1) / kTileB; ti >= 0 && diverged_at < 0; ti--) {
      const int t0 = ti * kTileB;
      const int cnt = (T - t0 < kTileB) ? T - t0 : kTileB;
      ex.lanes([&](int lane, Lane &) {
        for (int e = lane; e < cnt * N; e += 32) sc.xs[e] = tr.xs[t0 * N + e];
        for (int e = lane; e < cnt * M; e += 32) sc.us[e] = tr.us[t0 * M + e];
        for (int e = lane; e < cnt * NM * N; e += 32) sc.Ft[e] = ld_fresh(sl.F + (size_t)t0 * NM * N + e);
        if constexpr (CD == kCostFD) {
          for (int e = lane; e < cnt * NC; e += 32) sc.Ct[e] = ld_fresh(sl.C + (size_t)t0 * NC + e);
        }
      });
      int first_done = 0; /* tile entries [first_done, cnt) hold finished k / K */
      for (int tt = cnt - 1; tt >= 0; tt--) {
        if (!backward_step(tt, lam)) { /* the steps above this one have already written their k, K (:396-397) */
          diverged_at = t0 + tt;
          first_done = tt + 1;
          break;
        }
      }
      /* flush the tile's k / K and the gradient-norm terms of its timesteps (:405-412), one per lane */
      ex.lanes([&](int lane, Lane &) {
        for (int e = lane + first_done * M * N; e < cnt * M * N; e += 32) tr.K[t0 * M * N + e] = sc.K[e];
        for (int e = lane + first_done * M; e < cnt * M; e += 32) tr.k[t0 * M + e] = sc.k[e];
        for (int e = lane + first_done; e < cnt; e += 32) sl.gterm[t0 + e] = gn_term(sc.k + e * M, sc.us + e * M);
      });
    }
    if (diverged_at >= 0) return diverged_at;
    /* Vx[0], Vxx[0] are results of record for the tests (include/ilqr.h:76-77) */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < N * N + N; e += 32) {
        if (e < N * N) tr.Vxx0[e] = sc.Vxx[e];
        else tr.Vx0[e - N * N] = sc.Vx[e - N * N];
      }
    });
    return 0;
  }

  /* get_gradient_norm (:405-412): mean_t max_j |k_tj| / (|u_tj| + 1), summed in ascending t like the
   * reference, from the terms the backward pass left in gterm ... */
  ILQR_HD void gradient_norm_from_terms() {
    const int T = P.T;
    ex.lanes([&](int lane, Lane &) {
      if (lane != 0) return;
      S acc = 0;
      for (int t = 0; t < T; t++) acc += sl.gterm[t];
      sc.st.gnorm = acc / T;
    });
  }
  /* ... or from k and us in global memory (after a pass that stopped early) */
  ILQR_HD void gradient_norm_only() {
    const int T = P.T;
    ex.lanes([&](int lane, Lane &) {
      if (lane != 0) return;
      S acc = 0;
      for (int t = 0; t < T; t++) acc += gn_term(tr.k + t * M, tr.us + t * M);
      sc.st.gnorm = acc / T;
    });
  }
  ILQR_HD static S gn_term(const S *k, const S *u) {
    S mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < M; j++) {
      const S v = t_abs(k[j]) / (t_abs(u[j]) + 1);
      if (v > mx) mx = v;
    }
    return mx;
  }

  /* ---- rollouts ------------------------------------------------------------------------- */

  /* one step of iLQR::forward_pass (:314-326) for one lane; the applied control is left in L.uc */
  ILQR_HD void rollout_step(Lane &L, const S *xhat, const S *ubar, const S *kt, const S *Kt, S alpha, int mode) {
#pragma unroll
    for (int j = 0; j < M; j++) {
      S v = ubar[j];
      if (mode == kRollClosed) v = ubar[j] + kt[j] * alpha; /* :188-190 */
      if (mode != kRollOpen) {                              /* :316 */
        S a = 0;
#pragma unroll
        for (int i = 0; i < N; i++) a += Kt[j * N + i] * (L.x[i] - xhat[i]);
        v += a;
      }
      L.uc[j] = v;
    }
    L.cost += Model::cost(L.x, L.uc, P.mp); /* :324 */
    S x1[N];
    integrate<Model, S>(L.x, L.uc, P.mp, P.dt, x1); /* :325 */
#pragma unroll
    for (int i = 0; i < N; i++) L.x[i] = x1[i];
  }

  ILQR_HD void load_forward_tile(int t0, int cnt, int mode) {
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < cnt * N; e += 32) sc.xs[e] = tr.xs[t0 * N + e];
      for (int e = lane; e < cnt * M; e += 32) sc.us[e] = tr.us[t0 * M + e];
      if (mode != kRollOpen) {
        for (int e = lane; e < cnt * M * N; e += 32) sc.K[e] = tr.K[t0 * M * N + e];
        for (int e = lane; e < cnt * M; e += 32) sc.k[e] = tr.k[t0 * M + e];
      }
    });
  }

  /* The candidate rollouts of the line search, lane a <-> alpha[a]: costs land in sc.newcost, the
   * candidate's controls and states stream to the warp's candidate buffer. */
  ILQR_HD void rollout_candidates() {
    const int T = P.T;
    const int na = P.n_alpha;
    ex.lanes([&](int lane, Lane &L) {
#pragma unroll
      for (int i = 0; i < N; i++) L.x[i] = tr.x0[i];
      L.cost = 0;
    });
    for (int t0 = 0; t0 < T; t0 += kTile) {
      const int cnt = (T - t0 < kTile) ? T - t0 : kTile;
      load_forward_tile(t0, cnt, kRollClosed);
      ex.lanes([&](int lane, Lane &L) {
        if (lane >= na) return;
        const S alpha = P.alpha[lane];
        S *cx = sl.cand_x + ((size_t)lane * T + t0) * N;
        S *cu = sl.cand_u + ((size_t)lane * T + t0) * M;
        for (int tt = 0; tt < cnt; tt++) {
          rollout_step(L, sc.xs + tt * N, sc.us + tt * M, sc.k + tt * M, sc.K + tt * M * N, alpha, kRollClosed);
#pragma unroll
          for (int j = 0; j < M; j++) cu[tt * M + j] = L.uc[j];
#pragma unroll
          for (int i = 0; i < N; i++) cx[tt * N + i] = L.x[i];
        }
      });
    }
    ex.lanes([&](int lane, Lane &L) {
      if (lane >= na) return;
      L.cost += Model::final_cost(L.x, P.mp); /* :335 */
      sc.newcost[lane] = L.cost;
    });
  }

  /* accept candidate a: xs[1..T], us[0..T-1] <- its rollout (what forward_pass left in the member
   * arrays, :323,334); xs[0] = x0 already */
  ILQR_HD void commit_candidate(int a) {
    const int T = P.T;
    ex.lanes([&](int lane, Lane &) {
      const S *cx = sl.cand_x + (size_t)a * T * N;
      const S *cu = sl.cand_u + (size_t)a * T * M;
      for (int e = lane; e < T * N; e += 32) tr.xs[N + e] = ld_fresh(cx + e);
      for (int e = lane; e < T * M; e += 32) tr.us[e] = ld_fresh(cu + e);
    });
  }

  /* One rollout on lane 0 that rewrites xs, us in place (init_traj, warm start, test hook); the
   * cost is returned in sc.st.new_cost. */
  ILQR_HD void rollout_commit(S alpha, int mode) {
    const int T = P.T;
    ex.lanes([&](int lane, Lane &L) {
      if (lane != 0) return;
#pragma unroll
      for (int i = 0; i < N; i++) L.x[i] = tr.x0[i];
      L.cost = 0;
    });
    for (int t0 = 0; t0 < T; t0 += kTile) {
      const int cnt = (T - t0 < kTile) ? T - t0 : kTile;
      load_forward_tile(t0, cnt, mode);
      ex.lanes([&](int lane, Lane &L) {
        if (lane != 0) return;
        for (int tt = 0; tt < cnt; tt++) {
#pragma unroll
          for (int i = 0; i < N; i++) tr.xs[(t0 + tt) * N + i] = L.x[i];
          rollout_step(L, sc.xs + tt * N, sc.us + tt * M, sc.k + tt * M, sc.K + tt * M * N, alpha, mode);
#pragma unroll
          for (int j = 0; j < M; j++) tr.us[(t0 + tt) * M + j] = L.uc[j];
        }
      });
    }
    ex.lanes([&](int lane, Lane &L) {
      if (lane != 0) return;
#pragma unroll
      for (int i = 0; i < N; i++) tr.xs[T * N + i] = L.x[i];
      L.cost += Model::final_cost(L.x, P.mp);
      sc.st.new_cost = L.cost;
    });
  }

  /* ---- entry points (one per ABI call) -------------------------------------------------- */

  ILQR_HD void load_state() {
    ex.lanes([&](int lane, Lane &) {
      if (lane == 0) sc.st = *tr.st;
    });
  }
  ILQR_HD void store_state() {
    ex.lanes([&](int lane, Lane &) {
      if (lane == 0) *tr.st = sc.st;
    });
  }

  /* iLQR::init_traj (:11-56): open-loop rollout of u0 (already in tr.us), zeroed gains (zeroed by
   * the caller), fresh lambda schedule. */
  ILQR_HD void op_init() {
    ex.lanes([&](int lane, Lane &) {
      if (lane != 0) return;
      TrajState<S> z = {};
      sc.st = z;
    });
    rollout_commit(S(0), kRollOpen);
    ex.lanes([&](int lane, Lane &) {
      if (lane != 0) return;
      TrajState<S> &s = sc.st;
      s.cost = s.new_cost;
      s.dV0 = S(M); /* :32 */
      s.dV1 = 1;
      s.lam = P.lambda_init;
      s.dlam = P.dlambda_init;
      s.flg_change = 1;
      s.status = kRunning;
      s.alpha_index = -1;
    });
    store_state();
  }

  /* iLQR::generate_trajectory(x_0) up to the loop (:65-76): feedback rollout of the kept us around
   * the kept xs from the new x0; lambda/dlambda carry over. */
  ILQR_HD void op_warm_start() {
    load_state();
    rollout_commit(S(0), kRollWarm);
    ex.lanes([&](int lane, Lane &) {
      if (lane != 0) return;
      TrajState<S> &s = sc.st;
      s.cost = s.new_cost;
      s.flg_change = 1;
      s.iter = 0;
      s.status = kRunning;
    });
    store_state();
  }

  ILQR_HD void op_backward_once(S lam) {
    load_state();
    derivative_sweep();
    const int d = backward_pass(lam);
    ex.lanes([&](int lane, Lane &) {
      if (lane != 0) return;
      sc.st.lam = lam;
      sc.st.diverge = d;
    });
    if (d == 0) gradient_norm_from_terms();
    else gradient_norm_only();
    store_state();
  }

  ILQR_HD void op_rollout_once(S alpha) {
    load_state();
    rollout_commit(alpha, kRollClosed);
    ex.lanes([&](int lane, Lane &) {
      if (lane == 0) sc.st.cost = sc.st.new_cost;
    });
    store_state();
  }

  /* the loop body of iLQR::generate_trajectory() (:103-288), up to n_iters trips */
  ILQR_HD void op_iterate(int n_iters) {
    load_state();
    int done_here = 0;
    bool have_derivs = false; /* the warp's F / C buffers hold this trajectory's current derivatives */
    while (sc.st.iter < P.max_iter && done_here < n_iters && sc.st.status == kRunning) {
      done_here++;
      /* :115-120 */
      if (sc.st.flg_change || !have_derivs) {
        derivative_sweep();
        have_derivs = true;
      }
      ex.lanes([&](int lane, Lane &) {
        if (lane != 0) return;
        sc.st.trips++;
        if (sc.st.flg_change) {
          sc.st.flg_change = 0;
          sc.st.n_deriv++;
        }
        sc.flag = 0;
      });
      /* :136-150 */
      bool back_done = false;
      while (!back_done) {
        const int diverge = backward_pass(sc.st.lam);
        ex.lanes([&](int lane, Lane &) {
          if (lane != 0) return;
          TrajState<S> &s = sc.st;
          s.diverge = diverge;
          sc.flag = 0;
          if (diverge != 0) {
            s.dlam = fmax_(s.dlam * P.lambda_factor, P.lambda_factor);
            s.lam = fmax_(s.lam * s.dlam, P.lambda_min);
            if (s.lam > P.lambda_max) sc.flag = 1;
          }
        });
        if (diverge != 0) {
          if (sc.flag) break;
          continue;
        }
        back_done = true;
      }
      if (back_done) gradient_norm_from_terms();
      else gradient_norm_only();
      /* :153-159 */
      ex.lanes([&](int lane, Lane &) {
        if (lane != 0) return;
        sc.flag = 0;
        if (sc.st.gnorm < P.tol_grad && sc.st.lam < P.grad_lambda_gate) {
          sc.st.status = kExitGrad;
          sc.flag = 2;
        }
      });
      if (sc.flag == 2) break; /* gradient exit: `break` before iter++ */
      if (back_done) rollout_candidates();
      /* the acceptance test :199-213 in the reference's serial order */
      ex.lanes([&](int lane, Lane &) {
        if (lane != 0) return;
        TrajState<S> &s = sc.st;
        sc.flag = 0;
        s.alpha_index = -1;
        S alpha = 0;
        if (back_done) {
          for (int a = 0; a < P.n_alpha; a++) {
            alpha = P.alpha[a];
            s.new_cost = sc.newcost[a];
            s.n_rollouts++;
            s.dcost = s.cost - s.new_cost;
            s.expected = -alpha * (s.dV0 + alpha * s.dV1);
            S z;
            if (s.expected > 0) z = s.dcost / s.expected;
            else z = S((S(0) < s.dcost) - (s.dcost < S(0))); /* sgn, include/common.h:43-44 */
            if (z > P.z_min) {
              s.alpha_index = a;
              sc.flag = 1;
              break;
            }
          }
          if (!sc.flag) alpha = 0;
        }
        s.alpha = alpha;
      });
      const bool fwd_done = sc.flag == 1;
      if (fwd_done) commit_candidate(sc.st.alpha_index);
      ex.lanes([&](int lane, Lane &) {
        if (lane != 0) return;
        TrajState<S> &s = sc.st;
        sc.flag = 0;
        if (fwd_done) { /* :242-263 */
          s.dlam = fmin_(s.dlam / P.lambda_factor, 1 / P.lambda_factor);
          s.lam = s.lam * s.dlam * S(s.lam > P.lambda_min);
          s.cost = s.new_cost;
          s.flg_change = 1;
          s.n_accept++;
          if (s.dcost < P.tol_fun) {
            s.status = kExitTolFun;
            sc.flag = 1;
          }
        } else { /* :264-282 */
          s.dlam = fmax_(s.dlam * P.lambda_factor, P.lambda_factor);
          s.lam = fmax_(s.lam * s.dlam, P.lambda_min);
          s.n_reject++;
          if (s.lam > P.lambda_max) {
            s.status = kExitLambdaMax;
            sc.flag = 1;
          }
        }
        if (!sc.flag) s.iter++;
      });
      if (sc.flag) break;
    }
    ex.lanes([&](int lane, Lane &) {
      if (lane != 0) return;
      if (sc.st.status == kRunning && sc.st.iter >= P.max_iter) sc.st.status = kExitMaxIter;
    });
    store_state();
  }

  ILQR_HD static S fmax_(S a, S b) { return a < b ? b : a; } /* std::max(a, b) */
  ILQR_HD static S fmin_(S a, S b) { return b < a ? b : a; } /* std::min(a, b) */
};

}  // namespace ilqr
#endif
