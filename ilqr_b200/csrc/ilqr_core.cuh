/*
 * ilqr_core.cuh — one iLQR problem instance, solved by ONE WARP.
 *
 * This is the whole hot path of the reference for one trajectory, re-designed for a 32-lane warp
 * that owns the trajectory from the first rollout to termination:
 *
 *   backward pass (src/ilqr_core.cpp:350-401), serial in t, per timestep
 *     - lanes 0 .. 2(n+m)-1 evaluate the perturbed Euler steps of the central-difference
 *       Jacobians fx, fu (src/derivatives.cpp:15-26, include/finite_diff.h:35-47);
 *       the remaining lanes evaluate the cost stencils (src/derivatives.cpp:29-144,
 *       finite_diff.h:22-33,67-86) or the closed forms — derivatives are recomputed on the fly
 *       and never written to HBM;
 *     - the (n+m)^2 Q-function entries are spread one per lane (:359-367);
 *     - one lane runs the boxQP (src/boxqp.cpp) and the gains (:369-389);
 *     - the n + n^2 entries of Vx, Vxx are spread one per lane (:391-393);
 *     Vx/Vxx and every intermediate live in the warp's shared-memory scratch.
 *   line search (src/ilqr_core.cpp:184-226): the n_alpha candidate rollouts (:305-337) run
 *     concurrently, one per lane; the first accepted index is taken, which is result-identical
 *     to the reference's serial early-exit loop because every candidate starts from the same
 *     (xs, us, K, k);  the accepted candidate is then re-rolled to commit xs/us in place.
 *   outer loop, lambda schedule, termination (src/ilqr_core.cpp:103-288): per-trajectory scalars
 *     in the scratch; lambda/dlambda are per trajectory (process-wide statics in the reference).
 *
 * HBM traffic is the minimum the algorithm allows: a backward pass reads xs, us and writes K, k;
 * the line search reads xs, us, K, k; a commit rewrites xs, us.  The per-trajectory arrays are
 * contiguous in t, moved in 32-timestep tiles by coalesced warp-wide copies.
 *
 * Lanes communicate only through the scratch (never shuffles), in phases separated by a warp
 * barrier.  That lets tests/emu/ compile this same header with g++ and run the phases lane by
 * lane on the CPU (`HostExec`), so the control flow and the arithmetic order can be checked
 * bit-for-bit against the oracle without a GPU.  The emulator is test infrastructure; the
 * product only ever instantiates `WarpExec`.
 *
 * Arithmetic order follows the reference statement by statement (sequential accumulations from
 * zero, no FMA contraction: the .cu is built with -fmad=false) so that f64 results differ from
 * the reference only through sin/cos.
 */
#ifndef ILQR_CORE_CUH_
#define ILQR_CORE_CUH_

#include <type_traits>

#include "boxqp.cuh"
#include "models.cuh"

namespace ilqr {

constexpr int kTile = 32;      /* timesteps per staged tile */
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

/* the warp's working set */
#if defined(__CUDACC__)
#define ILQR_HDC __host__ __device__ constexpr
#else
#define ILQR_HDC constexpr
#endif
ILQR_HDC int popc_(unsigned v) { return v ? (int)(v & 1u) + popc_(v >> 1) : 0; }

/* f(integral_constant<int, 0>) ... f(integral_constant<int, G-1>): loops whose index must be a
 * compile-time constant (so per-model tables fold to literals in device code) */
template <int G, class F>
ILQR_HD void static_for(F &&f) {
  if constexpr (G > 0) {
    static_for<G - 1>(f);
    f(std::integral_constant<int, G - 1>{});
  }
}
ILQR_HD int popcnt(unsigned v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}

/* finite-difference variants of the model's trig arguments: argument g is needed at the base point
 * and at +-eps on each state variable it depends on */
template <class Model>
struct TrigVariants {
  static constexpr int KT = Model::kTrig;
  static ILQR_HDC int count(int g) { return 1 + 2 * popc_(Model::trig_deps(g)); }
  static ILQR_HDC int offset(int g) { return g <= 0 ? 0 : offset(g - 1) + count(g - 1); }
  static constexpr int total = offset(KT);
};

template <int N, int M, typename S, int NTV = 1, int KT = 1>
struct Scratch {
  static constexpr int NM = N + M;
  /* staged tiles */
  S xs[kTile * N], us[kTile * M], K[kTile * M * N], k[kTile * M];
  S xn[kTile * N], un[kTile * M];
  /* one timestep */
  S x[N], u[M];
  S E[2 * NM * N];      /* perturbed Euler steps: row 2j = +eps on variable j, 2j+1 = -eps */
  S F[N * NM];          /* [fx | fu], N x (N+M) row-major */
  S cx[N], cu[M], cxx[N * N], cxu[N * M], cuu[M * M];
  S Vx[N], Vxx[N * N];  /* value function at i+1, overwritten with i at the end of the step */
  S W[NM * N];          /* F^T Vxx' */
  S Qx[N], Qu[M], Qxx[N * N], Qux[M * N], Quu[M * M];
  S Kc[M * N], kc[M], kprev[M];
  S Vtmp[N * N], Vxn[N];
  S newcost[kMaxAlpha];
  S bsn[NTV > 0 ? NTV : 1], bcs[NTV > 0 ? NTV : 1];                   /* backward: sincos per (argument, FD variant) */
  QPWork<M, S> qp;
  TrajState<S> st;
  int flag;
};

template <int N, int M, typename S>
struct LaneRegs {
  S x[N];
  S uc[M];
  S cost;
  S gn;
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

template <class Model, typename S, int CD, class Exec>
struct Core {
  static constexpr int N = Model::N, M = Model::M, NM = N + M;
  static constexpr int KT = Model::kTrig;
  using TV = TrigVariants<Model>;
  using Sc = Scratch<N, M, S, TV::total, KT>;
  using Lane = LaneRegs<N, M, S>;

  const SolveParams<S> &P;
  Sc &sc;
  Exec &ex;
  TrajPtrs<S> tr;
  S *gterm; /* [T] per-timestep terms of the gradient norm, in the warp's shared memory after the scratch */

  ILQR_HD Core(const SolveParams<S> &p, Sc &s, S *g, Exec &e, const TrajPtrs<S> &t) : P(p), sc(s), ex(e), tr(t), gterm(g) {}

  /* ---- tiles ---------------------------------------------------------------------------- */
  ILQR_HD void copy_in(S *dst, const S *src, int count) {
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < count; e += 32) dst[e] = src[e];
    });
  }
  ILQR_HD void copy_out(S *dst, const S *src, int count) {
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < count; e += 32) dst[e] = src[e];
    });
  }

  /* ---- derivatives ---------------------------------------------------------------------- */

  /* perturbed copy helpers: v[q] (+ sa*eps if q == a) (+ sb*eps if q == b), in that order */
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

  /* One finite-difference cost-stencil output (src/derivatives.cpp:29-144).  Output ids:
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

  ILQR_HD void cost_stencil(int o, bool terminal) {
    const S eps = P.fd_eps;
    const S *mp = P.mp;
    S xa[N], ua[M];
    auto fx = [&](const S *xv, const S *uv) -> S { return terminal ? Model::final_cost(xv, mp) : Model::cost(xv, uv, mp); };
    if (o < N) { /* finite_diff_gradient wrt x, finite_diff.h:22-33 */
      perturb<N>(sc.x, o, eps, -1, S(0), xa);
      const S p = fx(xa, sc.u);
      perturb<N>(sc.x, o, -eps, -1, S(0), xa);
      const S m = fx(xa, sc.u);
      sc.cx[o] = (p - m) / (2 * eps);
      return;
    }
    o -= N;
    if (!terminal) {
      if (o < M) {
        perturb<M>(sc.u, o, eps, -1, S(0), ua);
        const S p = Model::cost(sc.x, ua, mp);
        perturb<M>(sc.u, o, -eps, -1, S(0), ua);
        const S m = Model::cost(sc.x, ua, mp);
        sc.cu[o] = (p - m) / (2 * eps);
        return;
      }
      o -= M;
    }
    if (o < kNxx) { /* finite_diff_hessian wrt x, finite_diff.h:67-86 */
      int i, j;
      tri_index(o, N, i, j);
      perturb<N>(sc.x, i, eps, j, eps, xa);
      const S pp = fx(xa, sc.u);
      perturb<N>(sc.x, i, -eps, j, eps, xa);
      const S mpv = fx(xa, sc.u);
      perturb<N>(sc.x, i, eps, j, -eps, xa);
      const S pm = fx(xa, sc.u);
      perturb<N>(sc.x, i, -eps, j, -eps, xa);
      const S mm = fx(xa, sc.u);
      const S v = (pp - mpv - pm + mm) / (4 * eps * eps);
      sc.cxx[i * N + j] = v;
      sc.cxx[j * N + i] = v;
      return;
    }
    o -= kNxx;
    if (terminal) return;
    if (o < kNuu) {
      int i, j;
      tri_index(o, M, i, j);
      perturb<M>(sc.u, i, eps, j, eps, ua);
      const S pp = Model::cost(sc.x, ua, mp);
      perturb<M>(sc.u, i, -eps, j, eps, ua);
      const S mpv = Model::cost(sc.x, ua, mp);
      perturb<M>(sc.u, i, eps, j, -eps, ua);
      const S pm = Model::cost(sc.x, ua, mp);
      perturb<M>(sc.u, i, -eps, j, -eps, ua);
      const S mm = Model::cost(sc.x, ua, mp);
      const S v = (pp - mpv - pm + mm) / (4 * eps * eps);
      sc.cuu[i * M + j] = v;
      sc.cuu[j * M + i] = v;
      return;
    }
    o -= kNuu;
    if (o < N * M) { /* calculate_cxu, src/derivatives.cpp:114-144 (its own stencil) */
      const int i = o / M, j = o % M;
      S xp[N], xm[N], up[M], um[M];
      perturb<N>(sc.x, i, eps, -1, S(0), xp);
      perturb<N>(sc.x, i, -eps, -1, S(0), xm);
      perturb<M>(sc.u, j, eps, -1, S(0), up);
      perturb<M>(sc.u, j, -eps, -1, S(0), um);
      sc.cxu[i * M + j] =
          (Model::cost(xp, up, mp) - Model::cost(xm, up, mp) - Model::cost(xp, um, mp) + Model::cost(xm, um, mp)) /
          (4 * (eps * eps));
    }
  }

  /* Phase: derivatives of timestep i at (sc.x, sc.u) -> sc.E (perturbed steps), sc.c*.
   * Two sub-phases when the model has trig arguments: every (argument, variant) sincos on its own
   * lane (the cost derivatives ride along on the lanes after those), then the 2(n+m) perturbed
   * Euler steps from the tabulated values. */
  ILQR_HD void phase_derivatives() {
    constexpr int nDyn = 2 * NM;
    constexpr int nCost = (CD == kCostFD) ? kStencilStep : 1;
    constexpr int nTrig = TV::total;
    ex.lanes([&](int lane, Lane &) {
      for (int task = lane; task < nTrig + nCost; task += 32) {
        if (KT > 0 && task < nTrig) {
          if constexpr (KT > 0) {
          int g = 0, off = 0;
          unsigned deps = 0;
          static_for<KT>([&](auto gc) {
            constexpr int G = decltype(gc)::value;
            constexpr int o = TV::offset(G);
            constexpr unsigned dd = Model::trig_deps(G);
            if (task >= o) {
              g = G;
              off = o;
              deps = dd;
            }
          });
          const int v = task - off;
          int var = -1;
          S d = 0;
          if (v > 0) {
            int rank = (v - 1) >> 1, seen = 0;
#pragma unroll
            for (int q = 0; q < N; q++)
              if ((deps >> q) & 1u) {
                if (seen == rank) var = q;
                seen++;
              }
            d = ((v - 1) & 1) ? -P.fd_eps : P.fd_eps;
          }
          S xa[N];
          perturb<N>(sc.x, var, d, -1, S(0), xa);
          sincos_det(Model::trig_arg(g, xa), &sc.bsn[task], &sc.bcs[task]);
          }
        } else if (CD == kCostFD) {
          cost_stencil(task - nTrig, false);
        } else {
          Model::cost_derivs(sc.x, sc.u, P.mp, false, sc.cx, sc.cu, sc.cxx, sc.cxu, sc.cuu);
        }
      }
    });
    ex.lanes([&](int lane, Lane &) {
      for (int task = lane; task < nDyn; task += 32) { /* finite_diff_jacobian of integrate_dynamics, finite_diff.h:35-47 */
        const int var = task >> 1, neg = task & 1;
        const S d = neg ? -P.fd_eps : P.fd_eps;
        S xa[N], ua[M], x1[N];
        perturb<N>(sc.x, var, d, -1, S(0), xa);
        perturb<M>(sc.u, var - N, d, -1, S(0), ua);
        if constexpr (KT > 0) {
          S sn[KT], cs[KT];
          static_for<KT>([&](auto gc) { /* variant 0 = base value when the argument does not depend on `var` */
            constexpr int G = decltype(gc)::value;
            constexpr int o = TV::offset(G);
            constexpr unsigned dd = Model::trig_deps(G);
            int variant = 0;
            if (var < N && ((dd >> var) & 1u)) variant = 1 + 2 * popcnt(dd & ((1u << var) - 1u)) + neg;
            sn[G] = sc.bsn[o + variant];
            cs[G] = sc.bcs[o + variant];
          });
          integrate_trig<Model, S>(xa, ua, P.mp, P.dt, sn, cs, x1);
        } else {
          integrate<Model, S>(xa, ua, P.mp, P.dt, x1);
        }
#pragma unroll
        for (int r = 0; r < N; r++) sc.E[task * N + r] = x1[r];
      }
    });
  }

  /* Vx[T] = cx[T], Vxx[T] = cxx[T]  (src/ilqr_core.cpp:353-354) from sc.x = xs[T] */
  ILQR_HD void phase_terminal() {
    ex.lanes([&](int lane, Lane &) {
      if (CD == kCostFD) {
        for (int o = lane; o < kStencilTerm; o += 32) cost_stencil(o, true);
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

  /* One timestep of the backward recursion at (sc.x, sc.u) = (xs[i], us[i]); returns false when
   * the boxQP reports failure (result < 1, src/ilqr_core.cpp:371).  On success sc.kc / sc.Kc hold
   * k_i / K_i and sc.Vx / sc.Vxx the value function at i. */
  ILQR_HD bool backward_step(S lam) {
    phase_derivatives();
    /* F = [fx | fu] column j = (f(+eps e_j) - f(-eps e_j)) / (2 eps) */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < N * NM; e += 32) {
        const int r = e / NM, j = e % NM;
        sc.F[e] = (sc.E[(2 * j) * N + r] - sc.E[(2 * j + 1) * N + r]) / (2 * P.fd_eps);
      }
    });
    /* W = F^T Vxx' ; Qx = cx + fx^T Vx' ; Qu = cu + fu^T Vx'   (:359-360) */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < NM * N + NM; e += 32) {
        if (e < NM * N) {
          const int c = e / N, b = e % N;
          S acc = 0;
#pragma unroll
          for (int r = 0; r < N; r++) acc += sc.F[r * NM + c] * sc.Vxx[r * N + b];
          sc.W[e] = acc;
        } else {
          const int c = e - NM * N;
          S acc = 0;
#pragma unroll
          for (int r = 0; r < N; r++) acc += sc.F[r * NM + c] * sc.Vx[r];
          if (c < N) sc.Qx[c] = sc.cx[c] + acc;
          else sc.Qu[c - N] = sc.cu[c - N] + acc;
        }
      }
    });
    /* Qxx, Qux, Quu and the regularised QuuF (:361-367); QuuF goes straight into the QP */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < N * N + M * N + M * M; e += 32) {
        if (e < N * N) {
          const int a = e / N, b = e % N;
          S acc = 0;
#pragma unroll
          for (int r = 0; r < N; r++) acc += sc.W[a * N + r] * sc.F[r * NM + b];
          sc.Qxx[e] = sc.cxx[e] + acc;
        } else if (e < N * N + M * N) {
          const int q = e - N * N, a = q / N, b = q % N;
          S acc = 0;
#pragma unroll
          for (int r = 0; r < N; r++) acc += sc.W[(N + a) * N + r] * sc.F[r * NM + b];
          sc.Qux[q] = sc.cxu[b * M + a] + acc;
        } else {
          const int q = e - N * N - M * N, a = q / M, b = q % M;
          S acc = 0;
#pragma unroll
          for (int r = 0; r < N; r++) acc += sc.W[(N + a) * N + r] * sc.F[r * NM + N + b];
          sc.Quu[q] = sc.cuu[q] + acc;
          sc.qp.Q[q] = (sc.cuu[q] + (a == b ? lam : S(0))) + acc;
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
        w.lo[j] = P.u_min[j] - sc.u[j];
        w.hi[j] = P.u_max[j] - sc.u[j];
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
    /* symmetrise (:393), roll the value function, remember k for the next warm start */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < N * N + N + M; e += 32) {
        if (e < N * N) {
          const int a = e / N, b = e % N;
          sc.Vxx[e] = S(0.5) * (sc.Vtmp[a * N + b] + sc.Vtmp[b * N + a]);
        } else if (e < N * N + N) {
          sc.Vx[e - N * N] = sc.Vxn[e - N * N];
        } else {
          sc.kprev[e - N * N - N] = sc.kc[e - N * N - N];
        }
      }
    });
    return true;
  }

  /* iLQR::backward_pass.  Returns the failing timestep or 0 (:371,400). */
  ILQR_HD int backward_pass(S lam) {
    const int T = P.T;
    copy_in(sc.x, tr.xs + T * N, N);
    ex.lanes([&](int lane, Lane &) {
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
    for (int ti = (T - 1) / kTile; ti >= 0; ti--) {
      const int t0 = ti * kTile;
      const int cnt = (T - t0 < kTile) ? T - t0 : kTile;
      ex.lanes([&](int lane, Lane &) {
        for (int e = lane; e < cnt * N; e += 32) sc.xs[e] = tr.xs[t0 * N + e];
        for (int e = lane; e < cnt * M; e += 32) sc.us[e] = tr.us[t0 * M + e];
      });
      for (int tt = cnt - 1; tt >= 0; tt--) {
        ex.lanes([&](int lane, Lane &) {
          if (lane < N) sc.x[lane] = sc.xs[tt * N + lane];
          else if (lane < NM) sc.u[lane - N] = sc.us[tt * M + lane - N];
        });
        const bool ok = backward_step(lam);
        if (!ok) { /* the steps above this one have already written their k, K (:396-397) */
          const int done0 = tt + 1;
          ex.lanes([&](int lane, Lane &) {
            for (int e = lane + done0 * M * N; e < cnt * M * N; e += 32) tr.K[t0 * M * N + e] = sc.K[e];
            for (int e = lane + done0 * M; e < cnt * M; e += 32) tr.k[t0 * M + e] = sc.k[e];
          });
          return t0 + tt;
        }
        ex.lanes([&](int lane, Lane &) {
          for (int e = lane; e < M * N + M + 1; e += 32) {
            if (e < M * N) sc.K[tt * M * N + e] = sc.Kc[e];
            else if (e < M * N + M) sc.k[tt * M + e - M * N] = sc.kc[e - M * N];
            else gterm[t0 + tt] = gn_term(sc.kc, sc.u);
          }
        });
      }
      ex.lanes([&](int lane, Lane &) {
        for (int e = lane; e < cnt * M * N; e += 32) tr.K[t0 * M * N + e] = sc.K[e];
        for (int e = lane; e < cnt * M; e += 32) tr.k[t0 * M + e] = sc.k[e];
      });
    }
    /* Vx[0], Vxx[0] are results of record for the tests (include/ilqr.h:76-77) */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < N * N + N; e += 32) {
        if (e < N * N) tr.Vxx0[e] = sc.Vxx[e];
        else tr.Vx0[e - N * N] = sc.Vx[e - N * N];
      }
    });
    return 0;
  }

  /* get_gradient_norm (:405-412) as its own ascending loop; the solve loop gets the same number
   * for free from the line-search rollout. */
  ILQR_HD void gradient_norm_only() {
    const int T = P.T;
    ex.lanes([&](int lane, Lane &L) {
      if (lane != 0) return;
      S acc = 0;
      for (int t = 0; t < T; t++) acc += gn_term(tr.k + t * M, tr.us + t * M);
      sc.st.gnorm = acc / T;
    });
  }
  /* the same number from the terms the backward pass just left in gterm (ascending t, like the reference) */
  ILQR_HD void gradient_norm_from_terms() {
    const int T = P.T;
    ex.lanes([&](int lane, Lane &) {
      if (lane != 0) return;
      S acc = 0;
      for (int t = 0; t < T; t++) acc += gterm[t];
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

  /* The candidate rollouts of the line search, lane a <-> alpha[a]; costs land in sc.newcost. */
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
        for (int tt = 0; tt < cnt; tt++)
          rollout_step(L, sc.xs + tt * N, sc.us + tt * M, sc.k + tt * M, sc.K + tt * M * N, alpha, kRollClosed);
      });
    }
    ex.lanes([&](int lane, Lane &L) {
      if (lane >= na) return;
      L.cost += Model::final_cost(L.x, P.mp); /* :335 */
      sc.newcost[lane] = L.cost;
    });
  }

  /* One rollout that commits xs, us in place (what forward_pass does to the member arrays,
   * :323,334); returns the cost in sc.st.new_cost. */
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
          for (int i = 0; i < N; i++) sc.xn[tt * N + i] = L.x[i];
          rollout_step(L, sc.xs + tt * N, sc.us + tt * M, sc.k + tt * M, sc.K + tt * M * N, alpha, mode);
#pragma unroll
          for (int j = 0; j < M; j++) sc.un[tt * M + j] = L.uc[j];
        }
      });
      ex.lanes([&](int lane, Lane &) {
        for (int e = lane; e < cnt * N; e += 32) tr.xs[t0 * N + e] = sc.xn[e];
        for (int e = lane; e < cnt * M; e += 32) tr.us[t0 * M + e] = sc.un[e];
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
    while (sc.st.iter < P.max_iter && done_here < n_iters && sc.st.status == kRunning) {
      done_here++;
      /* derivatives are recomputed inside every backward pass, so flgChange (:115-120) only
       * feeds the sweep counter */
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
      if (back_done) {
        gradient_norm_from_terms();
        rollout_candidates();
      } else {
        gradient_norm_only();
      }
      /* :153-159, then the acceptance test :199-213 in the reference's serial order */
      ex.lanes([&](int lane, Lane &) {
        if (lane != 0) return;
        TrajState<S> &s = sc.st;
        sc.flag = 0;
        if (s.gnorm < P.tol_grad && s.lam < P.grad_lambda_gate) {
          s.status = kExitGrad;
          sc.flag = 2;
          return;
        }
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
      if (sc.flag == 2) break; /* gradient exit: `break` before iter++ */
      const bool fwd_done = sc.flag == 1;
      if (fwd_done) rollout_commit(sc.st.alpha, kRollClosed);
      ex.lanes([&](int lane, Lane &) {
        if (lane != 0) return;
        TrajState<S> &s = sc.st;
        sc.flag = 0;
        if (fwd_done) { /* :242-263 */
          s.dlam = fmin_(s.dlam / P.lambda_factor, 1 / P.lambda_factor);
          s.lam = s.lam * s.dlam * S(s.lam > P.lambda_min);
          s.cost = sc.newcost[s.alpha_index];
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
