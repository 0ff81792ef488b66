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
 *      they are spread over the 32 lanes — one lane per (timestep, configuration variable) for full
 *      +eps / -eps Euler steps, one lane per timestep for the remaining variables, whose perturbed
 *      points share the configuration-dependent part of the dynamics — each writing Jacobian columns.  With finite-difference cost
 *      derivatives (src/derivatives.cpp:29-144) the T*20 stencil outputs are spread the same way.
 *      The columns go to a per-warp buffer that stays in L2 (acrobot, T = 200: 32 KB).
 *   2. backward pass (:350-401), serial in t, three warp phases per timestep (four when m > 1): F^T [Vxx' | Vx']
 *      entries one per lane; the (n+m)^2 Q-function entries one per lane (:359-367); then boxQP + gains
 *      (src/boxqp.cpp, :369-389) run by EVERY lane in registers when there is one control — same
 *      instructions, nothing to broadcast — each lane going straight on to its own entry of Vx, Vxx with the
 *      symmetrisation folded in (:391-393).  Vx, Vxx and the Q-function live in the warp's shared-memory
 *      scratch.  A lone warp issues in order and a trajectory is one long dependent chain, so the phases are
 *      written as single instruction streams (selects, no divergent branches) with their operands loaded up
 *      front; tools/ubench_backward.cu measures them in isolation.
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
enum { kFlagClampRollout = 2, kFlagAnalyticDyn = 4 }; /* = ILQR_FLAG_CLAMP_ROLLOUT, ILQR_FLAG_ANALYTIC_DYN */

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
  int flags;  /* ILQR_FLAG_* of include/ilqr_b200.h that change the arithmetic: kFlagClampRollout, kFlagAnalyticDyn */
  int bulk_f; /* host-set: every tile of the per-warp Jacobian buffer starts 16-byte aligned (bulk copy allowed) */
  int bulk_c; /* the same for the per-warp buffer of finite-difference cost derivatives */
};

/* per-trajectory solver state that survives between launches (one slot per trajectory) */
template <typename S>
struct TrajState {
  S cost, lam, dlam, dV0, dV1, gnorm, dcost, expected, alpha, new_cost;
  int iter, trips, status, alpha_index, n_accept, n_reject, n_backward, diverge, flg_change, n_rollouts, n_deriv;
  int roll; /* phase engine (ilqr_phases.cuh): what the backward phase decided for this trip's line search */
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
  S *C;      /* [T][NCF]      finite-difference cost derivatives, full layout, FD mode only          */
  S *cand_x; /* [n_alpha][T][n]  candidate states x_1..x_T of the line search                       */
  S *cand_u; /* [n_alpha][T][m]  candidate controls                                                  */
  S *gterm;  /* [T] per-timestep terms of the gradient norm (shared memory on the device)           */
};

/* the warp's working set */
template <int N, int M, typename S, int CD>
struct Scratch {
  static constexpr int NM = N + M;
  static constexpr int NA = N + 1;              /* columns of an x-indexed matrix augmented with its vector */
  static constexpr int NCF = NM + NM * NM;      /* cost derivatives of one timestep, full layout (Core::cost_stencil) */
  /* staged tiles: rollouts use kTile timesteps, the backward pass the first kTileB of the same arrays */
  S xs[kTile * N], us[kTile * M], K[kTile * M * N], k[kTile * M];
  alignas(16) S Ft[kTileB * NM * N];         /* backward: Jacobian columns of the tile (bulk-copy target) */
  alignas(16) S Ct[CD == kCostFD ? kTileB * NCF : 1]; /* backward: FD cost derivatives of the tile (full layout; bulk-copy target) */
  /* one timestep */
  S x[N], u[M];         /* xs[T] and a zero control for the terminal derivatives */
  S Cf[NCF];            /* terminal cost derivatives, full layout */
  S Va[N * NA];         /* [Vxx | Vx] at i+1, overwritten with i at the end of the step */
  S W[NM * NA];         /* F^T [Vxx' | Vx'] */
  S Qg[NM * (NM + 1)];  /* the Q-function over the stacked variable v = (x, u): Qg[c][d] = Q_{v_c v_d}, d < n + m, and
                           Qg[c][n + m] = Q_{v_c}  (Qxx, Qxu / Qux, Quu; Qx, Qu) */
  S Ka[M * NA];         /* [K_i | k_i]  (m > 1; with one control the gains stay in registers) */
  S kprev[M];           /* (m > 1) */
  S newcost[kMaxAlpha];
  QPWork<M, S> qp;
  TrajState<S> st;
  int flag;
};

template <int N, int M, typename S>
struct LaneRegs {
  S x[N];
  S cost;
  S dV[2];   /* backward pass: the expected-reduction sums (lane 0's copy is the one of record) */
  S kprev;   /* backward pass, one control: warm start of the next boxQP, kept by every lane */
};

#if defined(__CUDACC__)
#if defined(ILQR_PHASE_CLOCKS)
static __device__ unsigned long long g_phase_clk[32], g_phase_cnt[32];
#endif
/* device: the calling thread is one lane; a phase ends with a warp barrier.
 *
 * stage_issue / stage_wait move one contiguous tile global -> shared with the 1-D bulk form of TMA
 * (cp.async.bulk, completion counted in bytes on the warp's mbarrier): one instruction on lane 0 instead of
 * five load/store pairs on every lane for the 1.25 KB Jacobian tile, and the copy runs while the lanes
 * stage the small xs / us tiles themselves.  Unaligned tiles fall back to a coalesced lane copy.  (Staging
 * EVERY tile this way with double buffering and prefetch was measured and is slower at this occupancy:
 * profiles/experiments/.) */
template <int N, int M, typename S, int G = 32>
struct WarpExec {
  static constexpr int kLanes = G; /* lanes that cooperate on one trajectory: 32, or 16 (two trajectories per warp) */
  LaneRegs<N, M, S> regs;
  int lane;             /* 0 .. G-1 within the trajectory's lane group */
  unsigned mask;        /* the lane group's members, for the group barrier */
  unsigned bar;         /* shared-window address of the group's mbarrier */
  unsigned phase = 0;   /* bit 0: parity to wait for; bit 1: a bulk copy is outstanding */

  template <class Fn>
  __device__ __forceinline__ void lanes(Fn fn) {
    fn(lane, regs);
    __syncwarp(mask);
  }
  /* A value lane 0 left in the scratch, read by every lane of the group for a uniform branch.  The barrier AFTER the
   * read keeps a fast lane from overwriting the slot in the next phase before a slow lane has read it (lanes of a warp
   * are not guaranteed to run in lockstep; compute-sanitizer racecheck flags the pattern without it). */
  template <typename V>
  __device__ __forceinline__ V uniform(const V &slot) {
    const V v = slot;
    __syncwarp(mask);
    return v;
  }
  /* experiment builds only (-DILQR_PHASE_CLOCKS): cycles since the previous tick are charged to phase `id` */
#if defined(ILQR_PHASE_CLOCKS)
  long long last_clk = 0;
  __device__ __forceinline__ void tick(int id) {
    if (lane == 0) {
      const long long now = clock64();
      atomicAdd(&g_phase_clk[id], (unsigned long long)(now - last_clk));
      atomicAdd(&g_phase_cnt[id], 1ULL);
      last_clk = now;
    }
  }
#else
  __device__ __forceinline__ void tick(int) {}
#endif
  __device__ __forceinline__ static unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
  __device__ __forceinline__ void init_barrier(unsigned long long *b) {
    bar = s32(b);
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp(mask);
  }
  /* global writes of this lane group (generic proxy) that a later bulk copy (async proxy) will read */
  __device__ __forceinline__ void publish() {
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp(mask);
  }
  /* Any number of stage_issue calls, then one stage_wait: every copy announces its bytes on the group's mbarrier
   * (expect_tx) before it is issued, and the group's single arrival is made in stage_wait, so the barrier phase
   * completes exactly when all of them have landed. */
  __device__ __forceinline__ void stage_issue(S *dst, const S *src, int count, bool aligned) {
    const unsigned bytes = (unsigned)count * (unsigned)sizeof(S);
    if (aligned && (bytes & 15u) == 0) {
      if (lane == 0) {
        if (!(phase & 2u)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* earlier generic reads of the destinations */
        asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
                     "l"(src), "r"(bytes), "r"(bar)
                     : "memory");
      }
      phase |= 2u;
    } else {
      for (int e = lane; e < count; e += G) dst[e] = src[e];
    }
  }
  __device__ __forceinline__ void stage_wait() {
    if (phase & 2u) {
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
      unsigned done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(phase & 1u)
            : "memory");
      }
      phase = (phase ^ 1u) & 1u;
    }
    __syncwarp(mask);
  }
};
#endif
/* host (tests only): run the phase for lane 0..31 in turn; a staged tile is a plain copy */
template <int N, int M, typename S, int G = 32>
struct HostExec {
  static constexpr int kLanes = G;
  LaneRegs<N, M, S> regs[G];
  template <class Fn>
  void lanes(Fn fn) {
    for (int l = 0; l < G; l++) fn(l, regs[l]);
  }
  void publish() {}
  void stage_issue(S *dst, const S *src, int count, bool) {
    for (int e = 0; e < count; e++) dst[e] = src[e];
  }
  void stage_wait() {}
  void tick(int) {}
  template <typename V>
  V uniform(const V &slot) { return slot; }
};

/* Accumulates products in index order.  The reference's sums start from zero (Eigen zero-initialises, the
 * oracle writes `acc = 0; acc += ...`); 0 + p == p for every p except the sign of a zero, so the first
 * product is taken as it is and one dependent add per dot product is saved. */
template <typename S>
struct Acc {
  S v = 0;
  bool any = false;
  ILQR_HD void add(S p) {
    v = any ? v + p : p;
    any = true;
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

/* ---- finite differences ----------------------------------------------------------------------------------------
 * The reference's stencils (include/finite_diff.h:22-86, src/derivatives.cpp:15-144) with eps = 1e-3, evaluated in
 * DOUBLE whatever the scalar type S of the solve: in f32 a central difference of the Euler step carries
 * ulp(x) / (2 eps) ~ 3e-5 of relative noise, which the Riccati recursion over hundreds of unstable steps amplifies
 * until the f32 solve has nothing to do with the f64 one after a handful of trips (measured, tests: K off by O(1)
 * after 5 trips), and a cost Hessian stencil is pure noise (ulp(4000) / 4 eps^2 ~ 60).  Forming the differences in
 * double and rounding the RESULT to S keeps the derivative arrays accurate to S's own precision; everything else
 * of an f32 solve (rollouts, backward pass, boxQP, storage) stays f32.  For S = double nothing changes. */
template <int D, typename W>
ILQR_HD void fd_perturb(const W *v, int a, W da, int b, W db, W *out) { /* v[q] (+ da if q == a) (+ db if q == b) */
#pragma unroll
  for (int q = 0; q < D; q++) {
    W t = v[q];
    if (q == a) t += da;
    if (q == b) t += db;
    out[q] = t;
  }
}
/* does the model twin offer a closed-form Jacobian (Model::dynamics_jac)?  Optional: user models may leave it out */
template <class M, class = void>
struct HasDynamicsJac {
  static constexpr bool value = false;
};
template <class M>
struct HasDynamicsJac<M, decltype(M::template dynamics_jac<double>(nullptr, nullptr, nullptr, nullptr, nullptr), void())> {
  static constexpr bool value = true;
};

template <class Model, typename S>
struct FiniteDiff {
  typedef double W;
  static constexpr int N = Model::N, M = Model::M, NM = N + M;
  struct Point {
    W x[N], u[M], mp[16], dt, eps;
  };
  ILQR_HD static void load(const SolveParams<S> &P, const S *x, const S *u, Point &p) {
#pragma unroll
    for (int i = 0; i < N; i++) p.x[i] = W(x[i]);
#pragma unroll
    for (int i = 0; i < M; i++) p.u[i] = W(u[i]);
#pragma unroll
    for (int i = 0; i < 16; i++) p.mp[i] = W(P.mp[i]);
    p.dt = W(P.dt);
    p.eps = W(P.fd_eps);
  }
  /* column j < n of [fx | fu] by two full Euler steps (finite_diff_jacobian, finite_diff.h:35-47) */
  ILQR_HD static void column_full(const Point &p, int j, S *col) {
    W xa[N], fp[N], fm[N];
    fd_perturb<N, W>(p.x, j, p.eps, -1, W(0), xa);
    integrate<Model, W>(xa, p.u, p.mp, p.dt, fp);
    fd_perturb<N, W>(p.x, j, -p.eps, -1, W(0), xa);
    integrate<Model, W>(xa, p.u, p.mp, p.dt, fm);
#pragma unroll
    for (int r = 0; r < N; r++) col[r] = S((fp[r] - fm[r]) / (2 * p.eps));
  }
  /* OPT-IN (kFlagAnalyticDyn): every column of [fx | fu] of the Euler step x + dt f(x, u) from the model's closed-form
   * Jacobian; F[j * N + r] = d x'_r / d (x|u)_j.  Not the reference's arithmetic (it differs from the central
   * differences by their O(eps^2) truncation term, ~3e-8), hence a flag. */
  ILQR_HD static void jacobian_analytic(const Point &p, S *F) {
    if constexpr (HasDynamicsJac<Model>::value) {
      W A[N * N], Bm[N * M];
      Model::template dynamics_jac<W>(p.x, p.u, p.mp, A, Bm);
#pragma unroll
      for (int j = 0; j < N; j++)
#pragma unroll
        for (int r = 0; r < N; r++) F[j * N + r] = S((r == j ? W(1) : W(0)) + A[r * N + j] * p.dt);
#pragma unroll
      for (int j = 0; j < M; j++)
#pragma unroll
        for (int r = 0; r < N; r++) F[(N + j) * N + r] = S(Bm[r * M + j] * p.dt);
    }
  }
  /* column j (not a configuration variable) with the configuration-dependent part of the dynamics already formed */
  ILQR_HD static void column_shared(const Point &p, const typename Model::template Config<W> &cf, int j, S *col) {
    W xa[N], ua[M], fp[N], fm[N];
    fd_perturb<N, W>(p.x, j, p.eps, -1, W(0), xa);
    fd_perturb<M, W>(p.u, j - N, p.eps, -1, W(0), ua);
    integrate_cfg<Model, W>(cf, xa, ua, p.mp, p.dt, fp);
    fd_perturb<N, W>(p.x, j, -p.eps, -1, W(0), xa);
    fd_perturb<M, W>(p.u, j - N, -p.eps, -1, W(0), ua);
    integrate_cfg<Model, W>(cf, xa, ua, p.mp, p.dt, fm);
#pragma unroll
    for (int r = 0; r < N; r++) col[r] = S((fp[r] - fm[r]) / (2 * p.eps));
  }
};

template <class Model, typename S, int CD, class Exec>
struct Core {
  static constexpr int N = Model::N, M = Model::M, NM = N + M;
  static constexpr int G = Exec::kLanes; /* lanes cooperating on this trajectory */
  using Sc = Scratch<N, M, S, CD>;
  using Lane = LaneRegs<N, M, S>;
  static constexpr int NA = Sc::NA, NCF = Sc::NCF;
  static constexpr int NQ = NM + 1; /* row length of Scratch::Qg */
  /* column of Qg for column b of an x-indexed matrix augmented with its vector (b == n: the gradient) */
  ILQR_HD static int qcol(int b) { return b < N ? b : NM; }

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

  /* Cost derivatives of one timestep are kept in ONE array ("full layout", NCF scalars):
   *   cvec[c]     at [c]                 c < n: cx,  c >= n: cu
   *   Cfull[c][d] at [NM + c*NM + d]     the (n+m)^2 Hessian [[cxx, cxu], [cxu^T, cuu]]
   * so that every Q-function entry is the same expression  Cfull[c][d] + sum_r W[c][r] F[d][r]
   * and the warp phase that evaluates them has a single code path. */
  ILQR_HD static int ix_cxx(int i, int j) { return NM + i * NM + j; }
  ILQR_HD static int ix_cxu(int i, int j) { return NM + i * NM + N + j; }
  ILQR_HD static int ix_cux(int j, int i) { return NM + (N + j) * NM + i; }
  ILQR_HD static int ix_cuu(int i, int j) { return NM + (N + i) * NM + N + j; }

  ILQR_HD void cost_stencil(int o, bool terminal, const S *x, const S *u, S *cf) { cost_stencil_s(P, o, terminal, x, u, cf); }
  /* the same without an instance (the phase kernels of ilqr_phases.cuh run one stencil output per thread) */
  ILQR_HD static void cost_stencil_s(const SolveParams<S> &P, int o, bool terminal, const S *x_, const S *u_, S *cf) {
    typedef typename FiniteDiff<Model, S>::W W; /* evaluated in double, the result rounded to S (see FiniteDiff) */
    typename FiniteDiff<Model, S>::Point pt;
    FiniteDiff<Model, S>::load(P, x_, u_, pt);
    const W eps = pt.eps;
    const W *mp = pt.mp, *x = pt.x, *u = pt.u;
    W xa[N], ua[M];
    auto fx = [&](const W *xv, const W *uv) -> W { return terminal ? Model::final_cost(xv, mp) : Model::cost(xv, uv, mp); };
    if (o < N) { /* finite_diff_gradient wrt x, finite_diff.h:22-33 */
      fd_perturb<N, W>(x, o, eps, -1, W(0), xa);
      const W p = fx(xa, u);
      fd_perturb<N, W>(x, o, -eps, -1, W(0), xa);
      const W m = fx(xa, u);
      cf[o] = S((p - m) / (2 * eps));
      return;
    }
    o -= N;
    if (!terminal) {
      if (o < M) {
        fd_perturb<M, W>(u, o, eps, -1, W(0), ua);
        const W p = Model::cost(x, ua, mp);
        fd_perturb<M, W>(u, o, -eps, -1, W(0), ua);
        const W m = Model::cost(x, ua, mp);
        cf[N + o] = S((p - m) / (2 * eps));
        return;
      }
      o -= M;
    }
    if (o < kNxx) { /* finite_diff_hessian wrt x, finite_diff.h:67-86 */
      int i, j;
      tri_index(o, N, i, j);
      fd_perturb<N, W>(x, i, eps, j, eps, xa);
      const W pp = fx(xa, u);
      fd_perturb<N, W>(x, i, -eps, j, eps, xa);
      const W mpv = fx(xa, u);
      fd_perturb<N, W>(x, i, eps, j, -eps, xa);
      const W pm = fx(xa, u);
      fd_perturb<N, W>(x, i, -eps, j, -eps, xa);
      const W mm = fx(xa, u);
      const S v = S((pp - mpv - pm + mm) / (4 * eps * eps));
      cf[ix_cxx(i, j)] = v;
      cf[ix_cxx(j, i)] = v;
      return;
    }
    o -= kNxx;
    if (terminal) return;
    if (o < kNuu) {
      int i, j;
      tri_index(o, M, i, j);
      fd_perturb<M, W>(u, i, eps, j, eps, ua);
      const W pp = Model::cost(x, ua, mp);
      fd_perturb<M, W>(u, i, -eps, j, eps, ua);
      const W mpv = Model::cost(x, ua, mp);
      fd_perturb<M, W>(u, i, eps, j, -eps, ua);
      const W pm = Model::cost(x, ua, mp);
      fd_perturb<M, W>(u, i, -eps, j, -eps, ua);
      const W mm = Model::cost(x, ua, mp);
      const S v = S((pp - mpv - pm + mm) / (4 * eps * eps));
      cf[ix_cuu(i, j)] = v;
      cf[ix_cuu(j, i)] = v;
      return;
    }
    o -= kNuu;
    if (o < N * M) { /* calculate_cxu, src/derivatives.cpp:114-144 (its own stencil) */
      const int i = o / M, j = o % M;
      W xp[N], xm[N], up[M], um[M];
      fd_perturb<N, W>(x, i, eps, -1, W(0), xp);
      fd_perturb<N, W>(x, i, -eps, -1, W(0), xm);
      fd_perturb<M, W>(u, j, eps, -1, W(0), up);
      fd_perturb<M, W>(u, j, -eps, -1, W(0), um);
      const S v = S((Model::cost(xp, up, mp) - Model::cost(xm, up, mp) - Model::cost(xp, um, mp) + Model::cost(xm, um, mp)) /
                    (4 * (eps * eps)));
      cf[ix_cxu(i, j)] = v;
      cf[ix_cux(j, i)] = v;
    }
  }

  /* closed-form cost derivatives of the model twin in the full layout (one lane; terminal step only) */
  ILQR_HD void analytic_cost(const S *x, const S *u, bool terminal, S *cf) { analytic_cost_s(P, x, u, terminal, cf); }
  ILQR_HD static void analytic_cost_s(const SolveParams<S> &P, const S *x, const S *u, bool terminal, S *cf) {
    for (int c = 0; c < NM; c++) cf[c] = Model::cost_d1(c, x, u, P.mp, terminal);
    for (int c = 0; c < NM; c++)
      for (int d = 0; d < NM; d++) cf[NM + c * NM + d] = Model::cost_d2(c, d, x, u, P.mp, terminal);
  }

  /* get_dynamics_derivatives (+ get_cost_derivatives / get_cost_2nd_derivatives in FD mode) for the
   * whole horizon, parallel over the lanes: every column of [fx | fu] is the +eps / -eps pair of
   * finite_diff_jacobian (finite_diff.h:35-47) on the Euler step.  Two passes:
   *   A. one lane per (timestep, configuration variable): two full Euler steps (a perturbed angle changes the
   *      trigonometry, the mass matrix, everything);
   *   B. one lane per timestep for all the other variables (velocities, controls): the model's
   *      configuration-dependent part is formed once and shared by the 2 (n + m - nq) perturbed points, which
   *      then cost a few dozen operations each — same operands, same operations, same results as full steps. */
  static constexpr int kNumConfigVars = ((Model::kConfigVars >> 0) & 1) + ((Model::kConfigVars >> 1) & 1) +
                                        ((Model::kConfigVars >> 2) & 1) + ((Model::kConfigVars >> 3) & 1) +
                                        ((Model::kConfigVars >> 4) & 1) + ((Model::kConfigVars >> 5) & 1) +
                                        ((Model::kConfigVars >> 6) & 1) + ((Model::kConfigVars >> 7) & 1);
  ILQR_HD static bool is_config_var(int j) { return j < N && ((Model::kConfigVars >> j) & 1u) != 0; }
  ILQR_HD static int nth_config_var(int q) { /* index of the q-th set bit */
    int j = 0;
    for (int seen = 0; j < N; j++)
      if ((Model::kConfigVars >> j) & 1u) {
        if (seen == q) break;
        seen++;
      }
    return j;
  }

  ILQR_HD void derivative_sweep() {
    const int T = P.T;
    if ((P.flags & kFlagAnalyticDyn) && HasDynamicsJac<Model>::value) { /* opt-in: closed-form Jacobians, one lane per timestep */
      for (int base = 0; base < T; base += G) {
        ex.lanes([&](int lane, Lane &) {
          const int t = base + lane;
          if (t >= T) return;
          S x[N], u[M], Fc[NM * N];
#pragma unroll
          for (int i = 0; i < N; i++) x[i] = tr.xs[t * N + i];
#pragma unroll
          for (int i = 0; i < M; i++) u[i] = tr.us[t * M + i];
          typename FiniteDiff<Model, S>::Point pt;
          FiniteDiff<Model, S>::load(P, x, u, pt);
          FiniteDiff<Model, S>::jacobian_analytic(pt, Fc);
#pragma unroll
          for (int e = 0; e < NM * N; e++) sl.F[(size_t)t * NM * N + e] = Fc[e];
        });
      }
    } else {
    if constexpr (kNumConfigVars > 0) {
      const int n_a = T * kNumConfigVars;
      for (int base = 0; base < n_a; base += G) {
        ex.lanes([&](int lane, Lane &) {
          const int task = base + lane;
          if (task >= n_a) return;
          const int t = task / kNumConfigVars, j = nth_config_var(task - t * kNumConfigVars);
          S x[N], u[M], col[N];
#pragma unroll
          for (int i = 0; i < N; i++) x[i] = tr.xs[t * N + i];
#pragma unroll
          for (int i = 0; i < M; i++) u[i] = tr.us[t * M + i];
          typename FiniteDiff<Model, S>::Point pt;
          FiniteDiff<Model, S>::load(P, x, u, pt);
          FiniteDiff<Model, S>::column_full(pt, j, col);
#pragma unroll
          for (int r = 0; r < N; r++) sl.F[((size_t)t * NM + j) * N + r] = col[r];
        });
      }
    }
    for (int base = 0; base < T; base += G) {
      ex.lanes([&](int lane, Lane &) {
        const int t = base + lane;
        if (t >= T) return;
        S x[N], u[M], col[N];
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = tr.xs[t * N + i];
#pragma unroll
        for (int i = 0; i < M; i++) u[i] = tr.us[t * M + i];
        typename FiniteDiff<Model, S>::Point pt;
        FiniteDiff<Model, S>::load(P, x, u, pt);
        typename Model::template Config<typename FiniteDiff<Model, S>::W> cf;
        Model::configure(pt.x, pt.mp, cf);
#pragma unroll
        for (int j = 0; j < NM; j++) {
          if (is_config_var(j)) continue;
          FiniteDiff<Model, S>::column_shared(pt, cf, j, col);
#pragma unroll
          for (int r = 0; r < N; r++) sl.F[((size_t)t * NM + j) * N + r] = col[r];
        }
      });
    }
    }
    if constexpr (CD == kCostFD) {
      const int n_c = T * kStencilStep;
      for (int base = 0; base < n_c; base += G) {
        ex.lanes([&](int lane, Lane &) {
          const int task = base + lane;
          if (task >= n_c) return;
          const int t = task / kStencilStep, o = task - t * kStencilStep;
          S x[N], u[M];
#pragma unroll
          for (int i = 0; i < N; i++) x[i] = tr.xs[t * N + i];
#pragma unroll
          for (int i = 0; i < M; i++) u[i] = tr.us[t * M + i];
          cost_stencil(o, false, x, u, sl.C + (size_t)t * NCF);
        });
      }
    }
    ex.publish(); /* F is read by the backward pass's bulk copies */
  }

  /* Vx[T] = cx[T], Vxx[T] = cxx[T]  (src/ilqr_core.cpp:353-354) from sc.x = xs[T] */
  ILQR_HD void phase_terminal() {
    ex.lanes([&](int lane, Lane &) {
      if (CD == kCostFD) {
        for (int o = lane; o < kStencilTerm; o += G) cost_stencil(o, true, sc.x, sc.u, sc.Cf);
      } else if (lane == 0) {
        analytic_cost(sc.x, sc.u, true, sc.Cf);
      }
    });
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < N * NA; e += G) {
        const int r = e / NA, b = e % NA;
        sc.Va[e] = (b < N) ? sc.Cf[ix_cxx(r, b < N ? b : 0)] : sc.Cf[r];
      }
    });
  }

  /* ---- backward pass -------------------------------------------------------------------- */

  /* One timestep of the backward recursion for tile entry tt; returns false when the boxQP reports
   * failure (result < 1, src/ilqr_core.cpp:371).  The value function is kept augmented,
   * Va = [Vxx | Vx] (n x (n+1)).  Three warp phases when there is one control, four otherwise:
   *
   *   A. the Q-function (:359-367) as two products with one entry per lane each: W = F^T [Vxx' | Vx'],
   *      then Q = C + W F.  A lane needs 2n operands per entry, which it loads up front; forming a whole
   *      row of W per lane to save the exchange was measured and is slower (6n + n^2 operands under the
   *      register cap serialise on shared-memory latency).  The closed-form cost derivative of the
   *      entry is evaluated in place (Model::cost_d1/_d2).
   *   B. boxQP, gains, dV (:369-389) and the value-function update with its symmetrisation
   *      (:391-393).  With one control the boxQP is a few dozen scalar operations, so EVERY lane
   *      runs it in registers (same instructions, no divergence, nothing to broadcast) and goes
   *      straight on to its own entry of Vxx / Vx: the lane of (a, b) evaluates both Vt[a][b] and
   *      Vt[b][a] and averages them.  With several controls lane 0 solves the QP in the scratch
   *      (phase B1) and the update is a third phase.
   *
   * cfd = finite-difference cost derivatives of this timestep in the full layout, or nullptr. */
  ILQR_HD bool backward_step(int tt, S lam, const S *cfd) {
    const S *F = sc.Ft + tt * NM * N; /* F[j][r]: column j of [fx | fu] */
    const S *xt = sc.xs + tt * N;
    const S *ut = sc.us + tt * M;
    /* ---- A1: W = F^T [Vxx' | Vx'], one entry per lane.  The last column is Qx / Qu already but for the cost
     * gradient (:359-360): its lanes add that and store to Qa instead (same instructions, selected operands) ---- */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < NM * NA; e += G) {
        const int c = e / NA, b = e % NA;
        const bool is_vec = b == N;
        Acc<S> w;
#pragma unroll
        for (int q = 0; q < N; q++) w.add(F[c * N + q] * sc.Va[q * NA + b]);
        const S c1 = cfd ? cfd[c] : Model::cost_d1(c, xt, ut, P.mp, false);
        S *dst = is_vec ? &sc.Qg[c * NQ + NM] : &sc.W[e];
        *dst = is_vec ? c1 + w.v : w.v;
      }
    });
    ex.tick(1);
    /* ---- A2: Q[c][d] = C[c][d] + sum_r W[c][r] F[d][r] over the whole stacked variable (Qxx, Qux, Quu :361-363; the
     * Qxu block is computed too and never read: one instruction stream, because a divergent branch costs a lone
     * warp more than the wasted multiplies); the Quu lanes also store the regularised QuuF (:367) ---- */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < NM * NM; e += G) {
        const int c = e / NM, d = e % NM;
        Acc<S> acc;
#pragma unroll
        for (int r = 0; r < N; r++) acc.add(sc.W[c * NA + r] * F[d * N + r]);
        const S cc = cfd ? cfd[NM + e] : Model::cost_d2(c, d, xt, ut, P.mp, false);
        sc.Qg[c * NQ + d] = cc + acc.v;
        if (c >= N && d >= N) sc.qp.Q[(c - N) * M + (d - N)] = (cc + (c == d ? lam : S(0))) + acc.v;
      }
    });
    ex.tick(2);
    if constexpr (M == 1) {
      int result = 1;
      /* ---- B, one control ---- */
      ex.lanes([&](int lane, Lane &L) {
        const S Quu = sc.Qg[N * NQ + N], Qu = sc.Qg[N * NQ + NM];
        /* operands of this lane's value-function entry, loaded before the boxQP so that they are in flight behind it */
        const S *Qua = sc.Qg + N * NQ; /* [Qux | Quu | Qu] */
        const int e0 = lane < N * NA ? lane : 0, a0 = e0 / NA, b0 = e0 % NA;
        const S qa0 = Qua[a0], qb0 = Qua[qcol(b0)], qab0 = sc.Qg[a0 * NQ + qcol(b0)], qba0 = sc.Qg[(b0 < N ? b0 : 0) * NQ + a0];
        const S qg0 = Qua[lane < N ? lane : 0];
        const QPScalar<S> r = box_qp_scalar<S>(P.qp, sc.qp.Q[0], Qu, L.kprev, P.u_min[0] - ut[0], P.u_max[0] - ut[0]);
        result = r.result;
        if (r.result < 1) return;
        const S kk = r.x;
        const S nH = -r.Hinv;
        const bool fr = r.v_free != 0;
        L.dV[0] += kk * Qu;                 /* :388 */
        L.dV[1] += ((S(0.5) * kk) * Quu) * kk; /* :389, unregularised Quu */
        L.kprev = kk;                       /* warm start of the next boxQP (:369) */
        /* gains (:373-385, :396-397): entry b <= n of [K | k] is kept by the lane b */
        if (lane <= N) {
          const S g = lane < N ? (fr ? nH * qg0 : S(0)) : kk;
          S *dst = lane < N ? &sc.K[tt * N + lane] : &sc.k[tt];
          *dst = g;
        }
        /* value function (:391-393): the lane of (a, b) forms Vt[a][b] and Vt[b][a] and averages them; the Vx
         * column (b == n) runs the same instructions with k, Qu in the place of K[b], Qux[b] and copies
         * its value (0.5 * (v + v) == v exactly) */
        for (int e = lane; e < N * NA; e += G) {
          const int a = e / NA, b = e % NA;
          const bool col = b < N;
          const bool first = e == lane; /* G >= n (n + 1) lanes: the only pass, operands already loaded */
          const S qa = first ? qa0 : Qua[a], qb = first ? qb0 : Qua[qcol(b)];
          const S qab = first ? qab0 : sc.Qg[a * NQ + qcol(b)], qba = first ? qba0 : sc.Qg[(col ? b : 0) * NQ + a];
          const S Kga = fr ? nH * qa : S(0);
          const S Kgb = col ? (fr ? nH * qb : S(0)) : kk;
          const S v1 = qab + (Kga * Quu) * Kgb + Kga * qb + qa * Kgb;
          const S v2 = qba + (Kgb * Quu) * Kga + Kgb * qa + qb * Kga;
          sc.Va[e] = S(0.5) * (v1 + (col ? v2 : v1));
        }
      });
      ex.tick(3);
      return result >= 1;
    } else {
      /* ---- B1, several controls: boxQP, gains, dV on one lane; k / K go to Ka = [K | k] and to the tile ---- */
      ex.lanes([&](int lane, Lane &L) {
        if (lane != 0) return;
        QPWork<M, S> &w = sc.qp;
#pragma unroll
        for (int j = 0; j < M; j++) {
          w.c[j] = sc.Qg[(N + j) * NQ + NM];
          w.x0[j] = sc.kprev[j];
          w.lo[j] = P.u_min[j] - ut[j];
          w.hi[j] = P.u_max[j] - ut[j];
        }
        box_qp_generic<M, S>(P.qp, w);
        if (w.result < 1) return;
        for (int e = 0; e < M * NA; e++) sc.Ka[e] = 0;
#pragma unroll
        for (int j = 0; j < M; j++) sc.Ka[j * NA + N] = w.x[j];
        const int r = w.r_dim;
        int q = 0;
        for (int j = 0; j < M; j++)
          if (w.v_free[j]) w.idx[q++] = j;
        if (q > 0) {
          for (int a = 0; a < r && a < q; a++)
            for (int b = 0; b < N; b++) {
              S acc = 0;
              for (int c = 0; c < r && c < q; c++) acc += (-w.Hinv[a * r + c]) * sc.Qg[(N + w.idx[c]) * NQ + b];
              sc.Ka[w.idx[a] * NA + b] = acc;
            }
        }
        Acc<S> a0; /* :388-389, unregularised Quu */
#pragma unroll
        for (int j = 0; j < M; j++) a0.add(sc.Ka[j * NA + N] * sc.Qg[(N + j) * NQ + NM]);
        L.dV[0] += a0.v;
        Acc<S> a1;
        S row[M];
#pragma unroll
        for (int b = 0; b < M; b++) {
          Acc<S> acc;
#pragma unroll
          for (int a = 0; a < M; a++) acc.add((S(0.5) * sc.Ka[a * NA + N]) * sc.Qg[(N + a) * NQ + N + b]);
          row[b] = acc.v;
        }
#pragma unroll
        for (int b = 0; b < M; b++) a1.add(row[b] * sc.Ka[b * NA + N]);
        L.dV[1] += a1.v;
#pragma unroll
        for (int j = 0; j < M; j++) { /* :396-397, and the warm start of the next boxQP (:369) */
          sc.kprev[j] = sc.Ka[j * NA + N];
          sc.k[tt * M + j] = sc.Ka[j * NA + N];
#pragma unroll
          for (int b = 0; b < N; b++) sc.K[(tt * M + j) * N + b] = sc.Ka[j * NA + b];
        }
      });
      ex.tick(3);
      if (ex.uniform(sc.qp.result) < 1) return false;
      /* ---- B2: [Vxx | Vx] (:391-392) and the symmetrisation (:393); column n (Vx) is copied:
       * 0.5 * (v + v) == v exactly ---- */
      ex.lanes([&](int lane, Lane &) {
        for (int e = lane; e < N * NA; e += G) {
          const int a = e / NA, b = e % NA;
          S v[2];
#pragma unroll
          for (int side = 0; side < 2; side++) {
            const int aa = (side == 0 || b == N) ? a : b, bb = (side == 0 || b == N) ? b : a;
            S ktq[M]; /* row aa of K^T Quu */
#pragma unroll
            for (int j = 0; j < M; j++) {
              Acc<S> acc;
#pragma unroll
              for (int c = 0; c < M; c++) acc.add(sc.Ka[c * NA + aa] * sc.Qg[(N + c) * NQ + N + j]);
              ktq[j] = acc.v;
            }
            Acc<S> t1, t2, t3;
#pragma unroll
            for (int c = 0; c < M; c++) t1.add(ktq[c] * sc.Ka[c * NA + bb]);
#pragma unroll
            for (int c = 0; c < M; c++) t2.add(sc.Ka[c * NA + aa] * sc.Qg[(N + c) * NQ + qcol(bb)]);
#pragma unroll
            for (int c = 0; c < M; c++) t3.add(sc.Qg[(N + c) * NQ + aa] * sc.Ka[c * NA + bb]);
            v[side] = sc.Qg[aa * NQ + qcol(bb)] + t1.v + t2.v + t3.v;
          }
          sc.Va[e] = S(0.5) * (v[0] + v[1]);
        }
      });
      ex.tick(4);
      return true;
    }
  }

  /* iLQR::backward_pass.  Returns the failing timestep or 0 (:371,400). */
  ILQR_HD int backward_pass(S lam) {
    const int T = P.T;
    ex.lanes([&](int lane, Lane &L) {
      if (lane < N) sc.x[lane] = tr.xs[T * N + lane];
      if (lane < M) sc.u[lane] = 0;
      /* :369 warm start of i = T-1: the previous pass's k[T-1] */
      if constexpr (M == 1) L.kprev = tr.k[T - 1];
      else if (lane < M) sc.kprev[lane] = tr.k[(T - 1) * M + lane];
      L.dV[0] = 0; /* :356 */
      L.dV[1] = 0;
      if (lane == 0) sc.st.n_backward++;
    });
    phase_terminal();
    ex.tick(8);
    int diverged_at = -1;
    for (int ti = (T - 1) / kTileB; ti >= 0 && diverged_at < 0; ti--) {
      const int t0 = ti * kTileB;
      const int cnt = (T - t0 < kTileB) ? T - t0 : kTileB;
      ex.stage_issue(sc.Ft, sl.F + (size_t)t0 * NM * N, cnt * NM * N, P.bulk_f != 0);
      if constexpr (CD == kCostFD) ex.stage_issue(sc.Ct, sl.C + (size_t)t0 * NCF, cnt * NCF, P.bulk_c != 0);
      ex.lanes([&](int lane, Lane &) {
        for (int e = lane; e < cnt * N; e += G) sc.xs[e] = tr.xs[t0 * N + e];
        for (int e = lane; e < cnt * M; e += G) sc.us[e] = tr.us[t0 * M + e];
      });
      ex.stage_wait();
      ex.tick(6);
      int first_done = 0; /* tile entries [first_done, cnt) hold finished k / K */
      for (int tt = cnt - 1; tt >= 0; tt--) {
        const S *cfd = (CD == kCostFD) ? sc.Ct + tt * NCF : nullptr;
        if (!backward_step(tt, lam, cfd)) { /* the steps above this one have already written their k, K (:396-397) */
          diverged_at = t0 + tt;
          first_done = tt + 1;
          break;
        }
      }
      /* flush the tile's k / K and the gradient-norm terms of its timesteps (:405-412), one per lane */
      ex.lanes([&](int lane, Lane &) {
        for (int e = lane + first_done * M * N; e < cnt * M * N; e += G) tr.K[t0 * M * N + e] = sc.K[e];
        for (int e = lane + first_done * M; e < cnt * M; e += G) tr.k[t0 * M + e] = sc.k[e];
        for (int e = lane + first_done; e < cnt; e += G) sl.gterm[t0 + e] = gn_term(sc.k + e * M, sc.us + e * M);
      });
      ex.tick(7);
    }
    ex.lanes([&](int lane, Lane &L) {
      if (lane != 0) return;
      sc.st.dV0 = L.dV[0];
      sc.st.dV1 = L.dV[1];
    });
    pass_complete = diverged_at < 0;
    if (diverged_at >= 0) return diverged_at;
    /* Vx[0], Vxx[0] are results of record for the tests (include/ilqr.h:76-77) */
    ex.lanes([&](int lane, Lane &) {
      for (int e = lane; e < N * NA; e += G) {
        const int r = e / NA, b = e % NA;
        if (b < N) tr.Vxx0[r * N + b] = sc.Va[e];
        else tr.Vx0[r] = sc.Va[e];
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

  /* one step of iLQR::forward_pass (:314-326) for one lane: x, cost are the lane's running state (kept in
   * registers by the callers' tile loops), uc receives the applied control */
  ILQR_HD void rollout_step(S *x, S &cost, S *uc, const S *xhat, const S *ubar, const S *kt, const S *Kt, S alpha, int mode) {
#pragma unroll
    for (int j = 0; j < M; j++) {
      S v = ubar[j];
      if (mode == kRollClosed) v = ubar[j] + kt[j] * alpha; /* :188-190 */
      if (mode != kRollOpen) {                              /* :316 */
        Acc<S> a;
#pragma unroll
        for (int i = 0; i < N; i++) a.add(Kt[j * N + i] * (x[i] - xhat[i]));
        v += a.v;
      }
      if (P.flags & kFlagClampRollout) v = clampd(v, P.u_min[j], P.u_max[j]); /* opt-in: "the right way", :327-329 */
      uc[j] = v;
    }
    cost += Model::cost(x, uc, P.mp); /* :324 */
    S x1[N];
    integrate<Model, S>(x, uc, P.mp, P.dt, x1); /* :325 */
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = x1[i];
  }

  ILQR_HD void load_forward_tile(int t0, int cnt, int mode) {
    ex.lanes([&](int lane, Lane &) {
      /* all of the tile's loads in flight before the first store to the scratch */
      constexpr int U = (kTile * N + G - 1) / G, UK = (kTile * M * N + G - 1) / G, U1 = (kTile * M + G - 1) / G;
      S vx[U], vu[U1], vK[UK], vk[U1];
#pragma unroll
      for (int u = 0; u < U; u++) vx[u] = (lane + u * G < cnt * N) ? tr.xs[t0 * N + lane + u * G] : S(0);
#pragma unroll
      for (int u = 0; u < U1; u++) vu[u] = (lane + u * G < cnt * M) ? tr.us[t0 * M + lane + u * G] : S(0);
      if (mode != kRollOpen) {
#pragma unroll
        for (int u = 0; u < UK; u++) vK[u] = (lane + u * G < cnt * M * N) ? tr.K[t0 * M * N + lane + u * G] : S(0);
#pragma unroll
        for (int u = 0; u < U1; u++) vk[u] = (lane + u * G < cnt * M) ? tr.k[t0 * M + lane + u * G] : S(0);
      }
#pragma unroll
      for (int u = 0; u < U; u++)
        if (lane + u * G < cnt * N) sc.xs[lane + u * G] = vx[u];
#pragma unroll
      for (int u = 0; u < U1; u++)
        if (lane + u * G < cnt * M) sc.us[lane + u * G] = vu[u];
      if (mode != kRollOpen) {
#pragma unroll
        for (int u = 0; u < UK; u++)
          if (lane + u * G < cnt * M * N) sc.K[lane + u * G] = vK[u];
#pragma unroll
        for (int u = 0; u < U1; u++)
          if (lane + u * G < cnt * M) sc.k[lane + u * G] = vk[u];
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
      ex.tick(10);
      load_forward_tile(t0, cnt, kRollClosed);
      ex.tick(14);
      ex.lanes([&](int lane, Lane &L) {
        if (lane >= na) return;
        const S alpha = P.alpha[lane];
        S *cx = sl.cand_x + ((size_t)lane * T + t0) * N;
        S *cu = sl.cand_u + ((size_t)lane * T + t0) * M;
        S x[N], uc[M], cost = L.cost; /* the tile runs on register copies of the lane state */
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = L.x[i];
        for (int tt = 0; tt < cnt; tt++) {
          rollout_step(x, cost, uc, sc.xs + tt * N, sc.us + tt * M, sc.k + tt * M, sc.K + tt * M * N, alpha, kRollClosed);
#pragma unroll
          for (int j = 0; j < M; j++) cu[tt * M + j] = uc[j];
#pragma unroll
          for (int i = 0; i < N; i++) cx[tt * N + i] = x[i];
        }
#pragma unroll
        for (int i = 0; i < N; i++) L.x[i] = x[i];
        L.cost = cost;
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
      copy_lanes<8>(tr.xs + N, sl.cand_x + (size_t)a * T * N, T * N, lane);
      copy_lanes<8>(tr.us, sl.cand_u + (size_t)a * T * M, T * M, lane);
    });
  }
  /* dst[e] = src[e] for this lane's share e = lane, lane + G, ... with U loads in flight at a time.  Written out
   * because a plain loop cannot be pipelined by the compiler (dst might alias src, so every load waits for the store
   * before it): the commit of a 200-step trajectory took 25 full L2 round trips, 10 k cycles. */
  template <int U>
  ILQR_HD static void copy_lanes(S *dst, const S *src, int count, int lane) {
    int e = lane;
    for (; e + (U - 1) * G < count; e += U * G) {
      S v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = ld_fresh(src + e + u * G);
#pragma unroll
      for (int u = 0; u < U; u++) dst[e + u * G] = v[u];
    }
    for (; e < count; e += G) dst[e] = ld_fresh(src + e);
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
        S x[N], uc[M], cost = L.cost;
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = L.x[i];
        for (int tt = 0; tt < cnt; tt++) {
#pragma unroll
          for (int i = 0; i < N; i++) tr.xs[(t0 + tt) * N + i] = x[i];
          rollout_step(x, cost, uc, sc.xs + tt * N, sc.us + tt * M, sc.k + tt * M, sc.K + tt * M * N, alpha, mode);
#pragma unroll
          for (int j = 0; j < M; j++) tr.us[(t0 + tt) * M + j] = uc[j];
        }
#pragma unroll
        for (int i = 0; i < N; i++) L.x[i] = x[i];
        L.cost = cost;
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
    if (d == 0 && pass_complete) gradient_norm_from_terms();
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

  /* The loop of iLQR::generate_trajectory() (:103-288), up to n_iters trips, as three calls so that a warp
   * carrying two trajectories can advance both one trip at a time in lockstep:
   *   iterate_begin(n);  while (iterate_trip()) {}  iterate_end();                                      */
  int trips_left = 0;
  /* the last backward pass ran every timestep.  A pass that stops at timestep 0 returns 0 like a success (:371 vs :142)
   * but has not written this pass's gradient-norm term of timestep 0: the norm is then formed from k and us as they
   * stand (the stale k[0] with the current us[0], like the reference's get_gradient_norm) */
  bool pass_complete = false;
  bool have_derivs = false; /* the lane group's F / C buffers hold this trajectory's current derivatives */

  ILQR_HD void iterate_begin(int n_iters) {
    load_state();
    trips_left = n_iters;
    have_derivs = false;
  }
  ILQR_HD void iterate_end() {
    ex.lanes([&](int lane, Lane &) {
      if (lane != 0) return;
      if (sc.st.status == kRunning && sc.st.iter >= P.max_iter) sc.st.status = kExitMaxIter;
    });
    store_state();
  }
  /* One trip of the loop body, in three parts so that a warp carrying two trajectories can run the line searches of
   * both at once (ilqr_kernel.cuh):  trip_pre — derivatives and backward pass, up to the gradient test;
   * rollout_candidates — the line search's rollouts, only after kTripRoll;  trip_post — acceptance, commit, lambda
   * schedule, returns whether another trip follows.  iterate_trip() is the three in sequence. */
  enum { kTripStop = 0, kTripRoll = 1, kTripNoRoll = 2 };
  ILQR_HD int trip_pre() {
    {
      const int it = ex.uniform(sc.st.iter), status = ex.uniform(sc.st.status);
      if (!(it < P.max_iter && trips_left > 0 && status == kRunning)) return kTripStop;
    }
    trips_left--;
    /* :115-120 */
    if (ex.uniform(sc.st.flg_change) || !have_derivs) {
      ex.tick(15);
      derivative_sweep();
      have_derivs = true;
      ex.tick(0);
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
        if (ex.uniform(sc.flag)) break;
        continue;
      }
      back_done = true;
    }
    if (back_done && pass_complete) gradient_norm_from_terms();
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
    ex.tick(9);
    if (ex.uniform(sc.flag) == 2) return kTripStop; /* gradient exit: `break` before iter++ */
    return back_done ? kTripRoll : kTripNoRoll;
  }
  ILQR_HD bool trip_post(int pre) {
    const bool back_done = pre == kTripRoll;
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
    const bool fwd_done = ex.uniform(sc.flag) == 1;
    ex.tick(11);
    if (fwd_done) commit_candidate(ex.uniform(sc.st.alpha_index));
    ex.tick(12);
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
    ex.tick(13);
    return ex.uniform(sc.flag) == 0;
  }
  ILQR_HD bool iterate_trip() {
    const int pre = trip_pre();
    if (pre == kTripStop) return false;
    if (pre == kTripRoll) rollout_candidates();
    ex.tick(10);
    return trip_post(pre);
  }
  ILQR_HD void op_iterate(int n_iters) {
    iterate_begin(n_iters);
    while (iterate_trip()) {
    }
    iterate_end();
  }

  ILQR_HD static S fmax_(S a, S b) { return a < b ? b : a; } /* std::max(a, b) */
  ILQR_HD static S fmin_(S a, S b) { return b < a ? b : a; } /* std::min(a, b) */
};

}  // namespace ilqr
#endif
