/*
 * ilqr_phases.cuh — the batch-lockstep engine: one loop trip of src/ilqr_core.cpp:103-288 for EVERY running
 * trajectory of the batch as five kernels, each with the thread mapping that fills its lanes.
 *
 * The warp-per-trajectory kernel (ilqr_kernel.cuh) gives a trajectory 32 lanes for every phase of a trip, but the
 * phases do not have 32-way work: the line search has n_alpha = 11 rollouts, the backward step (n+m)(n+1) = 25
 * four-term dot products and then a scalar boxQP that every lane repeats, so two thirds of the fp64 lane slots
 * it issues carry nothing (profiles/r1n: 129 k warp instructions per trip, 20.8 lanes active, boxQP 19 % of them
 * redundant).  Here the unit of parallelism changes with the phase instead:
 *
 *   sweep     one thread per (trajectory, timestep, variable group): central differences of the Euler step
 *             (src/derivatives.cpp:15-26) and, in FD-cost mode, one cost-stencil output per thread (:29-144);
 *   backward  (large active sets; a few thousand run phase_backward_rows_kernel, 8 lanes per trajectory, below)
 *             one thread per trajectory: the whole recursion (:350-401) in registers — Va = [Vxx | Vx], the
 *             Jacobian of the step, the Q-function — with boxQP (src/boxqp.cpp) inline; 32 trajectories per warp,
 *             every fp64 instruction 32 useful lanes, no shared memory, no barriers.  The next timestep's
 *             Jacobian / state / control are loaded while the current step computes;
 *   rollout   one thread per (trajectory, alpha): the n_alpha candidates of the line search (:184-226, :305-337),
 *             adjacent lanes share the trajectory's nominal arrays (broadcast loads), each streams its candidate
 *             to the trajectory's candidate buffer;
 *   accept    one warp per trajectory: lane 0 applies the reference's serial acceptance order and the lambda
 *             schedule (:199-282), all lanes commit the accepted candidate by a copy of whole timesteps, and the
 *             trajectories that go on are flagged;
 *   compact   one CTA: the next trip's active list = the flagged entries of this one, in order (the lists stay
 *             ascending, so the trajectories of a warp stay neighbours in memory).
 *
 * The batch advances one trip per round of launches; finished trajectories leave the active list, so every launch
 * covers exactly the running ones (ragged trip counts cost nothing but the per-trip latency floor).  Arithmetic is
 * the warp kernel's, entry for entry, in the same order (the helpers are Core's static functions), so both engines
 * return identical bits; the per-thread functions are plain ILQR_HD code and tests/emu runs them on the CPU against
 * the oracle.
 *
 * HBM (per handle, all per TRAJECTORY): F [B][T][n+m][n] Jacobian columns, C [B][T][NCF] (FD-cost mode),
 * cand_x [B][T][n_alpha][n], cand_u [B][T][n_alpha][m] (candidate-interleaved: the n_alpha stores of one timestep are
 * contiguous — with every candidate's block contiguous instead, each store instruction touched eleven DRAM pages and
 * configs[4] lost a quarter of its rate), newcost [B][16], act [5][B] active lists and flags (PhaseBufs).
 */
#ifndef ILQR_PHASES_CUH_
#define ILQR_PHASES_CUH_

#include "ilqr_core.cuh"
#if defined(__CUDACC__)
#include "ilqr_kernel.cuh"
#endif

namespace ilqr {

struct NoExec {
  static constexpr int kLanes = 1;
};

enum { kRollStop = 0, kRollGo = 1, kRollSkip = 2 }; /* TrajState::roll */

/* One thread per trajectory means the lanes of a warp take different branches inside the boxQP of a timestep.  They
 * must come back together before the next timestep's (identical) arithmetic: left to the compiler, the loop with its
 * early exits reconverged only at the end of the kernel and every lane ran the whole recursion alone (ncu: 1.16
 * threads per instruction, 5.4 ms per backward launch instead of 0.14).  `mask` = the lanes that carry a trajectory. */
ILQR_HD void reconverge(unsigned mask) {
#if defined(__CUDA_ARCH__)
  __syncwarp(mask);
#else
  (void)mask;
#endif
}
ILQR_HD bool lanes_any(unsigned mask, bool p) {
#if defined(__CUDA_ARCH__)
  return __any_sync(mask, p) != 0;
#else
  (void)mask;
  return p;
#endif
}

#ifndef ILQR_ROLLOUT_PREFETCH
#define ILQR_ROLLOUT_PREFETCH 4
#endif
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

/* dst[0..CNT) = src[0..CNT).  `src` points into a dense array of CNT-element runs whose base is 256-byte aligned, so
 * when a run is a multiple of 16 bytes every run is 16-byte aligned and moves as 128-bit accesses. */
template <int CNT>
ILQR_HD void load_run(double *dst, const double *src) {
#if defined(__CUDA_ARCH__)
  if constexpr (CNT % 2 == 0) {
#pragma unroll
    for (int i = 0; i < CNT / 2; i++) {
      const double2 v = reinterpret_cast<const double2 *>(src)[i];
      dst[2 * i] = v.x;
      dst[2 * i + 1] = v.y;
    }
  } else
#endif
  {
#pragma unroll
    for (int i = 0; i < CNT; i++) dst[i] = src[i];
  }
}
template <int CNT>
ILQR_HD void load_run(float *dst, const float *src) {
#if defined(__CUDA_ARCH__)
  if constexpr (CNT % 4 == 0) {
#pragma unroll
    for (int i = 0; i < CNT / 4; i++) {
      const float4 v = reinterpret_cast<const float4 *>(src)[i];
      dst[4 * i] = v.x;
      dst[4 * i + 1] = v.y;
      dst[4 * i + 2] = v.z;
      dst[4 * i + 3] = v.w;
    }
  } else
#endif
  {
#pragma unroll
    for (int i = 0; i < CNT; i++) dst[i] = src[i];
  }
}
template <int CNT>
ILQR_HD void store_run(double *dst, const double *src) {
#if defined(__CUDA_ARCH__)
  if constexpr (CNT % 2 == 0) {
#pragma unroll
    for (int i = 0; i < CNT / 2; i++) reinterpret_cast<double2 *>(dst)[i] = make_double2(src[2 * i], src[2 * i + 1]);
  } else
#endif
  {
#pragma unroll
    for (int i = 0; i < CNT; i++) dst[i] = src[i];
  }
}
template <int CNT>
ILQR_HD void store_run(float *dst, const float *src) {
#if defined(__CUDA_ARCH__)
  if constexpr (CNT % 4 == 0) {
#pragma unroll
    for (int i = 0; i < CNT / 4; i++)
      reinterpret_cast<float4 *>(dst)[i] = make_float4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
  } else
#endif
  {
#pragma unroll
    for (int i = 0; i < CNT; i++) dst[i] = src[i];
  }
}

/* per-trajectory work arrays of the phase engine (base pointers of the whole batch) */
template <typename S>
struct PhaseBufs {
  S *F;       /* [B][T][n+m][n]      */
  S *C;       /* [B][T][NCF]         FD-cost mode only */
  S *cand_x;  /* [B][T][n_alpha][n]  candidate states x_1..x_T (allocated on first use: not needed in re-roll mode) */
  S *cand_u;  /* [B][T][n_alpha][m]  candidate controls */
  S *newcost; /* [B][kMaxAlpha]      */
  S *gterm;   /* [B][T]              per-timestep terms of the gradient norm, written by the backward phase */
  int *act;   /* [5][B]              active lists: two double-buffered by trip parity; [2] = stage 2 of a staged line
                                     search; [3] = accepted steps to re-roll (phase_commit_kernel); [4] = per list
                                     position, "runs another trip" (phase_compact_kernel) */
  int *n_act; /* [4]                 their lengths */
};

template <class Model, typename S, int CD>
struct Phases {
  static constexpr int N = Model::N, M = Model::M, NM = N + M, NA = N + 1;
  using CoreT = Core<Model, S, CD, NoExec>; /* static helpers only: stencils, index maps, gn_term */
  static constexpr int NCF = NM + NM * NM;
  static constexpr int kParts = CoreT::kNumConfigVars + 1; /* sweep tasks per timestep */
  static constexpr int kStencilStep = CoreT::kStencilStep;

  /* ---- sweep ---------------------------------------------------------------------------------------------- */

  /* One task of get_dynamics_derivatives (src/derivatives.cpp:15-26, finite_diff.h:35-47) at timestep t:
   * part < kNumConfigVars: the column of that configuration variable (two full Euler steps);
   * part == kNumConfigVars: every other column, the configuration-dependent part of the dynamics formed once
   * (Core::derivative_sweep, passes A and B). */
  ILQR_HD static void sweep_task(const SolveParams<S> &P, const S *xs, const S *us, S *F, int part, int t) {
    S x[N], u[M], col[N];
    load_run<N>(x, xs + (size_t)t * N);
    load_run<M>(u, us + (size_t)t * M);
    typedef FiniteDiff<Model, S> FD; /* differences formed in double, rounded to S (ilqr_core.cuh) */
    typename FD::Point pt;
    FD::load(P, x, u, pt);
    if ((P.flags & kFlagAnalyticDyn) && HasDynamicsJac<Model>::value) { /* opt-in: closed-form Jacobians, all columns at once */
      if (part != 0) return;
      S Fc[NM * N];
      FD::jacobian_analytic(pt, Fc);
#pragma unroll
      for (int j = 0; j < NM; j++) store_run<N>(F + ((size_t)t * NM + j) * N, Fc + j * N);
      return;
    }
    if (part < CoreT::kNumConfigVars) {
      const int j = CoreT::nth_config_var(part);
      FD::column_full(pt, j, col);
      store_run<N>(F + ((size_t)t * NM + j) * N, col);
      return;
    }
    typename Model::template Config<typename FD::W> cf;
    Model::configure(pt.x, pt.mp, cf);
#pragma unroll
    for (int j = 0; j < NM; j++) {
      if (CoreT::is_config_var(j)) continue;
      FD::column_shared(pt, cf, j, col);
      store_run<N>(F + ((size_t)t * NM + j) * N, col);
    }
  }
  /* one output of the finite-difference cost stencils (src/derivatives.cpp:29-144) at timestep t */
  ILQR_HD static void stencil_task(const SolveParams<S> &P, const S *xs, const S *us, S *C, int o, int t) {
    S x[N], u[M];
    load_run<N>(x, xs + (size_t)t * N);
    load_run<M>(u, us + (size_t)t * M);
    CoreT::cost_stencil_s(P, o, false, x, u, C + (size_t)t * NCF);
  }

  /* ---- backward ------------------------------------------------------------------------------------------- */

  /* iLQR::backward_pass (src/ilqr_core.cpp:350-401) for one trajectory, by ONE thread, everything in registers.
   * Returns the failing timestep or 0 (:371,400).  Entry for entry the arithmetic of Core::backward_step. */
  ILQR_HD static int backward_pass(const SolveParams<S> &P, const TrajPtrs<S> &tr, const S *F, const S *Cfd, S *gterm, S lam,
                                   S &dV0, S &dV1, bool &complete, unsigned mask, bool enabled) {
    const int T = P.T;
    /* every lane of `mask` walks all T timesteps and meets the others at the top of each; a lane that is not
     * `enabled` (its trajectory needs no further pass) or whose boxQP has failed skips the bodies */
    int failed_at = enabled ? -1 : T;
    const S *mp = P.mp;
    S Va[N][NA]; /* [Vxx | Vx] at i+1, then at i */
    {            /* Vx[T] = cx[T], Vxx[T] = cxx[T]  (:353-354) */
      S xT[N], uz[M], Cf[NCF];
      load_run<N>(xT, tr.xs + (size_t)T * N);
#pragma unroll
      for (int j = 0; j < M; j++) uz[j] = 0;
      if constexpr (CD == kCostFD) {
        for (int o = 0; o < CoreT::kStencilTerm; o++) CoreT::cost_stencil_s(P, o, true, xT, uz, Cf);
      } else {
        CoreT::analytic_cost_s(P, xT, uz, true, Cf);
      }
#pragma unroll
      for (int r = 0; r < N; r++)
#pragma unroll
        for (int b = 0; b < NA; b++) Va[r][b] = (b < N) ? Cf[CoreT::ix_cxx(r, b < N ? b : 0)] : Cf[r];
    }
    S kprev[M]; /* :369 warm start of i = T-1: the previous pass's k[T-1] */
#pragma unroll
    for (int j = 0; j < M; j++) kprev[j] = tr.k[(size_t)(T - 1) * M + j];
    dV0 = 0; /* :356 */
    dV1 = 0;
    S Fn[NM * N], xn[N], un[M]; /* the next step's operands, in flight while this one computes */
    load_run<NM * N>(Fn, F + (size_t)(T - 1) * NM * N);
    load_run<N>(xn, tr.xs + (size_t)(T - 1) * N);
    load_run<M>(un, tr.us + (size_t)(T - 1) * M);
    for (int t = T - 1; t >= 0; t--) {
      reconverge(mask);
      if (failed_at >= 0) continue;
      S Fm[NM * N], xt[N], ut[M]; /* Fm[c * N + q]: column c of [fx | fu] */
#pragma unroll
      for (int e = 0; e < NM * N; e++) Fm[e] = Fn[e];
#pragma unroll
      for (int e = 0; e < N; e++) xt[e] = xn[e];
#pragma unroll
      for (int e = 0; e < M; e++) ut[e] = un[e];
      if (t > 0) {
        load_run<NM * N>(Fn, F + (size_t)(t - 1) * NM * N);
        load_run<N>(xn, tr.xs + (size_t)(t - 1) * N);
        load_run<M>(un, tr.us + (size_t)(t - 1) * M);
      }
      S cfd[CD == kCostFD ? NCF : 1];
      if constexpr (CD == kCostFD) load_run<NCF>(cfd, Cfd + (size_t)t * NCF);
      /* W = F^T [Vxx' | Vx']; its last column is Qx / Qu but for the cost gradient (:359-360) */
      S W[NM][N], Qv[NM];
#pragma unroll
      for (int c = 0; c < NM; c++) {
#pragma unroll
        for (int b = 0; b < NA; b++) {
          S w = Fm[c * N] * Va[0][b];
#pragma unroll
          for (int q = 1; q < N; q++) w = w + Fm[c * N + q] * Va[q][b];
          if (b < N) {
            W[c][b] = w;
          } else {
            S c1;
          if constexpr (CD == kCostFD) c1 = cfd[c];
          else c1 = Model::cost_d1(c, xt, ut, mp, false);
            Qv[c] = c1 + w;
          }
        }
      }
      /* Q[c][d] = C[c][d] + sum_r W[c][r] F[d][r]  (Qxx, Qux, Quu :361-363; the Qxu block is never read), and the
       * regularised QuuF (:367) */
      S Q[NM][NM], QuuF[M * M];
#pragma unroll
      for (int c = 0; c < NM; c++) {
#pragma unroll
        for (int d = 0; d < NM; d++) {
          if (c < N && d >= N) continue;
          S acc = W[c][0] * Fm[d * N];
#pragma unroll
          for (int r = 1; r < N; r++) acc = acc + W[c][r] * Fm[d * N + r];
          S cc;
          if constexpr (CD == kCostFD) cc = cfd[NM + c * NM + d];
          else cc = Model::cost_d2(c, d, xt, ut, mp, false);
          Q[c][d] = cc + acc;
          if (c >= N && d >= N) QuuF[(c - N) * M + (d - N)] = (cc + (c == d ? lam : S(0))) + acc;
        }
      }
      S Vt[N][NA]; /* the unsymmetrised value function of this step (:391-392) */
      if constexpr (M == 1) {
        const S Quu = Q[N][N], Qu = Qv[N];
        const QPScalar<S> r = box_qp_scalar<S>(P.qp, QuuF[0], Qu, kprev[0], P.u_min[0] - ut[0], P.u_max[0] - ut[0]);
        if (r.result < 1) { /* :371 */
          failed_at = t;
          continue;
        }
        const S kk = r.x;
        const S nH = -r.Hinv;
        const bool fr = r.v_free != 0;
        dV0 += kk * Qu;                     /* :388 */
        dV1 += ((S(0.5) * kk) * Quu) * kk;  /* :389, unregularised Quu */
        kprev[0] = kk;
        S Kg[N]; /* gains (:373-385) */
#pragma unroll
        for (int b = 0; b < N; b++) Kg[b] = fr ? nH * Q[N][b] : S(0);
        store_run<N>(tr.K + (size_t)t * N, Kg); /* :396-397 */
        tr.k[t] = kk;
        gterm[t] = CoreT::gn_term(&kk, ut); /* :405-412, summed in ascending order after the pass */
#pragma unroll
        for (int a = 0; a < N; a++) {
          const S qa = Q[N][a];
#pragma unroll
          for (int b = 0; b < NA; b++) {
            const bool col = b < N;
            const S Kgb = col ? Kg[col ? b : 0] : kk;
            const S qb = col ? Q[N][col ? b : 0] : Qu;
            const S qab = col ? Q[a][col ? b : 0] : Qv[a];
            Vt[a][b] = qab + (Kg[a] * Quu) * Kgb + Kg[a] * qb + qa * Kgb;
          }
        }
      } else {
        QPWork<M, S> w;
#pragma unroll
        for (int e = 0; e < M * M; e++) w.Q[e] = QuuF[e];
#pragma unroll
        for (int j = 0; j < M; j++) {
          w.c[j] = Qv[N + j];
          w.x0[j] = kprev[j];
          w.lo[j] = P.u_min[j] - ut[j];
          w.hi[j] = P.u_max[j] - ut[j];
        }
        box_qp_generic<M, S>(P.qp, w);
        if (w.result < 1) {
          failed_at = t;
          continue;
        }
        S Ka[M][NA]; /* [K | k] */
#pragma unroll
        for (int j = 0; j < M; j++)
#pragma unroll
          for (int b = 0; b < NA; b++) Ka[j][b] = 0;
#pragma unroll
        for (int j = 0; j < M; j++) Ka[j][N] = w.x[j];
        const int rd = w.r_dim;
        int nf = 0;
        for (int j = 0; j < M; j++)
          if (w.v_free[j]) w.idx[nf++] = j;
        if (nf > 0) {
          /* K on the free dimensions: -R^-1 R^-T Qux[free] (:376-385); Ka is indexed dynamically through idx */
          S Kfree[M][N];
          for (int a = 0; a < rd && a < nf; a++)
            for (int b = 0; b < N; b++) {
              S acc = 0;
              for (int c = 0; c < rd && c < nf; c++) {
                S qux = 0;
#pragma unroll
                for (int jj = 0; jj < M; jj++)
                  if (jj == w.idx[c]) qux = Q[N + jj][b];
                acc += (-w.Hinv[a * rd + c]) * qux;
              }
              Kfree[a][b] = acc;
            }
#pragma unroll
          for (int j = 0; j < M; j++)
            for (int a = 0; a < rd && a < nf; a++)
              if (w.idx[a] == j) {
#pragma unroll
                for (int b = 0; b < N; b++) Ka[j][b] = Kfree[a][b];
              }
        }
        {
          Acc<S> a0; /* :388-389, unregularised Quu */
#pragma unroll
          for (int j = 0; j < M; j++) a0.add(Ka[j][N] * Qv[N + j]);
          dV0 += a0.v;
          Acc<S> a1;
          S row[M];
#pragma unroll
          for (int b = 0; b < M; b++) {
            Acc<S> acc;
#pragma unroll
            for (int a = 0; a < M; a++) acc.add((S(0.5) * Ka[a][N]) * Q[N + a][N + b]);
            row[b] = acc.v;
          }
#pragma unroll
          for (int b = 0; b < M; b++) a1.add(row[b] * Ka[b][N]);
          dV1 += a1.v;
        }
#pragma unroll
        for (int j = 0; j < M; j++) { /* :396-397, and the warm start of the next boxQP (:369) */
          kprev[j] = Ka[j][N];
          tr.k[(size_t)t * M + j] = Ka[j][N];
          store_run<N>(tr.K + ((size_t)t * M + j) * N, Ka[j]);
        }
        gterm[t] = CoreT::gn_term(kprev, ut);
        /* [Vxx | Vx] (:391-392) */
        S ktq[N][M]; /* row aa of K^T Quu */
#pragma unroll
        for (int aa = 0; aa < N; aa++)
#pragma unroll
          for (int j = 0; j < M; j++) {
            Acc<S> acc;
#pragma unroll
            for (int c = 0; c < M; c++) acc.add(Ka[c][aa] * Q[N + c][N + j]);
            ktq[aa][j] = acc.v;
          }
#pragma unroll
        for (int aa = 0; aa < N; aa++)
#pragma unroll
          for (int bb = 0; bb < NA; bb++) {
            Acc<S> t1, t2, t3;
#pragma unroll
            for (int c = 0; c < M; c++) t1.add(ktq[aa][c] * Ka[c][bb]);
#pragma unroll
            for (int c = 0; c < M; c++) t2.add(Ka[c][aa] * (bb < N ? Q[N + c][bb < N ? bb : 0] : Qv[N + c]));
#pragma unroll
            for (int c = 0; c < M; c++) t3.add(Q[N + c][aa] * Ka[c][bb]);
            Vt[aa][bb] = (bb < N ? Q[aa][bb < N ? bb : 0] : Qv[aa]) + t1.v + t2.v + t3.v;
          }
      }
      /* the symmetrisation (:393); the Vx column is copied: 0.5 * (v + v) */
#pragma unroll
      for (int a = 0; a < N; a++)
#pragma unroll
        for (int b = 0; b < NA; b++) Va[a][b] = S(0.5) * (Vt[a][b] + (b < N ? Vt[b < N ? b : 0][a] : Vt[a][b]));
    }
    reconverge(mask);
    complete = failed_at < 0; /* a failure at timestep 0 returns 0 like a success (:371 vs :142) but is not complete */
    if (failed_at >= 0) return failed_at < T ? failed_at : 0;
    /* Vx[0], Vxx[0] are results of record for the tests (include/ilqr.h:76-77) */
#pragma unroll
    for (int r = 0; r < N; r++) {
      tr.Vx0[r] = Va[r][N];
#pragma unroll
      for (int b = 0; b < N; b++) tr.Vxx0[r * N + b] = Va[r][b];
    }
    return 0;
  }

  /* get_gradient_norm (:405-412): mean_t max_j |k_tj| / (|u_tj| + 1), ascending t, from k and us in global memory.
   * After a pass that stopped at timestep d the entries at and below d are the previous pass's, as in the
   * reference (its k is only overwritten down to the failing step). */
  /* the same from the terms a COMPLETE pass left in gterm (same operands, same operations as gn_term on k / us) */
  ILQR_HD static S gradient_norm_terms(const SolveParams<S> &P, const S *gterm) {
    const int T = P.T;
    S acc = 0;
    int t = 0;
    for (; t + 4 <= T; t += 4) {
      S g[4];
#pragma unroll
      for (int q = 0; q < 4; q++) g[q] = gterm[t + q];
#pragma unroll
      for (int q = 0; q < 4; q++) acc += g[q];
    }
    for (; t < T; t++) acc += gterm[t];
    return acc / T;
  }
  ILQR_HD static S gradient_norm(const SolveParams<S> &P, const TrajPtrs<S> &tr) {
    const int T = P.T;
    S acc = 0;
    int t = 0;
    for (; t + 4 <= T; t += 4) { /* four timesteps' loads and divisions in flight; the sum stays in ascending order */
      S g[4];
#pragma unroll
      for (int q = 0; q < 4; q++) g[q] = CoreT::gn_term(tr.k + (size_t)(t + q) * M, tr.us + (size_t)(t + q) * M);
#pragma unroll
      for (int q = 0; q < 4; q++) acc += g[q];
    }
    for (; t < T; t++) acc += CoreT::gn_term(tr.k + (size_t)t * M, tr.us + (size_t)t * M);
    return acc / T;
  }

  /* The head of a loop trip for one running trajectory (:115-159): bookkeeping of the derivative refresh (done by
   * the sweep phase just before), backward pass with the lambda retries, gradient-norm exit.  Leaves in s.roll what
   * the line search has to do.  s lives in registers / local memory; the caller stores it back. */
  ILQR_HD static void backward_trip(const SolveParams<S> &P, const TrajPtrs<S> &tr, const S *F, const S *Cfd, S *gterm,
                                    TrajState<S> &s, unsigned mask) {
    s.trips++;
    if (s.flg_change) {
      s.flg_change = 0;
      s.n_deriv++;
    }
    bool back_done = false, need = true, complete = false;
    while (lanes_any(mask, need)) { /* :136-150; the lanes of a warp pass through together (see reconverge) */
      S dV0 = 0, dV1 = 0;
      bool done_all = false;
      const int diverge = backward_pass(P, tr, F, Cfd, gterm, s.lam, dV0, dV1, done_all, mask, need);
      if (!need) continue;
      complete = done_all;
      s.n_backward++;
      s.dV0 = dV0;
      s.dV1 = dV1;
      s.diverge = diverge;
      if (diverge != 0) {
        s.dlam = CoreT::fmax_(s.dlam * P.lambda_factor, P.lambda_factor);
        s.lam = CoreT::fmax_(s.lam * s.dlam, P.lambda_min);
        if (s.lam > P.lambda_max) need = false;
        continue;
      }
      back_done = true;
      need = false;
    }
    /* a pass that stopped early (or at timestep 0, which the reference cannot tell from success) left stale entries:
     * then the norm is formed from k and us as they stand, like the reference's */
    s.gnorm = (back_done && complete) ? gradient_norm_terms(P, gterm) : gradient_norm(P, tr);
    if (s.gnorm < P.tol_grad && s.lam < P.grad_lambda_gate) { /* :153-159: `break` before iter++ */
      s.status = kExitGrad;
      s.roll = kRollStop;
      return;
    }
    s.roll = back_done ? kRollGo : kRollSkip;
  }

  /* ---- rollout -------------------------------------------------------------------------------------------- */

  /* iLQR::forward_pass (:305-337) for candidate `a` of the line search (:188-197): u_t = us_t + alpha k_t +
   * K_t (x_t - xs_t), unclamped; controls and states stream to the trajectory's candidate buffer, the cost is
   * returned.  Where the rollout goes (MODE):
   *   kToCand    to the trajectory's candidate buffer (cand_x / cand_u = its [T][n_alpha][.] block);
   *   kCostOnly  nowhere — only the cost is wanted;
   *   kInPlace   over xs[1..T] / us themselves: the commit of an accepted candidate by re-rolling it (the nominal
   *              x_{t+1}, u_{t+1} are in registers before the step overwrites them).
   * Small active sets keep every candidate (kToCand) and commit by a copy: a round of launches costs its latency
   * floor and a second rollout would add to it.  Large ones are throughput- and HBM-bound — the candidates were 40 %
   * of all DRAM traffic of configs[4] and ten of eleven are never read — so they run kCostOnly and re-roll the one
   * accepted candidate (one extra rollout in eleven; same operations, same bits). */
  enum { kToCand = 0, kCostOnly = 1, kInPlace = 2 };
  template <int MODE>
  ILQR_HD static S rollout_task(const SolveParams<S> &P, const TrajPtrs<S> &tr, S *cand_x, S *cand_u, int a, int cand_w = -1) {
    const int T = P.T;
    if (cand_w < 0) cand_w = P.n_alpha; /* candidates kept per timestep (the first stage of a staged line search keeps fewer) */
    const S alpha = P.alpha[a];
    const S *mp = P.mp;
    S x[N], cost = 0;
    load_run<N>(x, tr.x0);
    S xh_n[N], ub_n[M], kt_n[M], Kt_n[M * N];
    load_run<N>(xh_n, tr.xs);
    load_run<M>(ub_n, tr.us);
    load_run<M>(kt_n, tr.k);
    load_run<M * N>(Kt_n, tr.K);
    for (int t = 0; t < T; t++) {
      S xh[N], ub[M], kt[M], Kt[M * N];
#pragma unroll
      for (int e = 0; e < N; e++) xh[e] = xh_n[e];
#pragma unroll
      for (int e = 0; e < M; e++) ub[e] = ub_n[e];
#pragma unroll
      for (int e = 0; e < M; e++) kt[e] = kt_n[e];
#pragma unroll
      for (int e = 0; e < M * N; e++) Kt[e] = Kt_n[e];
      if (t + 1 < T) {
        load_run<N>(xh_n, tr.xs + (size_t)(t + 1) * N);
        load_run<M>(ub_n, tr.us + (size_t)(t + 1) * M);
        load_run<M>(kt_n, tr.k + (size_t)(t + 1) * M);
        load_run<M * N>(Kt_n, tr.K + (size_t)(t + 1) * M * N);
      }
#if defined(__CUDA_ARCH__) && ILQR_ROLLOUT_PREFETCH > 0
      /* the nominal arrays of a few steps ahead, into L2: with a full machine the one-step register prefetch above
       * does not cover a DRAM round trip (ncu, configs[4] shard: 37 % of this kernel's stall samples waited here) */
      if (t + ILQR_ROLLOUT_PREFETCH < T) {
        prefetch_l2(tr.xs + (size_t)(t + ILQR_ROLLOUT_PREFETCH) * N);
        prefetch_l2(tr.K + (size_t)(t + ILQR_ROLLOUT_PREFETCH) * M * N);
        if ((t & 3) == 0) {
          prefetch_l2(tr.us + (size_t)(t + ILQR_ROLLOUT_PREFETCH) * M);
          prefetch_l2(tr.k + (size_t)(t + ILQR_ROLLOUT_PREFETCH) * M);
        }
      }
#endif
      S uc[M];
#pragma unroll
      for (int j = 0; j < M; j++) {
        S v = ub[j] + kt[j] * alpha; /* :188-190 */
        Acc<S> acc;                  /* :316 */
#pragma unroll
        for (int i = 0; i < N; i++) acc.add(Kt[j * N + i] * (x[i] - xh[i]));
        v += acc.v;
        if (P.flags & kFlagClampRollout) v = clampd(v, P.u_min[j], P.u_max[j]); /* opt-in: "the right way", :327-329 */
        uc[j] = v;
      }
      cost += Model::cost(x, uc, mp); /* :324 */
      S x1[N];
      integrate<Model, S>(x, uc, mp, P.dt, x1); /* :325 */
#pragma unroll
      for (int i = 0; i < N; i++) x[i] = x1[i];
      if constexpr (MODE == kToCand) {
        store_run<M>(cand_u + ((size_t)t * cand_w + a) * M, uc);
        store_run<N>(cand_x + ((size_t)t * cand_w + a) * N, x);
      } else if constexpr (MODE == kInPlace) {
        store_run<M>(tr.us + (size_t)t * M, uc);
        store_run<N>(tr.xs + (size_t)(t + 1) * N, x);
      }
    }
    cost += Model::final_cost(x, mp); /* :335 */
    return cost;
  }

  /* ---- accept --------------------------------------------------------------------------------------------- */

  /* the acceptance test (:199-213) in the reference's serial order over the candidates' costs; true = a step was
   * accepted (s.alpha_index says which) */
  ILQR_HD static bool accept(const SolveParams<S> &P, TrajState<S> &s, const S *newcost, int n_try = -1) {
    const bool back_done = s.roll == kRollGo;
    if (n_try < 0) n_try = P.n_alpha; /* a staged line search (phase_accept_kernel) looks at the first few only */
    bool fwd_done = false;
    s.alpha_index = -1;
    S alpha = 0;
    if (back_done) {
      for (int a = 0; a < n_try; a++) {
        alpha = P.alpha[a];
        s.new_cost = newcost[a];
        s.n_rollouts++;
        s.dcost = s.cost - s.new_cost;
        s.expected = -alpha * (s.dV0 + alpha * s.dV1);
        S z;
        if (s.expected > 0) z = s.dcost / s.expected;
        else z = S((S(0) < s.dcost) - (s.dcost < S(0))); /* sgn, include/common.h:43-44 */
        if (z > P.z_min) {
          s.alpha_index = a;
          fwd_done = true;
          break;
        }
      }
      if (!fwd_done) alpha = 0;
    }
    s.alpha = alpha;
    return fwd_done;
  }
  /* lambda schedule and termination (:242-282); true = the trajectory goes on to another trip */
  ILQR_HD static bool schedule(const SolveParams<S> &P, TrajState<S> &s, bool fwd_done) {
    bool stop = false;
    if (fwd_done) { /* :242-263 */
      s.dlam = CoreT::fmin_(s.dlam / P.lambda_factor, 1 / P.lambda_factor);
      s.lam = s.lam * s.dlam * S(s.lam > P.lambda_min);
      s.cost = s.new_cost;
      s.flg_change = 1;
      s.n_accept++;
      if (s.dcost < P.tol_fun) {
        s.status = kExitTolFun;
        stop = true;
      }
    } else { /* :264-282 */
      s.dlam = CoreT::fmax_(s.dlam * P.lambda_factor, P.lambda_factor);
      s.lam = CoreT::fmax_(s.lam * s.dlam, P.lambda_min);
      s.n_reject++;
      if (s.lam > P.lambda_max) {
        s.status = kExitLambdaMax;
        stop = true;
      }
    }
    if (!stop) s.iter++;
    if (!stop && s.iter >= P.max_iter) { /* the loop counter ran out (:103) */
      s.status = kExitMaxIter;
      stop = true;
    }
    return !stop;
  }
};

#if defined(__CUDACC__)

template <typename S>
struct PArgs {
  SolveParams<S> P;
  const S *x0;
  S *xs, *us, *K, *k, *Vx0, *Vxx0;
  TrajState<S> *st;
  PhaseBufs<S> buf;
  long long B;
  int parity;      /* which active list this trip reads */
  int force_sweep; /* first trip of an ilqr_iterate call: F / C may be stale (set_initial, warm start, test hooks) */
  int reroll;      /* this trip's line search keeps no candidates: the accepted one is re-rolled (phase_commit_kernel) */
  /* Staged line search (large active sets): stage 1 rolls out candidates [0, cand_hi) of every trajectory, keeps them
   * ([T][cand_hi][.]: with four, a timestep's states are one 128-byte line) and accepts where one of them passes;
   * the trajectories where none does go to list 2, and stage 2 rolls out [cand_lo, n_alpha) for them, cost only.  A
   * step accepted there (one line search in ten) goes to list 3 and is re-rolled over xs / us (phase_commit_kernel).
   * The reference tries the candidates in order and stops at the first that passes (:186-214), so the candidates
   * after it were never needed: 62 % of the line searches of the synthetic batch end within four, 28 % reject all.
   * stage 0 = all candidates at once (small active sets: latency, not work, is what a round costs). */
  int stage, cand_lo, cand_hi;
  int ordered; /* the next active list is built in the order of this one (phase_compact_kernel), not by atomic append */
  int keep; /* this stage's rollouts were stored (cand_hi per timestep): an accepted one is copied, otherwise re-rolled */
};

template <typename S>
__device__ __forceinline__ TrajPtrs<S> phase_pointers(const PArgs<S> &a, long long b, int N, int M) {
  const size_t T = (size_t)a.P.T;
  TrajPtrs<S> tr;
  tr.x0 = a.x0 + b * N;
  tr.xs = a.xs + b * (T + 1) * N;
  tr.us = a.us + b * T * M;
  tr.K = a.K + b * T * M * N;
  tr.k = a.k + b * T * M;
  tr.Vx0 = a.Vx0 + b * N;
  tr.Vxx0 = a.Vxx0 + b * N * N;
  tr.st = a.st + b;
  return tr;
}

/* start of an ilqr_iterate call: the running trajectories, in index order within a warp */
template <typename S>
__global__ void phase_begin_kernel(const __grid_constant__ PArgs<S> a) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  bool run = false;
  if (b < a.B) {
    TrajState<S> &s = a.st[b];
    if (s.status == kRunning && s.iter >= a.P.max_iter) s.status = kExitMaxIter;
    run = s.status == kRunning;
  }
  const unsigned m = __ballot_sync(0xffffffffu, run);
  if (!m) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(&a.buf.n_act[a.parity], __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (run) a.buf.act[(size_t)a.parity * a.B + base + __popc(m & ((1u << lane) - 1))] = (int)b;
}

constexpr int kSweepThreads = 128;
template <class Model, typename S, int CD>
__global__ void __launch_bounds__(kSweepThreads) phase_sweep_kernel(const __grid_constant__ PArgs<S> a) {
  using Ph = Phases<Model, S, CD>;
  constexpr int N = Model::N, M = Model::M, NM = N + M;
  const int n_act = a.buf.n_act[a.parity];
  const int T = a.P.T;
  for (int i = blockIdx.x; i < n_act; i += gridDim.x) {
    const long long b = a.buf.act[(size_t)a.parity * a.B + i];
    if (!(a.st[b].flg_change || a.force_sweep)) continue; /* :115-120 */
    const S *xs = a.xs + b * (size_t)(T + 1) * N;
    const S *us = a.us + b * (size_t)T * M;
    S *F = a.buf.F + b * (size_t)T * NM * N;
    /* part-major so that the lanes of a warp run the same code path */
    for (int task = threadIdx.x; task < T * Ph::kParts; task += kSweepThreads) {
      const int part = task / T, t = task - part * T;
      Ph::sweep_task(a.P, xs, us, F, part, t);
    }
    if constexpr (CD == kCostFD) {
      S *C = a.buf.C + b * (size_t)T * Ph::NCF;
      for (int task = threadIdx.x; task < T * Ph::kStencilStep; task += kSweepThreads) {
        const int o = task / T, t = task - o * T;
        Ph::stencil_task(a.P, xs, us, C, o, t);
      }
    }
  }
}

constexpr int kBackwardThreads = 32;
#ifndef ILQR_BACKWARD_MINB
#define ILQR_BACKWARD_MINB 1
#endif
template <class Model, typename S, int CD>
__global__ void __launch_bounds__(kBackwardThreads, ILQR_BACKWARD_MINB) phase_backward_kernel(const __grid_constant__ PArgs<S> a) {
  using Ph = Phases<Model, S, CD>;
  constexpr int N = Model::N, M = Model::M, NM = N + M;
  const int n_act = a.buf.n_act[a.parity];
  const int i = blockIdx.x * kBackwardThreads + threadIdx.x;
  if (i == 0) a.buf.n_act[a.parity ^ 1] = a.buf.n_act[2] = a.buf.n_act[3] = 0; /* the lists this trip's accept phase fills */
  const unsigned mask = __ballot_sync(0xffffffffu, i < n_act);
  if (i >= n_act) return;
  const long long b = a.buf.act[(size_t)a.parity * a.B + i];
  const TrajPtrs<S> tr = phase_pointers(a, b, N, M);
  const S *F = a.buf.F + b * (size_t)a.P.T * NM * N;
  const S *Cfd = CD == kCostFD ? a.buf.C + b * (size_t)a.P.T * Ph::NCF : nullptr;
  TrajState<S> s = *tr.st;
  Ph::backward_trip(a.P, tr, F, Cfd, a.buf.gterm + b * (size_t)a.P.T, s, mask);
  *tr.st = s;
}

/* Model::cost_d1 / cost_d2 for a row index known only at run time (the row a lane owns): the models' functions take
 * the index as a plain argument, so this is just a call; kept as a named hook for models that specialise it. */
template <class Model, typename S>
__device__ __forceinline__ S cost_d1_rt(int c, const S *x, const S *u, const S *mp) { return Model::cost_d1(c, x, u, mp, false); }
template <class Model, typename S>
__device__ __forceinline__ S cost_d2_rt(int c, int d, const S *x, const S *u, const S *mp) { return Model::cost_d2(c, d, x, u, mp, false); }

/* The backward phase for MID-SIZED active sets: kRowLanes = 8 lanes per trajectory, lane c owning ROW c of the stacked
 * variable (x | u) for the whole pass.  With its row of F and a copy of [Vxx' | Vx'] a lane forms its row of
 * W = F^T [Vxx' | Vx'] and then its row of the Q-function entirely in registers (no exchange between the two
 * products); the lane of the first control row runs the boxQP on its own entries and publishes [K | k] and the
 * control rows of Q; the state-row lanes then form their rows of the value function, swap the off-diagonal entries
 * through shared memory for the symmetrisation, and publish the new [Vxx | Vx].  Three warp barriers per timestep,
 * ~50 scalars of shared memory per trajectory, four trajectories per warp advancing in lockstep.  The chain per timestep is a quarter of
 * the one-thread-per-trajectory kernel's and the instruction count per trajectory a fifth of the 32-lane
 * decomposition's, which is what a batch of a few thousand trajectories (BASELINE configs[1]) needs: too few to
 * fill the machine one thread each, too many to give each a warp.  Same expressions, same bits. */
#if defined(ILQR_ROWS_CLOCKS)
static __device__ unsigned long long g_rows_clk[4] = {0, 0, 0, ~0ULL};
#endif
constexpr int kRowLanes = 8;
constexpr int kRowThreads = 128;
template <class Model, typename S, int CD>
struct RowScratch {
  static constexpr int N = Model::N, M = Model::M, NM = N + M, NA = N + 1, NCF = NM + NM * NM;
  alignas(16) S Va[N * NA];        /* [Vxx | Vx] of the step above */
  alignas(16) S Vt[N * N];         /* unsymmetrised Vxx of this step, for the transposed read */
  alignas(16) S Uq[M * (NM + 1)];  /* the control rows of the Q-function: [Qux | Quu | Qu] */
  alignas(16) S Kk[M * NA];        /* [K | k] */
  S Cf[NCF];                       /* terminal cost derivatives, full layout */
  S QuuF[M * M];                   /* regularised Quu */
  S lam;
  int fail; /* failing timestep of the current pass, or -1 */
  int need; /* another pass is wanted (lambda retry, :142-148) */
};

/* GPW = trajectories (lane groups) per warp: 4 fills the lanes; fewer leaves lanes idle but serialises fewer boxQP
 * lanes per warp and spreads a small active set over more schedulers */
template <class Model, typename S, int CD, int GPW>
__global__ void __launch_bounds__(kRowThreads, 2) phase_backward_rows_kernel(const __grid_constant__ PArgs<S> a) {
  using Ph = Phases<Model, S, CD>;
  using CoreT = typename Ph::CoreT;
  constexpr int N = Model::N, M = Model::M, NM = N + M, NA = N + 1, NCF = NM + NM * NM, G = kRowLanes;
  static_assert(NM <= G, "one row of the stacked variable per lane");
  constexpr int kPerCta = (kRowThreads / 32) * GPW;
  __shared__ RowScratch<Model, S, CD> scr[kPerCta];
  const int wl = threadIdx.x & 31;
  const int grp = (threadIdx.x >> 5) * GPW + wl / G, lane = wl % G;
  const int n_act = a.buf.n_act[a.parity];
  const int i = (wl / G < GPW) ? blockIdx.x * kPerCta + grp : 0x7fffffff;
  if (blockIdx.x == 0 && threadIdx.x == 0) a.buf.n_act[a.parity ^ 1] = a.buf.n_act[2] = a.buf.n_act[3] = 0; /* the lists this trip's accept phase fills */
  /* The four trajectories of a warp advance in lockstep: every barrier below is warp-wide (over the lanes that carry a
   * trajectory), so the row products and the value update of the four groups issue as ONE instruction stream with
   * 4 x (n + m) lanes active, and only the boxQP lanes ever run apart.  (With group-wide barriers the groups drifted
   * and the warp issued four streams in turn: ncu 15 threads per instruction, 3.6 k cycles per timestep.) */
  const unsigned gmask = __ballot_sync(0xffffffffu, i < n_act);
  if (i >= n_act) return;
#if defined(ILQR_ROWS_CLOCKS)
  const long long rows_t0 = clock64();
#endif
  const long long b = a.buf.act[(size_t)a.parity * a.B + i];
  const TrajPtrs<S> tr = phase_pointers(a, b, N, M);
  const SolveParams<S> &P = a.P;
  const int T = P.T;
  const S *mp = P.mp;
  const S *F = a.buf.F + b * (size_t)T * NM * N;
  const S *Cfd = CD == kCostFD ? a.buf.C + b * (size_t)T * NCF : nullptr;
  RowScratch<Model, S, CD> &sc = scr[grp];
  S *gterm = a.buf.gterm + b * (size_t)T;
  const bool qp_lane = lane == N; /* first control row: boxQP, scalars, state record */
  const int c = lane < NM ? lane : 0; /* this lane's row (lanes >= n + m idle along on row 0 and store nothing) */
  const bool has_row = lane < NM;
  TrajState<S> s;
  if (qp_lane) {
    s = *tr.st;
    s.trips++;
    if (s.flg_change) {
      s.flg_change = 0;
      s.n_deriv++;
    }
    sc.lam = s.lam;
    sc.need = 1;
  }
  bool back_done = false, complete = false;
  __syncwarp(gmask);
  while (__any_sync(gmask, sc.need != 0)) { /* :136-150; a group that needs no further pass idles through it */
    const bool enabled = sc.need != 0;
    const S lam = sc.lam;
    /* Vx[T] = cx[T], Vxx[T] = cxx[T]  (:353-354) */
    {
      S xT[N], uz[M];
      load_run<N>(xT, tr.xs + (size_t)T * N);
#pragma unroll
      for (int j = 0; j < M; j++) uz[j] = 0;
      if constexpr (CD == kCostFD) {
        for (int o = lane; o < CoreT::kStencilTerm; o += G) CoreT::cost_stencil_s(P, o, true, xT, uz, sc.Cf);
      } else {
        if (lane == 0) CoreT::analytic_cost_s(P, xT, uz, true, sc.Cf);
      }
    }
    __syncwarp(gmask);
    for (int e = lane; e < N * NA; e += G) {
      const int r = e / NA, bb = e % NA;
      sc.Va[e] = (bb < N) ? sc.Cf[CoreT::ix_cxx(r, bb < N ? bb : 0)] : sc.Cf[r];
    }
    if (qp_lane) sc.fail = enabled ? -1 : T;
    S kprev[M]; /* :369 warm start of i = T-1: the previous pass's k[T-1] */
#pragma unroll
    for (int j = 0; j < M; j++) kprev[j] = tr.k[(size_t)(T - 1) * M + j];
    S dV0 = 0, dV1 = 0; /* :356 */
    S Vrow[NA];         /* this lane's row of [Vxx | Vx] after the last step */
#pragma unroll
    for (int e = 0; e < NA; e++) Vrow[e] = 0;
    S Fn[NM * N], xn[N], un[M];
    load_run<NM * N>(Fn, F + (size_t)(T - 1) * NM * N);
    load_run<N>(xn, tr.xs + (size_t)(T - 1) * N);
    load_run<M>(un, tr.us + (size_t)(T - 1) * M);
    __syncwarp(gmask);
    for (int t = T - 1; t >= 0; t--) {
      const bool live = sc.fail < 0; /* uniform in the group: written before a barrier every lane has passed */
      S Fm[NM * N], xt[N], ut[M];
#pragma unroll
      for (int e = 0; e < NM * N; e++) Fm[e] = Fn[e];
#pragma unroll
      for (int e = 0; e < N; e++) xt[e] = xn[e];
#pragma unroll
      for (int e = 0; e < M; e++) ut[e] = un[e];
      if (t > 0 && live) {
        load_run<NM * N>(Fn, F + (size_t)(t - 1) * NM * N);
        load_run<N>(xn, tr.xs + (size_t)(t - 1) * N);
        load_run<M>(un, tr.us + (size_t)(t - 1) * M);
      }
      S cvec = 0, crow[NM]; /* this row's finite-difference cost derivatives */
      if constexpr (CD == kCostFD) {
        cvec = Cfd[(size_t)t * NCF + c];
#pragma unroll
        for (int d = 0; d < NM; d++) crow[d] = Cfd[(size_t)t * NCF + NM + c * NM + d];
      }
      S Va[N][NA];
#pragma unroll
      for (int q = 0; q < N; q++) load_run<NA>(Va[q], sc.Va + q * NA);
      /* this lane's row of F (selected once: c is fixed for the kernel, the select chain is cheap next to the products) */
      S Fc[N];
#pragma unroll
      for (int q = 0; q < N; q++) {
        S v = Fm[q];
#pragma unroll
        for (int cc = 1; cc < NM; cc++) v = (c == cc) ? Fm[cc * N + q] : v;
        Fc[q] = v;
      }
      /* row c of W = F^T [Vxx' | Vx'] and of the Q-function (:359-367) */
      S W[N], Qv, Q[NM];
#pragma unroll
      for (int bb = 0; bb < NA; bb++) {
        S w = Fc[0] * Va[0][bb];
#pragma unroll
        for (int q = 1; q < N; q++) w = w + Fc[q] * Va[q][bb];
        if (bb < N) W[bb < N ? bb : 0] = w;
        else {
          S c1;
          if constexpr (CD == kCostFD) c1 = cvec;
          else c1 = cost_d1_rt<Model, S>(c, xt, ut, mp);
          Qv = c1 + w;
        }
      }
      S QuuFrow[M];
#pragma unroll
      for (int d = 0; d < NM; d++) {
        S acc = W[0] * Fm[d * N];
#pragma unroll
        for (int r = 1; r < N; r++) acc = acc + W[r] * Fm[d * N + r];
        S cc;
        if constexpr (CD == kCostFD) cc = crow[d];
        else cc = cost_d2_rt<Model, S>(c, d, xt, ut, mp);
        Q[d] = cc + acc;
        if (d >= N) QuuFrow[d - N] = (cc + (c == d ? lam : S(0))) + acc;
      }
      /* the control rows go to shared memory; with one control its lane solves the QP on its own registers first */
      if (live && has_row && c >= N) {
        const int j = c - N;
#pragma unroll
        for (int d = 0; d < NM; d++) sc.Uq[j * (NM + 1) + d] = Q[d];
        sc.Uq[j * (NM + 1) + NM] = Qv;
#pragma unroll
        for (int d = 0; d < M; d++) sc.QuuF[j * M + d] = QuuFrow[d];
      }
      if constexpr (M == 1) {
        /* The QP's three inputs go from the control row's lane to every lane of its group, and every lane solves it
         * (same instructions on the same operands; the warm start kprev is tracked by every lane): the four groups of
         * the warp then run the common path of the solver as one instruction stream instead of four lone lanes. */
        const int src = (wl & ~(G - 1)) + N;
        const S QuuF_b = __shfl_sync(gmask, QuuFrow[0], src);
        const S Qu_b = __shfl_sync(gmask, Qv, src);
        QPScalar<S> r;
        r.result = 1;
        r.x = 0;
        r.Hinv = 0;
        r.v_free = 0;
        if (live) r = box_qp_scalar<S>(P.qp, QuuF_b, Qu_b, kprev[0], P.u_min[0] - ut[0], P.u_max[0] - ut[0]);
        if (live && r.result >= 1) kprev[0] = r.x;
        if (live && qp_lane) {
          const S Quu = Q[N], Qu = Qv;
          if (r.result < 1) {
            sc.fail = t; /* :371 */
          } else {
            const S kk = r.x;
            const S nH = -r.Hinv;
            const bool fr = r.v_free != 0;
            dV0 += kk * Qu;                    /* :388 */
            dV1 += ((S(0.5) * kk) * Quu) * kk; /* :389, unregularised Quu */
            gterm[t] = CoreT::gn_term(&kk, ut);
            S Kg[NA];
#pragma unroll
            for (int bb = 0; bb < N; bb++) Kg[bb] = fr ? nH * Q[bb] : S(0);
            Kg[N] = kk;
#pragma unroll
            for (int bb = 0; bb < NA; bb++) sc.Kk[bb] = Kg[bb];
            store_run<N>(tr.K + (size_t)t * N, Kg); /* :396-397 */
            tr.k[t] = kk;
          }
        }
      } else {
        __syncwarp(gmask);
        if (live && qp_lane) {
          QPWork<M, S> w;
#pragma unroll
          for (int e = 0; e < M * M; e++) w.Q[e] = sc.QuuF[e];
#pragma unroll
          for (int j = 0; j < M; j++) {
            w.c[j] = sc.Uq[j * (NM + 1) + NM];
            w.x0[j] = kprev[j];
            w.lo[j] = P.u_min[j] - ut[j];
            w.hi[j] = P.u_max[j] - ut[j];
          }
          box_qp_generic<M, S>(P.qp, w);
          if (w.result < 1) {
            sc.fail = t;
          } else {
            S Ka[M][NA];
#pragma unroll
            for (int j = 0; j < M; j++)
#pragma unroll
              for (int bb = 0; bb < NA; bb++) Ka[j][bb] = 0;
#pragma unroll
            for (int j = 0; j < M; j++) Ka[j][N] = w.x[j];
            const int rd = w.r_dim;
            int nf = 0;
            for (int j = 0; j < M; j++)
              if (w.v_free[j]) w.idx[nf++] = j;
            if (nf > 0) { /* K on the free dimensions: -R^-1 R^-T Qux[free] (:376-385) */
              for (int aa = 0; aa < rd && aa < nf; aa++)
                for (int bb = 0; bb < N; bb++) {
                  S acc = 0;
                  for (int cc = 0; cc < rd && cc < nf; cc++) acc += (-w.Hinv[aa * rd + cc]) * sc.Uq[w.idx[cc] * (NM + 1) + bb];
#pragma unroll
                  for (int j = 0; j < M; j++)
                    if (w.idx[aa] == j) Ka[j][bb] = acc;
                }
            }
            {
              Acc<S> a0; /* :388-389, unregularised Quu */
#pragma unroll
              for (int j = 0; j < M; j++) a0.add(Ka[j][N] * sc.Uq[j * (NM + 1) + NM]);
              dV0 += a0.v;
              Acc<S> a1;
              S row[M];
#pragma unroll
              for (int bb = 0; bb < M; bb++) {
                Acc<S> acc;
#pragma unroll
                for (int aa = 0; aa < M; aa++) acc.add((S(0.5) * Ka[aa][N]) * sc.Uq[aa * (NM + 1) + N + bb]);
                row[bb] = acc.v;
              }
#pragma unroll
              for (int bb = 0; bb < M; bb++) a1.add(row[bb] * Ka[bb][N]);
              dV1 += a1.v;
            }
#pragma unroll
            for (int j = 0; j < M; j++) { /* :396-397, and the warm start of the next boxQP (:369) */
              kprev[j] = Ka[j][N];
              tr.k[(size_t)t * M + j] = Ka[j][N];
              store_run<N>(tr.K + ((size_t)t * M + j) * N, Ka[j]);
#pragma unroll
              for (int bb = 0; bb < NA; bb++) sc.Kk[j * NA + bb] = Ka[j][bb];
            }
            gterm[t] = CoreT::gn_term(kprev, ut);
          }
        }
      }
      __syncwarp(gmask); /* S1: [K | k], the control rows of Q and the verdict of the QP are published */
      const bool live2 = live && sc.fail < 0;
      /* row a = c of the unsymmetrised value function (:391-392), by the state-row lanes */
      S Vt[NA];
      {
        S Ka[M][NA], Uq[M][NM + 1];
#pragma unroll
        for (int j = 0; j < M; j++) {
          load_run<NA>(Ka[j], sc.Kk + j * NA);
          load_run<NM + 1>(Uq[j], sc.Uq + j * (NM + 1));
        }
        S Kca[M], Uca[M]; /* K[.][a] and Qux[.][a] for this lane's a (a select chain: a is fixed for the kernel) */
#pragma unroll
        for (int j = 0; j < M; j++) {
          S kv = Ka[j][0], uv = Uq[j][0];
#pragma unroll
          for (int q = 1; q < N; q++) {
            kv = (c == q) ? Ka[j][q] : kv;
            uv = (c == q) ? Uq[j][q] : uv;
          }
          Kca[j] = kv;
          Uca[j] = uv;
        }
        S ktq[M]; /* row a of K^T Quu */
#pragma unroll
        for (int j = 0; j < M; j++) {
          Acc<S> acc;
#pragma unroll
          for (int cc = 0; cc < M; cc++) acc.add(Kca[cc] * Uq[cc][N + j]);
          ktq[j] = acc.v;
        }
#pragma unroll
        for (int bb = 0; bb < NA; bb++) {
          Acc<S> t1, t2, t3;
#pragma unroll
          for (int cc = 0; cc < M; cc++) t1.add(ktq[cc] * Ka[cc][bb]);
#pragma unroll
          for (int cc = 0; cc < M; cc++) t2.add(Kca[cc] * (bb < N ? Uq[cc][bb < N ? bb : 0] : Uq[cc][NM]));
#pragma unroll
          for (int cc = 0; cc < M; cc++) t3.add(Uca[cc] * Ka[cc][bb]);
          Vt[bb] = (bb < N ? Q[bb < N ? bb : 0] : Qv) + t1.v + t2.v + t3.v;
        }
      }
      if (live2 && lane < N) {
#pragma unroll
        for (int bb = 0; bb < N; bb++) sc.Vt[lane * N + bb] = Vt[bb];
      }
      __syncwarp(gmask); /* S2 */
      if (live2 && lane < N) { /* the symmetrisation (:393); the Vx column is copied: 0.5 * (v + v) */
#pragma unroll
        for (int bb = 0; bb < NA; bb++) {
          const S other = bb < N ? sc.Vt[(bb < N ? bb : 0) * N + lane] : Vt[bb];
          Vrow[bb] = S(0.5) * (Vt[bb] + other);
        }
#pragma unroll
        for (int bb = 0; bb < NA; bb++) sc.Va[lane * NA + bb] = Vrow[bb];
      }
      __syncwarp(gmask); /* S3: [Vxx | Vx] of this step is published */
    }
    const int failed_at = sc.fail;
    if (enabled && failed_at < 0 && lane < N) { /* Vx[0], Vxx[0] are results of record for the tests (include/ilqr.h:76-77) */
      tr.Vx0[lane] = Vrow[N];
#pragma unroll
      for (int bb = 0; bb < N; bb++) tr.Vxx0[lane * N + bb] = Vrow[bb];
    }
    __syncwarp(gmask); /* every lane has read sc.fail / sc.lam / sc.need before the QP lane rewrites them */
    if (enabled && qp_lane) {
      const int diverge = failed_at >= 0 ? failed_at : 0;
      complete = failed_at < 0;
      s.n_backward++;
      s.dV0 = dV0;
      s.dV1 = dV1;
      s.diverge = diverge;
      int need = 0;
      if (diverge != 0) {
        s.dlam = CoreT::fmax_(s.dlam * P.lambda_factor, P.lambda_factor);
        s.lam = CoreT::fmax_(s.lam * s.dlam, P.lambda_min);
        need = !(s.lam > P.lambda_max);
      } else {
        back_done = true;
      }
      sc.lam = s.lam;
      sc.need = need;
    }
    __syncwarp(gmask);
  }
#if defined(ILQR_ROWS_CLOCKS) /* experiment builds: spread of the per-warp durations of this launch */
  if (wl == 0) {
    const unsigned long long dt = (unsigned long long)(clock64() - rows_t0);
    atomicAdd(&g_rows_clk[0], dt);
    atomicMax(&g_rows_clk[1], dt);
    atomicMin(&g_rows_clk[3], dt);
    const unsigned long long done = atomicAdd(&g_rows_clk[2], 1ULL) + 1;
    const unsigned long long total = ((unsigned long long)n_act + kPerCta - 1) / kPerCta * (kRowThreads / 32) -
                                     ((unsigned long long)(((n_act + kPerCta - 1) / kPerCta) * kPerCta - n_act) / GPW);
    if (done == total) {
      printf("rows launch: %llu warps, mean %.0f cycles, min %llu, max %llu (n_act %d)\n", done, (double)g_rows_clk[0] / done,
             g_rows_clk[3], g_rows_clk[1], n_act);
      g_rows_clk[0] = g_rows_clk[1] = g_rows_clk[2] = 0;
      g_rows_clk[3] = ~0ULL;
    }
  }
#endif
  if (qp_lane) {
    __threadfence_block(); /* K / k were written by this lane; the reads below are its own */
    s.gnorm = (back_done && complete) ? Ph::gradient_norm_terms(P, gterm) : Ph::gradient_norm(P, tr);
    if (s.gnorm < P.tol_grad && s.lam < P.grad_lambda_gate) { /* :153-159 */
      s.status = kExitGrad;
      s.roll = kRollStop;
    } else {
      s.roll = back_done ? kRollGo : kRollSkip;
    }
    *tr.st = s;
  }
}

/* The head of a trip for SMALL active sets: one warp per trajectory runs Core::trip_pre — derivative sweep (when the
 * trajectory changed), backward pass, gradient test — exactly as the persistent warp kernel does (ilqr_kernel.cuh),
 * against the trajectory's own F / C arrays.  One thread per trajectory (phase_backward_kernel) has the lowest cost
 * per trajectory but the longest chain per timestep (one lane issues the whole step: ~3 us); with a few thousand
 * trajectories in flight the machine is far from full and the 32-lane decomposition (~0.7 us per step) finishes the
 * phase sooner.  Same arithmetic, same bits; the host picks per trip from the size of the active list. */
template <class Model, typename S, int CD>
__global__ void __launch_bounds__(kThreads, ILQR_MIN_BLOCKS) phase_pre_warp_kernel(const __grid_constant__ PArgs<S> a) {
  constexpr int N = Model::N, M = Model::M, NM = N + M;
  using Ex = WarpExec<N, M, S, 32>;
  using CoreT = Core<Model, S, CD, Ex>;
  using Sc = typename CoreT::Sc;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t per_group = warp_smem_bytes<Sc, S>(a.P.T);
  const int group = threadIdx.x >> 5;
  unsigned char *mine = smem_raw + group * per_group;
  Sc &sc = *reinterpret_cast<Sc *>(mine);
  const int n_act = a.buf.n_act[a.parity];
  const int i = blockIdx.x * kWarpsPerCta + group;
  if (blockIdx.x == 0 && threadIdx.x == 0) a.buf.n_act[a.parity ^ 1] = a.buf.n_act[2] = a.buf.n_act[3] = 0; /* the lists this trip's accept phase fills */
  if (i >= n_act) return;
  const long long b = a.buf.act[(size_t)a.parity * a.B + i];
  const size_t T = (size_t)a.P.T;
  SlotPtrs<S> sl;
  sl.F = a.buf.F + b * T * NM * N;
  sl.C = CD == kCostFD ? a.buf.C + b * T * Sc::NCF : nullptr;
  sl.cand_x = nullptr;
  sl.cand_u = nullptr;
  sl.gterm = reinterpret_cast<S *>(mine + sizeof(Sc));
  Ex ex;
  ex.lane = threadIdx.x & 31;
  ex.mask = 0xffffffffu;
  ex.init_barrier(reinterpret_cast<unsigned long long *>(mine + per_group - 16));
  CoreT core(a.P, sc, ex, phase_pointers(a, b, N, M), sl);
  core.load_state();
  core.trips_left = 1;
  core.have_derivs = !a.force_sweep;
  const int pre = core.trip_pre();
  ex.lanes([&](int lane, typename CoreT::Lane &) {
    if (lane == 0) sc.st.roll = pre == CoreT::kTripRoll ? kRollGo : (pre == CoreT::kTripNoRoll ? kRollSkip : kRollStop);
  });
  core.store_state();
}

#ifndef ILQR_ROLLOUT_THREADS
#define ILQR_ROLLOUT_THREADS 64
#endif
#ifndef ILQR_ROLLOUT_MINB
#define ILQR_ROLLOUT_MINB 1
#endif
constexpr int kRolloutThreads = ILQR_ROLLOUT_THREADS;
template <class Model, typename S, int CD, int MODE>
__global__ void __launch_bounds__(kRolloutThreads, ILQR_ROLLOUT_MINB) phase_rollout_kernel(const __grid_constant__ PArgs<S> a) {
  using Ph = Phases<Model, S, CD>;
  constexpr int N = Model::N, M = Model::M;
  const int list = a.stage == 2 ? 2 : a.parity;
  const int n_act = a.buf.n_act[list];
  const int na = a.P.n_alpha, nc = a.cand_hi - a.cand_lo;
  const long long tid = blockIdx.x * (long long)kRolloutThreads + threadIdx.x;
  const long long i = tid / nc;
  if (i >= n_act) return;
  const int cand = a.cand_lo + (int)(tid - i * nc);
  const long long b = a.buf.act[(size_t)list * a.B + i];
  if (a.st[b].roll != kRollGo) return;
  const TrajPtrs<S> tr = phase_pointers(a, b, N, M);
  const size_t T = (size_t)a.P.T;
  /* MODE is a template parameter: with both variants behind a run-time branch the kernel carried two copies of the hot
   * loop and configs[4] lost 20 % */
  const S c = Ph::template rollout_task<MODE>(a.P, tr, MODE == Ph::kToCand ? a.buf.cand_x + b * T * na * N : nullptr,
                                              MODE == Ph::kToCand ? a.buf.cand_u + b * T * na * M : nullptr, cand, a.cand_hi);
  a.buf.newcost[b * kMaxAlpha + cand] = c;
}

/* the steps accepted from rollouts that were not kept (re-roll mode; stage 2 of a staged line search), rolled out once
 * more over xs / us: one thread per trajectory of list 3, which the accept phase has filled */
template <class Model, typename S, int CD>
__global__ void __launch_bounds__(kRolloutThreads, ILQR_ROLLOUT_MINB) phase_commit_kernel(const __grid_constant__ PArgs<S> a) {
  using Ph = Phases<Model, S, CD>;
  constexpr int N = Model::N, M = Model::M;
  const int n_act = a.buf.n_act[3];
  const long long i = blockIdx.x * (long long)kRolloutThreads + threadIdx.x;
  if (i >= n_act) return;
  const long long b = a.buf.act[(size_t)3 * a.B + i];
  const TrajPtrs<S> tr = phase_pointers(a, b, N, M);
  Ph::template rollout_task<Ph::kInPlace>(a.P, tr, nullptr, nullptr, a.st[b].alpha_index);
}

constexpr int kAcceptThreads = 128;
template <class Model, typename S, int CD>
__global__ void __launch_bounds__(kAcceptThreads) phase_accept_kernel(const __grid_constant__ PArgs<S> a) {
  using Ph = Phases<Model, S, CD>;
  constexpr int N = Model::N, M = Model::M;
  const int list = a.stage == 2 ? 2 : a.parity;
  const int n_act = a.buf.n_act[list];
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * kAcceptThreads) >> 5;
  const int na = a.P.n_alpha;
  const size_t T = (size_t)a.P.T;
  for (int i = (blockIdx.x * kAcceptThreads + threadIdx.x) >> 5; i < n_act; i += warps) {
    const long long b = a.buf.act[(size_t)list * a.B + i];
    TrajState<S> *st = a.st + b;
    int code = 0; /* bit 0: accepted; bits 8..: alpha index */
    if (lane == 0 && a.ordered) a.buf.act[(size_t)4 * a.B + i] = 0;
    if (lane == 0 && st->status == kRunning) {
      TrajState<S> s = *st;
      const bool fwd = Ph::accept(a.P, s, a.buf.newcost + b * kMaxAlpha, a.stage == 1 ? a.cand_hi : na);
      if (a.stage == 1 && !fwd && s.roll == kRollGo) {
        /* none of the first few passed: stage 2 decides, from the state as it was (nothing is written back here) */
        a.buf.act[(size_t)2 * a.B + atomicAdd(&a.buf.n_act[2], 1)] = (int)b;
      } else {
        code = fwd ? (1 | (s.alpha_index << 8)) : 0;
        const bool go_on = Ph::schedule(a.P, s, fwd);
        *st = s;
        if (go_on && a.ordered) a.buf.act[(size_t)4 * a.B + i] = 1;
        else if (go_on) a.buf.act[(size_t)(a.parity ^ 1) * a.B + atomicAdd(&a.buf.n_act[a.parity ^ 1], 1)] = (int)b;
      }
    }
    code = __shfl_sync(0xffffffffu, code, 0);
    if ((code & 1) && !a.keep) { /* not kept: phase_commit_kernel re-rolls it */
      if (lane == 0) a.buf.act[(size_t)3 * a.B + atomicAdd(&a.buf.n_act[3], 1)] = (int)b;
    } else if (code & 1) { /* xs[1..T], us[0..T-1] <- the accepted candidate (what forward_pass left, :323,334) */
      const int ai = code >> 8, cw = a.cand_hi;
      const S *__restrict__ cx = a.buf.cand_x + b * T * na * N + (size_t)ai * N;
      const S *__restrict__ cu = a.buf.cand_u + b * T * na * M + (size_t)ai * M;
      S *__restrict__ xs = a.xs + b * (T + 1) * N + N;
      S *__restrict__ us = a.us + b * T * M;
      /* a lane moves whole timesteps (one 128-bit run or two per state), kCopyAhead of them in flight: the candidate's
       * states are one sector every n_alpha, so the copy is a chain of DRAM round trips unless the loads overlap
       * (one element per lane and iteration: 25 dependent-looking iterations, ncu long scoreboard 96 per issue) */
      constexpr int kCopyAhead = 4;
      for (int t0 = lane; t0 < (int)T; t0 += 32 * kCopyAhead) {
        S vx[kCopyAhead][N], vu[kCopyAhead][M];
#pragma unroll
        for (int j = 0; j < kCopyAhead; j++) {
          const int t = t0 + 32 * j;
          if (t < (int)T) {
            load_run<N>(vx[j], cx + (size_t)t * cw * N);
            load_run<M>(vu[j], cu + (size_t)t * cw * M);
          }
        }
#pragma unroll
        for (int j = 0; j < kCopyAhead; j++) {
          const int t = t0 + 32 * j;
          if (t < (int)T) {
            store_run<N>(xs + (size_t)t * N, vx[j]);
            store_run<M>(us + (size_t)t * M, vu[j]);
          }
        }
      }
    }
  }
}

/* The next trip's active list in the ORDER of this one (the first is 0 .. B-1, so every list stays ascending): one
 * CTA walks the flags the accept phase left and appends the survivors.  The 32 trajectories of a warp of the
 * thread-per-trajectory kernels are then neighbours in memory, not a sample of whatever the atomics of ~2000
 * concurrent warps interleaved. */
constexpr int kCompactThreads = 1024, kCompactPer = 8;
template <typename S>
__global__ void __launch_bounds__(kCompactThreads) phase_compact_kernel(const __grid_constant__ PArgs<S> a) {
  __shared__ int warp_total[kCompactThreads / 32];
  __shared__ int base;
  const int n = a.buf.n_act[a.parity];
  const int *__restrict__ src = a.buf.act + (size_t)a.parity * a.B;
  const int *__restrict__ flag = a.buf.act + (size_t)4 * a.B;
  int *__restrict__ dst = a.buf.act + (size_t)(a.parity ^ 1) * a.B;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int start = 0; start < n; start += kCompactThreads * kCompactPer) { /* a thread: kCompactPer consecutive entries */
    const int i0 = start + threadIdx.x * kCompactPer;
    int f[kCompactPer], v[kCompactPer], cnt = 0;
#pragma unroll
    for (int j = 0; j < kCompactPer; j++) {
      const bool in = i0 + j < n;
      f[j] = in ? flag[i0 + j] : 0;
      v[j] = in ? src[i0 + j] : 0;
      cnt += f[j] != 0;
    }
    int incl = cnt; /* inclusive prefix over the warp, then over the warps */
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += y;
    }
    if (lane == 31) warp_total[w] = incl;
    __syncthreads();
    int off = base + incl - cnt;
    for (int q = 0; q < w; q++) off += warp_total[q];
#pragma unroll
    for (int j = 0; j < kCompactPer; j++)
      if (f[j] != 0) dst[off++] = v[j];
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int q = 0; q < kCompactThreads / 32; q++) tot += warp_total[q];
      base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) a.buf.n_act[a.parity ^ 1] = base;
}

#endif /* __CUDACC__ */

}  // namespace ilqr
#endif
