/*
 * ilqr_phases.cuh — the batch-lockstep engine: one loop trip of src/ilqr_core.cpp:103-288 for EVERY running
 * trajectory of the batch as four kernels, each with the thread mapping that fills its lanes.
 *
 * The warp-per-trajectory kernel (ilqr_kernel.cuh) gives a trajectory 32 lanes for every phase of a trip, but the
 * phases do not have 32-way work: the line search has n_alpha = 11 rollouts, the backward step (n+m)(n+1) = 25
 * four-term dot products and then a scalar boxQP that every lane repeats, so two thirds of the fp64 lane slots
 * it issues carry nothing (profiles/r1n: 129 k warp instructions per trip, 20.8 lanes active, boxQP 19 % of them
 * redundant).  Here the unit of parallelism changes with the phase instead:
 *
 *   sweep     one thread per (trajectory, timestep, variable group): central differences of the Euler step
 *             (src/derivatives.cpp:15-26) and, in FD-cost mode, one cost-stencil output per thread (:29-144);
 *   backward  one thread per trajectory: the whole recursion (:350-401) in registers — Va = [Vxx | Vx], the
 *             Jacobian of the step, the Q-function — with boxQP (src/boxqp.cpp) inline; 32 trajectories per warp,
 *             every fp64 instruction 32 useful lanes, no shared memory, no barriers.  The next timestep's
 *             Jacobian / state / control are loaded while the current step computes;
 *   rollout   one thread per (trajectory, alpha): the n_alpha candidates of the line search (:184-226, :305-337),
 *             adjacent lanes share the trajectory's nominal arrays (broadcast loads), each streams its candidate
 *             to the trajectory's candidate buffer;
 *   accept    one warp per trajectory: lane 0 applies the reference's serial acceptance order and the lambda
 *             schedule (:199-282), all lanes commit the accepted candidate by a coalesced copy, and trajectories
 *             that go on are appended to the next trip's active list.
 *
 * The batch advances one trip per round of launches; finished trajectories leave the active list, so every launch
 * covers exactly the running ones (ragged trip counts cost nothing but the per-trip latency floor).  Arithmetic is
 * the warp kernel's, entry for entry, in the same order (the helpers are Core's static functions), so both engines
 * return identical bits; the per-thread functions are plain ILQR_HD code and tests/emu runs them on the CPU against
 * the oracle.
 *
 * HBM (per handle, all per TRAJECTORY): F [B][T][n+m][n] Jacobian columns, C [B][T][NCF] (FD-cost mode),
 * cand_x [B][T][n_alpha][n], cand_u [B][T][n_alpha][m] (candidate-interleaved: the n_alpha stores of one timestep
 * are contiguous), newcost [B][16], act [2][B] active lists.
 */
#ifndef ILQR_PHASES_CUH_
#define ILQR_PHASES_CUH_

#include "ilqr_core.cuh"

namespace ilqr {

struct NoExec {
  static constexpr int kLanes = 1;
};

enum { kRollStop = 0, kRollGo = 1, kRollSkip = 2 }; /* TrajState::roll */

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
  S *cand_x;  /* [B][T][n_alpha][n]  candidate states x_1..x_T */
  S *cand_u;  /* [B][T][n_alpha][m]  candidate controls */
  S *newcost; /* [B][kMaxAlpha]      */
  int *act;   /* [2][B]              active lists, double-buffered by trip parity */
  int *n_act; /* [2]                 their lengths */
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
    S x[N], u[M], xa[N], fp[N], fm[N];
    load_run<N>(x, xs + (size_t)t * N);
    load_run<M>(u, us + (size_t)t * M);
    if (part < CoreT::kNumConfigVars) {
      const int j = CoreT::nth_config_var(part);
      CoreT::template perturb<N>(x, j, P.fd_eps, -1, S(0), xa);
      integrate<Model, S>(xa, u, P.mp, P.dt, fp);
      CoreT::template perturb<N>(x, j, -P.fd_eps, -1, S(0), xa);
      integrate<Model, S>(xa, u, P.mp, P.dt, fm);
      S col[N];
#pragma unroll
      for (int r = 0; r < N; r++) col[r] = (fp[r] - fm[r]) / (2 * P.fd_eps);
      store_run<N>(F + ((size_t)t * NM + j) * N, col);
      return;
    }
    S ua[M];
    typename Model::template Config<S> cf;
    Model::configure(x, P.mp, cf);
#pragma unroll
    for (int j = 0; j < NM; j++) {
      if (CoreT::is_config_var(j)) continue;
      CoreT::template perturb<N>(x, j, P.fd_eps, -1, S(0), xa);
      CoreT::template perturb<M>(u, j - N, P.fd_eps, -1, S(0), ua);
      integrate_cfg<Model, S>(cf, xa, ua, P.mp, P.dt, fp);
      CoreT::template perturb<N>(x, j, -P.fd_eps, -1, S(0), xa);
      CoreT::template perturb<M>(u, j - N, -P.fd_eps, -1, S(0), ua);
      integrate_cfg<Model, S>(cf, xa, ua, P.mp, P.dt, fm);
      S col[N];
#pragma unroll
      for (int r = 0; r < N; r++) col[r] = (fp[r] - fm[r]) / (2 * P.fd_eps);
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
  ILQR_HD static int backward_pass(const SolveParams<S> &P, const TrajPtrs<S> &tr, const S *F, const S *Cfd, S lam, S &dV0,
                                   S &dV1) {
    const int T = P.T;
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
        if (r.result < 1) return t; /* :371 */
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
        if (w.result < 1) return t;
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
  ILQR_HD static S gradient_norm(const SolveParams<S> &P, const TrajPtrs<S> &tr) {
    const int T = P.T;
    S acc = 0;
    for (int t = 0; t < T; t++) acc += CoreT::gn_term(tr.k + (size_t)t * M, tr.us + (size_t)t * M);
    return acc / T;
  }

  /* The head of a loop trip for one running trajectory (:115-159): bookkeeping of the derivative refresh (done by
   * the sweep phase just before), backward pass with the lambda retries, gradient-norm exit.  Leaves in s.roll what
   * the line search has to do.  s lives in registers / local memory; the caller stores it back. */
  ILQR_HD static void backward_trip(const SolveParams<S> &P, const TrajPtrs<S> &tr, const S *F, const S *Cfd, TrajState<S> &s) {
    s.trips++;
    if (s.flg_change) {
      s.flg_change = 0;
      s.n_deriv++;
    }
    bool back_done = false;
    for (;;) { /* :136-150 */
      S dV0, dV1;
      s.n_backward++;
      const int diverge = backward_pass(P, tr, F, Cfd, s.lam, dV0, dV1);
      s.dV0 = dV0;
      s.dV1 = dV1;
      s.diverge = diverge;
      if (diverge != 0) {
        s.dlam = CoreT::fmax_(s.dlam * P.lambda_factor, P.lambda_factor);
        s.lam = CoreT::fmax_(s.lam * s.dlam, P.lambda_min);
        if (s.lam > P.lambda_max) break;
        continue;
      }
      back_done = true;
      break;
    }
    s.gnorm = gradient_norm(P, tr);
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
   * returned.  cand_x / cand_u point at the trajectory's [T][na][.] block. */
  ILQR_HD static S rollout_task(const SolveParams<S> &P, const TrajPtrs<S> &tr, S *cand_x, S *cand_u, int a) {
    const int T = P.T, na = P.n_alpha;
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
      S uc[M];
#pragma unroll
      for (int j = 0; j < M; j++) {
        S v = ub[j] + kt[j] * alpha; /* :188-190 */
        Acc<S> acc;                  /* :316 */
#pragma unroll
        for (int i = 0; i < N; i++) acc.add(Kt[j * N + i] * (x[i] - xh[i]));
        v += acc.v;
        uc[j] = v;
      }
      cost += Model::cost(x, uc, mp); /* :324 */
      S x1[N];
      integrate<Model, S>(x, uc, mp, P.dt, x1); /* :325 */
#pragma unroll
      for (int i = 0; i < N; i++) x[i] = x1[i];
      store_run<M>(cand_u + ((size_t)t * na + a) * M, uc);
      store_run<N>(cand_x + ((size_t)t * na + a) * N, x);
    }
    cost += Model::final_cost(x, mp); /* :335 */
    return cost;
  }

  /* ---- accept --------------------------------------------------------------------------------------------- */

  /* the acceptance test (:199-213) in the reference's serial order over the candidates' costs; true = a step was
   * accepted (s.alpha_index says which) */
  ILQR_HD static bool accept(const SolveParams<S> &P, TrajState<S> &s, const S *newcost) {
    const bool back_done = s.roll == kRollGo;
    bool fwd_done = false;
    s.alpha_index = -1;
    S alpha = 0;
    if (back_done) {
      for (int a = 0; a < P.n_alpha; a++) {
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
template <class Model, typename S, int CD>
__global__ void __launch_bounds__(kBackwardThreads) phase_backward_kernel(const __grid_constant__ PArgs<S> a) {
  using Ph = Phases<Model, S, CD>;
  constexpr int N = Model::N, M = Model::M, NM = N + M;
  const int n_act = a.buf.n_act[a.parity];
  const int i = blockIdx.x * kBackwardThreads + threadIdx.x;
  if (i == 0) a.buf.n_act[a.parity ^ 1] = 0; /* the list this trip's accept phase fills */
  if (i >= n_act) return;
  const long long b = a.buf.act[(size_t)a.parity * a.B + i];
  const TrajPtrs<S> tr = phase_pointers(a, b, N, M);
  const S *F = a.buf.F + b * (size_t)a.P.T * NM * N;
  const S *Cfd = CD == kCostFD ? a.buf.C + b * (size_t)a.P.T * Ph::NCF : nullptr;
  TrajState<S> s = *tr.st;
  Ph::backward_trip(a.P, tr, F, Cfd, s);
  *tr.st = s;
}

constexpr int kRolloutThreads = 64;
template <class Model, typename S, int CD>
__global__ void __launch_bounds__(kRolloutThreads) phase_rollout_kernel(const __grid_constant__ PArgs<S> a) {
  using Ph = Phases<Model, S, CD>;
  constexpr int N = Model::N, M = Model::M;
  const int n_act = a.buf.n_act[a.parity];
  const int na = a.P.n_alpha;
  const long long tid = blockIdx.x * (long long)kRolloutThreads + threadIdx.x;
  const long long i = tid / na;
  if (i >= n_act) return;
  const int cand = (int)(tid - i * na);
  const long long b = a.buf.act[(size_t)a.parity * a.B + i];
  if (a.st[b].roll != kRollGo) return;
  const TrajPtrs<S> tr = phase_pointers(a, b, N, M);
  const size_t T = (size_t)a.P.T;
  const S c = Ph::rollout_task(a.P, tr, a.buf.cand_x + b * T * na * N, a.buf.cand_u + b * T * na * M, cand);
  a.buf.newcost[b * kMaxAlpha + cand] = c;
}

constexpr int kAcceptThreads = 128;
template <class Model, typename S, int CD>
__global__ void __launch_bounds__(kAcceptThreads) phase_accept_kernel(const __grid_constant__ PArgs<S> a) {
  using Ph = Phases<Model, S, CD>;
  constexpr int N = Model::N, M = Model::M;
  const int n_act = a.buf.n_act[a.parity];
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * kAcceptThreads) >> 5;
  const int na = a.P.n_alpha;
  const size_t T = (size_t)a.P.T;
  for (int i = (blockIdx.x * kAcceptThreads + threadIdx.x) >> 5; i < n_act; i += warps) {
    const long long b = a.buf.act[(size_t)a.parity * a.B + i];
    TrajState<S> *st = a.st + b;
    int code = 0; /* bit 0: accepted; bits 8..: alpha index */
    if (lane == 0 && st->status == kRunning) {
      TrajState<S> s = *st;
      const bool fwd = Ph::accept(a.P, s, a.buf.newcost + b * kMaxAlpha);
      code = fwd ? (1 | (s.alpha_index << 8)) : 0;
      const bool go_on = Ph::schedule(a.P, s, fwd);
      *st = s;
      if (go_on) a.buf.act[(size_t)(a.parity ^ 1) * a.B + atomicAdd(&a.buf.n_act[a.parity ^ 1], 1)] = (int)b;
    }
    code = __shfl_sync(0xffffffffu, code, 0);
    if (code & 1) { /* xs[1..T], us[0..T-1] <- the accepted candidate (what forward_pass left, :323,334) */
      const int ai = code >> 8;
      const S *cx = a.buf.cand_x + b * T * na * N;
      const S *cu = a.buf.cand_u + b * T * na * M;
      S *xs = a.xs + b * (T + 1) * N + N;
      S *us = a.us + b * T * M;
      const int nx = (int)T * N, nu = (int)T * M;
      for (int e = lane; e < nx; e += 32) {
        const int t = e / N, c = e - t * N;
        xs[e] = cx[((size_t)t * na + ai) * N + c];
      }
      for (int e = lane; e < nu; e += 32) {
        const int t = e / M, c = e - t * M;
        us[e] = cu[((size_t)t * na + ai) * M + c];
      }
    }
  }
}

#endif /* __CUDACC__ */

}  // namespace ilqr
#endif
