// tests/emu/ilqr_emu.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the kernel core (ilqr_b200/csrc/ilqr_core.cuh, the exact header the CUDA kernels
// instantiate) with g++ and runs its warp phases lane by lane on the CPU (ilqr::HostExec), one
// trajectory at a time.  Built with -ffp-contract=off, so for f64 the result must be
// bit-identical to the oracle (oracle/ilqr_oracle.c): that checks the control flow, the lane
// decomposition and the arithmetic order of the kernel source without a GPU.  It is never
// linked into or called by the product (libilqr_b200.so has no CPU path).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../ilqr_b200/csrc/ilqr_core.cuh"
#include "../../ilqr_b200/csrc/ilqr_phases.cuh"
#include "../../ilqr_b200/csrc/params.h"

using namespace ilqr;

namespace {

struct EmuBase {
  virtual ~EmuBase() {}
  virtual double init(const double *x0, const double *u0, int T) = 0;
  virtual double warm_start(const double *x0) = 0;
  virtual void iterate(int n) = 0;
  virtual void resume() = 0;
  virtual int backward_once(double lam) = 0;
  virtual double rollout_once(double alpha) = 0;
  virtual int get(int field, double *dst) = 0;
  virtual double scalar(int which) = 0;
  virtual long integer(int which) = 0;
  int n = 0, m = 0;
};

template <class Model, typename S, int CD, int G>
struct Emu : EmuBase {
  static constexpr int N = Model::N, M = Model::M;
  ilqr_desc desc;
  SolveParams<S> P;
  typename Core<Model, S, CD, HostExec<N, M, S, G>>::Sc sc;
  HostExec<N, M, S, G> ex;
  std::vector<S> x0, xs, us, K, k, Vx0, Vxx0, gterm, bufF, bufC, candX, candU;
  TrajState<S> st;
  int T = 0;

  explicit Emu(const ilqr_desc &d) : desc(d) {
    n = N;
    m = M;
    memset(&sc, 0, sizeof(sc));
    memset(&st, 0, sizeof(st));
  }
  TrajPtrs<S> ptrs() {
    TrajPtrs<S> t;
    t.x0 = x0.data();
    t.xs = xs.data();
    t.us = us.data();
    t.K = K.data();
    t.k = k.data();
    t.Vx0 = Vx0.data();
    t.Vxx0 = Vxx0.data();
    t.st = &st;
    return t;
  }
  SlotPtrs<S> slot() {
    SlotPtrs<S> w;
    w.F = bufF.data();
    w.C = bufC.data();
    w.cand_x = candX.data();
    w.cand_u = candU.data();
    w.gterm = gterm.data();
    return w;
  }
  Core<Model, S, CD, HostExec<N, M, S, G>> core() { return Core<Model, S, CD, HostExec<N, M, S, G>>(P, sc, ex, ptrs(), slot()); }

  double init(const double *x0_, const double *u0_, int T_) override {
    T = T_;
    desc.T = T;
    make_solve_params<S>(desc, &P);
    x0.assign(N, 0);
    xs.assign((size_t)(T + 1) * N, 0);
    us.assign((size_t)T * M, 0);
    K.assign((size_t)T * M * N, 0);
    k.assign((size_t)T * M, 0);
    Vx0.assign(N, 0);
    Vxx0.assign(N * N, 0);
    gterm.assign(T, 0);
    bufF.assign((size_t)T * (N + M) * N, 0);
    bufC.assign((size_t)T * Scratch<N, M, S, CD>::NCF, 0);
    candX.assign((size_t)P.n_alpha * T * N, 0);
    candU.assign((size_t)P.n_alpha * T * M, 0);
    for (int i = 0; i < N; i++) x0[i] = S(x0_[i]);
    for (int i = 0; i < T * M; i++) us[i] = S(u0_[i]);
    core().op_init();
    return st.cost;
  }
  double warm_start(const double *x0_) override {
    for (int i = 0; i < N; i++) x0[i] = S(x0_[i]);
    core().op_warm_start();
    return st.cost;
  }
  void resume() override { /* ilqr_resume_kernel */
    st.iter = 0;
    st.flg_change = 1;
    st.status = kRunning;
  }
  void iterate(int cnt) override {
    if (phases) iterate_phases(cnt);
    else core().op_iterate(cnt);
  }
  /* The batch-lockstep engine (ilqr_phases.cuh) on one trajectory: what phase_begin / sweep / backward / rollout /
   * accept kernels do per trip, every thread's task run in turn by the same ILQR_HD functions the kernels call. */
  bool phases = false;
  void iterate_phases(int cnt) {
    using Ph = Phases<Model, S, CD>;
    const TrajPtrs<S> tr = ptrs();
    if (st.status == kRunning && st.iter >= P.max_iter) st.status = kExitMaxIter;
    bool run = st.status == kRunning;
    const int na = P.n_alpha;
    S newcost[kMaxAlpha] = {0};
    for (int trip = 0; trip < cnt && run; trip++) {
      if (st.flg_change || trip == 0) {
        for (int part = 0; part < Ph::kParts; part++)
          for (int t = 0; t < T; t++) Ph::sweep_task(P, xs.data(), us.data(), bufF.data(), part, t);
        if (CD == kCostFD)
          for (int o = 0; o < Ph::kStencilStep; o++)
            for (int t = 0; t < T; t++) Ph::stencil_task(P, xs.data(), us.data(), bufC.data(), o, t);
      }
      Ph::backward_trip(P, tr, bufF.data(), bufC.data(), gterm.data(), st, 1u);
      /* alternate the three line-search modes of the phase kernels: every one must give the oracle's bits */
      const int mode = trip % 3; /* 0: all candidates kept; 1: cost only + re-roll; 2: staged (ilqr_phases.cuh: PArgs::stage) */
      const bool reroll = mode == 1;
      bool fwd = false;
      if (mode == 2) {
        const int k = na < 4 ? na : 4; /* kPhaseStageK */
        if (st.roll == kRollGo)
          for (int a = 0; a < k; a++) newcost[a] = Ph::template rollout_task<Ph::kToCand>(P, tr, candX.data(), candU.data(), a, k);
        if (st.status != kRunning) break;
        TrajState<S> s1 = st; /* stage 1 decides on a copy: nothing is written back unless it accepts or there was no search */
        const bool f1 = Ph::accept(P, s1, newcost, k);
        if (f1 || st.roll != kRollGo) {
          st = s1;
          fwd = f1;
          if (fwd) {
            const int ai = st.alpha_index;
            for (int t = 0; t < T; t++) {
              for (int c = 0; c < N; c++) xs[(size_t)(t + 1) * N + c] = candX[((size_t)t * k + ai) * N + c];
              for (int c = 0; c < M; c++) us[(size_t)t * M + c] = candU[((size_t)t * k + ai) * M + c];
            }
          }
        } else { /* stage 2: the remaining candidates, cost only; an accepted one is re-rolled */
          for (int a = k; a < na; a++) newcost[a] = Ph::template rollout_task<Ph::kCostOnly>(P, tr, nullptr, nullptr, a);
          fwd = Ph::accept(P, st, newcost);
          if (fwd) Ph::template rollout_task<Ph::kInPlace>(P, tr, nullptr, nullptr, st.alpha_index);
        }
      } else {
        if (st.roll == kRollGo)
          for (int a = 0; a < na; a++)
            newcost[a] = reroll ? Ph::template rollout_task<Ph::kCostOnly>(P, tr, nullptr, nullptr, a)
                                : Ph::template rollout_task<Ph::kToCand>(P, tr, candX.data(), candU.data(), a);
        if (st.status != kRunning) break;
        fwd = Ph::accept(P, st, newcost);
        if (fwd) {
          const int ai = st.alpha_index;
          if (reroll) {
            Ph::template rollout_task<Ph::kInPlace>(P, tr, nullptr, nullptr, ai);
          } else {
            for (int t = 0; t < T; t++) {
              for (int c = 0; c < N; c++) xs[(size_t)(t + 1) * N + c] = candX[((size_t)t * na + ai) * N + c];
              for (int c = 0; c < M; c++) us[(size_t)t * M + c] = candU[((size_t)t * na + ai) * M + c];
            }
          }
        }
      }
      run = Ph::schedule(P, st, fwd);
    }
  }
  int backward_once(double lam) override {
    core().op_backward_once(S(lam));
    return st.diverge;
  }
  double rollout_once(double alpha) override {
    core().op_rollout_once(S(alpha));
    return st.cost;
  }
  int get(int field, double *dst) override {
    const std::vector<S> *v = nullptr;
    switch (field) {
      case 0: v = &xs; break;
      case 1: v = &us; break;
      case 2: v = &K; break;
      case 3: v = &k; break;
      case 4: dst[0] = st.cost; return 1;
      case 5: dst[0] = st.dV0; dst[1] = st.dV1; return 2;
      case 6: v = &Vx0; break;
      case 7: v = &Vxx0; break;
      default: return -1;
    }
    for (size_t i = 0; i < v->size(); i++) dst[i] = (*v)[i];
    return (int)v->size();
  }
  double scalar(int which) override {
    switch (which) {
      case 0: return st.lam;
      case 1: return st.dlam;
      case 2: return st.gnorm;
      case 3: return st.dcost;
      case 4: return st.expected;
      case 5: return st.alpha;
      case 6: return st.new_cost;
      default: return 0;
    }
  }
  long integer(int which) override {
    switch (which) {
      case 0: return st.iter;
      case 1: return st.trips;
      case 2: return st.status;
      case 3: return st.alpha_index;
      case 4: return st.n_accept;
      case 5: return st.n_reject;
      case 6: return st.n_rollouts;
      case 7: return st.n_backward;
      case 8: return st.n_deriv;
      case 9: return T;
      case 10: return st.diverge;
      default: return -1;
    }
  }
};

int g_lanes = 32; /* lanes per trajectory of the next emu_new: 32 (one trajectory per warp), 16 (two per warp), or
                    1 = the batch-lockstep phase engine (one thread per task) */

template <class Model, typename S>
EmuBase *make_cd(const ilqr_desc &d) {
  if (g_lanes == 1) { /* the phase engine: thread-per-task functions of ilqr_phases.cuh */
    if (d.cost_deriv == ILQR_COST_ANALYTIC) {
      auto *e = new Emu<Model, S, kCostAnalytic, 32>(d);
      e->phases = true;
      return e;
    }
    auto *e = new Emu<Model, S, kCostFD, 32>(d);
    e->phases = true;
    return e;
  }
  if (g_lanes == 16) {
    if (d.cost_deriv == ILQR_COST_ANALYTIC) return new Emu<Model, S, kCostAnalytic, 16>(d);
    return new Emu<Model, S, kCostFD, 16>(d);
  }
  if (d.cost_deriv == ILQR_COST_ANALYTIC) return new Emu<Model, S, kCostAnalytic, 32>(d);
  return new Emu<Model, S, kCostFD, 32>(d);
}
template <class Model>
EmuBase *make_dtype(const ilqr_desc &d) {
  if (d.dtype == ILQR_F32) return make_cd<Model, float>(d);
  return make_cd<Model, double>(d);
}

template <int M>
int run_qp(const ilqr_params &p, int generic, const double *Q, const double *c, const double *x0, const double *lo,
           const double *hi, double *x_opt, int *v_free, double *R_free, int *r_dim) {
  ilqr_desc d;
  memset(&d, 0, sizeof(d));
  d.params = p;
  d.T = 1;
  d.model_id = ILQR_MODEL_ACROBOT;
  SolveParams<double> P;
  make_solve_params<double>(d, &P);
  static QPWork<M, double> w;
  memset(&w, 0, sizeof(w));
  for (int i = 0; i < M * M; i++) w.Q[i] = Q[i];
  for (int i = 0; i < M; i++) {
    w.c[i] = c[i];
    w.x0[i] = x0[i];
    w.lo[i] = lo[i];
    w.hi[i] = hi[i];
  }
  if (generic) box_qp_generic<M, double>(P.qp, w);
  else box_qp<M, double>(P.qp, w);
  for (int i = 0; i < M; i++) {
    x_opt[i] = w.x[i];
    v_free[i] = w.v_free[i];
  }
  *r_dim = w.r_dim;
  for (int i = 0; i < w.r_dim * w.r_dim; i++) R_free[i] = w.R[i];
  return w.result;
}

}  // namespace

extern "C" {

void emu_set_lanes(int lanes) { g_lanes = lanes == 16 ? 16 : (lanes == 1 ? 1 : 32); }

void *emu_new(const ilqr_desc *d) {
  if (d->model_id == ILQR_MODEL_ACROBOT) return make_dtype<Acrobot>(*d);
  if (d->model_id == ILQR_MODEL_DOUBLE_INTEGRATOR) return make_dtype<DoubleIntegrator>(*d);
  return nullptr;
}
void emu_free(void *h) { delete (EmuBase *)h; }
void emu_dims(void *h, int *n, int *m) {
  *n = ((EmuBase *)h)->n;
  *m = ((EmuBase *)h)->m;
}
double emu_init(void *h, const double *x0, const double *u0, int T) { return ((EmuBase *)h)->init(x0, u0, T); }
double emu_warm_start(void *h, const double *x0) { return ((EmuBase *)h)->warm_start(x0); }
int emu_iterate(void *h, int n) {
  ((EmuBase *)h)->iterate(n);
  return 0;
}
void emu_resume(void *h) { ((EmuBase *)h)->resume(); }
int emu_backward_once(void *h, double lam) { return ((EmuBase *)h)->backward_once(lam); }
double emu_rollout_once(void *h, double alpha) { return ((EmuBase *)h)->rollout_once(alpha); }
int emu_get(void *h, int field, double *dst) { return ((EmuBase *)h)->get(field, dst); }
double emu_scalar(void *h, int which) { return ((EmuBase *)h)->scalar(which); }
long emu_int(void *h, int which) { return ((EmuBase *)h)->integer(which); }

/* boxQP leaf: generic != 0 forces the general-m code path even for m == 1 */
int emu_boxqp(const ilqr_params *p, int m, int generic, const double *Q, const double *c, const double *x0,
              const double *lo, const double *hi, double *x_opt, int *v_free, double *R_free, int *r_dim) {
  ilqr_params dp;
  if (!p) {
    default_params(&dp);
    p = &dp;
  }
  switch (m) {
    case 1: return run_qp<1>(*p, generic, Q, c, x0, lo, hi, x_opt, v_free, R_free, r_dim);
    case 2: return run_qp<2>(*p, generic, Q, c, x0, lo, hi, x_opt, v_free, R_free, r_dim);
    case 3: return run_qp<3>(*p, generic, Q, c, x0, lo, hi, x_opt, v_free, R_free, r_dim);
    case 4: return run_qp<4>(*p, generic, Q, c, x0, lo, hi, x_opt, v_free, R_free, r_dim);
    default: return -100;
  }
}

}  // extern "C"
