/*
 * ilqr_oracle.h — CPU ORACLE for the batched-iLQR hot path.  TEST INFRASTRUCTURE.
 *
 * A plain-C restatement of kazuotani14/iLQR's solve path (src/ilqr_core.cpp,
 * src/derivatives.cpp, src/boxqp.cpp, include/finite_diff.h, include/acrobot.h,
 * include/double_integrator.h); every function in ilqr_oracle.c cites the lines it follows.
 * Parity of this port is PINNED: tests/test_oracle_port.py checks it against the reference's
 * own known-answer tests (test/test_boxqp.cpp, test_finite_diff.cpp, test_dynamicsmodels.cpp,
 * test_ilqr_forward_pass.cpp, test_ilqr_derivatives.cpp) and against golden vectors produced
 * by the unmodified reference (oracle/_ref, tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg
 * may load this library.  The product (libilqr_b200.so) never links or calls it.
 */
#ifndef ILQR_ORACLE_H_
#define ILQR_ORACLE_H_

#include "ilqr_b200.h" /* ilqr_params, ilqr_desc, field ids, status codes: the API being checked */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_solver orc_solver;

/* uses desc->model_id, cost_deriv, dt, T-independent; B, dtype, device are ignored (one instance, f64) */
orc_solver *orc_new(const ilqr_desc *desc);
void orc_free(orc_solver *s);
void orc_dims(const orc_solver *s, int *n, int *m);

double orc_init(orc_solver *s, const double *x0, const double *u0, int T);
double orc_warm_start(orc_solver *s, const double *x0);
void orc_resume(orc_solver *s);
int orc_iterate(orc_solver *s, int n_iters);
int orc_backward_once(orc_solver *s, double lambda, int recompute_derivs);
double orc_rollout_once(orc_solver *s, double alpha);

/* fields 0..14 as tests/refharness.py FIELDS: xs us K k cost dV Vx Vxx fx fu cx cu cxx cxu cuu */
int orc_get(const orc_solver *s, int field, double *dst);
/* 0 lambda 1 dlambda 2 gnorm 3 dcost 4 expected 5 alpha 6 new_cost */
double orc_scalar(const orc_solver *s, int which);
/* 0 iter 1 loop_trips 2 status 3 alpha_index 4 accepts 5 rejects 6 rollouts 7 backwards 8 derivs 9 T 10 diverge */
long orc_int(const orc_solver *s, int which);

/* leaf functions */
void orc_dynamics(const orc_solver *s, const double *x, const double *u, double *dx);
void orc_integrate(const orc_solver *s, const double *x, const double *u, double dt, double *x1);
double orc_cost(const orc_solver *s, const double *x, const double *u);
double orc_final_cost(const orc_solver *s, const double *x);
int orc_fd(const orc_solver *s, int which, const double *x, const double *u, double dt, double *out);

int orc_boxqp(const ilqr_params *p, int m, const double *Q, const double *c, const double *x0, const double *lo,
              const double *hi, double *x_opt, int *v_free, double *R_free, int *r_dim);
int orc_quadclamp(const ilqr_params *p, int m, const double *x0, const double *dir, const double *Q, const double *c,
                  const double *lo, const double *hi, double *x_opt, double *v_opt, int *n_steps);
double orc_quadcost(int m, const double *Q, const double *c, const double *x);

/* Solve instances [b0, b1) of a batch one after another on the calling thread (x0[B][n], u0[B][T][m]).
 * max_trips < 0: run to termination.  Outputs (may be NULL) are indexed by b - b0.  Returns the
 * total number of loop trips executed. */
long orc_solve_range(const ilqr_desc *desc, long b0, long b1, const double *x0, const double *u0, int max_trips,
                     double *cost, int *iters, int *status, long *n_accept, long *n_reject);

#ifdef __cplusplus
}
#endif
#endif
