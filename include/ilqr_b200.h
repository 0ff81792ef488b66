/*
 * ilqr_b200.h — C ABI of the B200-native batched iLQR solver (libilqr_b200.so).
 *
 * This is the drop-in boundary for ONE path of kazuotani14/iLQR: a batch of independent
 * `iLQR::generate_trajectory(x0, u0)` solves (reference include/ilqr.h:49-54,
 * src/ilqr_core.cpp:11-401, src/derivatives.cpp:15-144, src/boxqp.cpp:26-178).  The
 * reference has no FFI of its own (it is a C++ class called in-process, src/run_ilqr.cpp:56-59);
 * the entry points below are what a binding for that class would call, one per public method
 * plus getters for the results the reference keeps private (include/ilqr.h:57-85).
 * INTEGRATION.md shows the reference-side shim (a `class iLQR` with the reference's
 * signatures forwarding here) and ilqr_b200/host/ holds that shim.
 *
 * Conventions
 *   - plain C, no CUDA/torch types; every function returns 0 on success and a negative
 *     ILQR_E_* code on failure, never throws; ilqr_last_error() gives the message.
 *   - the caller owns every buffer it passes; the handle owns all device memory.
 *   - host-side array layout is trajectory-major, row-major, exactly how the reference's
 *     VecOfVecXd / VecOfMatXd would serialise:  x0[B][n], u0[B][T][m], xs[B][T+1][n],
 *     us[B][T][m], K[B][T][m][n], k[B][T][m].  `on_device != 0` means the pointer is a CUDA
 *     device pointer on the handle's device (same layout); scalars are `dtype` (f64 or f32).
 *   - lambda/dlambda are PER TRAJECTORY and start at 1 ("fresh process per trajectory");
 *     in the reference they are process-wide statics (include/ilqr.h:17-18).
 *   - a handle is bound to one device and one CUDA stream; distinct handles may be used
 *     from distinct threads.
 */
#ifndef ILQR_B200_H_
#define ILQR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ILQR_MAX_N 8      /* state dimension limit  (n_x) */
#define ILQR_MAX_M 4      /* control dimension limit (n_u) */
#define ILQR_MAX_ALPHA 16 /* line-search table limit */

/* model_id: device twins of the reference's Model subclasses */
#define ILQR_MODEL_ACROBOT 0            /* include/acrobot.h            n=4 m=1 */
#define ILQR_MODEL_DOUBLE_INTEGRATOR 1  /* include/double_integrator.h  n=4 m=2, model_params[0..3] = goal */
#define ILQR_MODEL_USER_BASE 100        /* ids >= 100: models registered at run time, ilqr_register_model() */

#define ILQR_F64 0
#define ILQR_F32 1

/* how cx, cu, cxx, cxu, cuu are obtained */
#define ILQR_COST_FD 0        /* the reference's stencils, src/derivatives.cpp:29-144, include/finite_diff.h:22-86 */
#define ILQR_COST_ANALYTIC 1  /* closed form from the model twin (BASELINE configs 2, 3, 5) */

/* ilqr_desc.flags */
#define ILQR_FLAG_ENGINE_WARP 1    /* run ilqr_iterate / ilqr_solve on the persistent warp-per-trajectory kernel instead of
                                      the batch-lockstep phase kernels (identical results) */
#define ILQR_FLAG_CLAMP_ROLLOUT 2  /* OPT-IN, changes results: the control a rollout applies is clamped to [u_min, u_max]
                                      ("the right way" the reference comments out, src/ilqr_core.cpp:322-329; by
                                      default rollouts are unclamped like the reference's) */
#define ILQR_FLAG_ANALYTIC_DYN 4   /* OPT-IN, changes results: fx, fu from the model twin's closed-form Jacobian
                                      (Model::dynamics_jac) instead of the reference's central differences
                                      (src/derivatives.cpp:15-26; the reference lists this as future work, notes.md:15,45) */
#define ILQR_FLAG_FAST_FMA 8       /* kernels built WITH fused multiply-add contraction: faster (dot products lose half their
                                      dependent chain), results within 1e-6 of the default build after a handful of trips
                                      but not bit-identical to the reference's no-FMA arithmetic.  Built-in models only. */
#define ILQR_FLAG_ALL (ILQR_FLAG_ENGINE_WARP | ILQR_FLAG_CLAMP_ROLLOUT | ILQR_FLAG_ANALYTIC_DYN | ILQR_FLAG_FAST_FMA)

/* error codes */
#define ILQR_OK 0
#define ILQR_E_INVALID -1   /* bad argument / unsupported combination */
#define ILQR_E_CUDA -2      /* CUDA runtime error (message in ilqr_last_error) */
#define ILQR_E_STATE -3     /* call out of order (e.g. iterate before set_initial) */
#define ILQR_E_NOMEM -4

/* per-trajectory exit reason (the reference only prints these: src/ilqr_core.cpp:156,259,278,285) */
#define ILQR_RUNNING 0
#define ILQR_EXIT_GRAD 1        /* gnorm < tolGrad && lambda < 1e-5      :154 */
#define ILQR_EXIT_TOLFUN 2      /* accepted step with dcost < tolFun     :257 */
#define ILQR_EXIT_LAMBDA_MAX 3  /* rejected step and lambda > lambdaMax  :276 */
#define ILQR_EXIT_MAXITER 4     /* loop counter reached maxIter          :103 */

/* Every tunable the reference hard-codes; ilqr_default_params() fills in its values. */
typedef struct ilqr_params {
  int32_t max_iter;        /* include/ilqr.h:14   100   */
  int32_t n_alpha;         /* include/ilqr.h:24   11    */
  double tol_fun;          /* include/ilqr.h:15   1e-6  */
  double tol_grad;         /* include/ilqr.h:16   1e-6  */
  double lambda_init;      /* include/ilqr.h:17   1     */
  double dlambda_init;     /* include/ilqr.h:18   1     */
  double lambda_factor;    /* include/ilqr.h:19   1.6   */
  double lambda_max;       /* include/ilqr.h:20   1e11  */
  double lambda_min;       /* include/ilqr.h:21   1e-8  */
  double z_min;            /* include/ilqr.h:22   0     */
  double grad_lambda_gate; /* src/ilqr_core.cpp:154  1e-5 */
  double alpha[ILQR_MAX_ALPHA]; /* include/ilqr.h:24 (the 11 literal values) */
  int32_t qp_max_iter;     /* include/boxqp.h:19  100   */
  int32_t reserved0;
  double qp_min_grad;        /* include/boxqp.h:20  1e-8  */
  double qp_min_rel_improve; /* include/boxqp.h:21  1e-8  */
  double qp_step_dec;        /* include/boxqp.h:22  0.6   */
  double qp_min_step;        /* include/boxqp.h:23  1e-22 */
  double qp_armijo;          /* include/boxqp.h:24  0.1   */
  double qp_clamp_tol;       /* include/boxqp.h:61-64  1e-4 (approx_eq) */
  double fd_eps;             /* include/finite_diff.h:9 and src/derivatives.cpp:10  1e-3 */
} ilqr_params;

typedef struct ilqr_desc {
  int32_t model_id;   /* ILQR_MODEL_* */
  int32_t dtype;      /* ILQR_F64 | ILQR_F32 */
  int32_t cost_deriv; /* ILQR_COST_FD | ILQR_COST_ANALYTIC */
  int32_t device;     /* CUDA device ordinal */
  int32_t T;          /* number of controls = u0.size() (src/ilqr_core.cpp:12); knots = T+1 */
  int32_t override_limits; /* 0: the model's own u_min/u_max (acrobot.h:37, double_integrator.h:25-26) */
  int32_t flags;      /* ILQR_FLAG_* bits; 0 = the reference's behaviour on the default engine */
  int32_t reserved1;  /* must be 0 */
  int64_t B;          /* number of independent problem instances */
  double dt;          /* iLQR(Model*, double timeDelta), include/ilqr.h:30 */
  double u_min[ILQR_MAX_M];
  double u_max[ILQR_MAX_M];
  double model_params[16]; /* DOUBLE_INTEGRATOR: [0..3] goal state; ACROBOT: [0..3] goal state, all zero = the
                              reference's (3.1415, 0, 0, 0) (acrobot.h:20-21); user models: passed through as `mp` */
  ilqr_params params;
} ilqr_desc;

typedef struct ilqr_handle ilqr_handle;

/* fields of ilqr_get(); element type is the handle's dtype unless noted */
#define ILQR_F_XS 0        /* [B][T+1][n]  current trajectory                       (ilqr.h:62)  */
#define ILQR_F_US 1        /* [B][T][m]                                             (ilqr.h:63)  */
#define ILQR_F_K 2         /* [B][T][m][n] feedback gains                           (ilqr.h:79)  */
#define ILQR_F_KFF 3       /* [B][T][m]    feed-forward k                           (ilqr.h:78)  */
#define ILQR_F_COST 4      /* [B]          cost_s                                   (ilqr.h:66)  */
#define ILQR_F_DV 5        /* [B][2]       expected-reduction terms                 (ilqr.h:75)  */
#define ILQR_F_VX0 6       /* [B][n]       Vx[0]                                    (ilqr.h:76)  */
#define ILQR_F_VXX0 7      /* [B][n][n]    Vxx[0]                                   (ilqr.h:77)  */
#define ILQR_F_LAMBDA 8    /* [B] */
#define ILQR_F_DLAMBDA 9   /* [B] */
#define ILQR_F_GNORM 10    /* [B]  last get_gradient_norm (src/ilqr_core.cpp:405-412) */
#define ILQR_F_ITERS 11    /* [B] int32: loop bodies of src/ilqr_core.cpp:103-288 entered */
#define ILQR_F_STATUS 12   /* [B] int32: ILQR_RUNNING / ILQR_EXIT_* */
#define ILQR_F_ALPHA_INDEX 13 /* [B] int32: index into alpha[] accepted by the last line search, -1 = NO STEP */
#define ILQR_F_N_ACCEPT 14    /* [B] int32: accepted iterations so far  */
#define ILQR_F_N_REJECT 15    /* [B] int32: rejected ("NO STEP") iterations so far */
#define ILQR_F_N_BACKWARD 16  /* [B] int32: backward passes run (incl. lambda retries, :137-150) */
#define ILQR_F_DIVERGE 17     /* [B] int32: return value of the last backward pass (:371,400) */

int ilqr_default_params(ilqr_params *p);
/* n, m, default limits of a model twin (Model::x_dims/u_dims/u_min/u_max, include/model.h:17-20) */
int ilqr_model_info(int32_t model_id, int32_t *n, int32_t *m, double *u_min, double *u_max);

/* The plugin surface on the GPU side (include/model.h:6-21: a user writes dynamics / cost / final_cost and hands
 * the Model to the solver).  A host vtable cannot run in a kernel, so a user model is its device twin as CUDA source:
 * a struct `struct_name` with the static interface of ilqr_b200/csrc/models.cuh — N, M, dynamics, cost, final_cost,
 * cost_d1, cost_d2 (closed-form cost derivatives; may return 0 if only ILQR_COST_FD is used), kConfigVars, Config,
 * configure, dynamics_cfg — templated on the scalar type and marked ILQR_HD.  The library compiles its own solver
 * kernel around that struct with NVRTC (sm_100a, no FMA contraction) the first time a handle of that model
 * launches; ilqr_desc.model_params[0..15] reach the struct's functions as `mp`.  u_min / u_max [m] are the model's
 * own limits (Model::u_min/u_max).  Returns the id to put in ilqr_desc.model_id (>= ILQR_MODEL_USER_BASE). */
int ilqr_register_model(const char *struct_name, const char *cuda_source, int32_t n, int32_t m, const double *u_min,
                        const double *u_max, int32_t *model_id);
/* Compile a registered model now (NVRTC only: needs no GPU) and report the compiler's log; 0 = it compiles. */
int ilqr_compile_model(int32_t model_id, int32_t dtype, int32_t cost_deriv, char *log, size_t log_bytes);

/* `new iLQR(model, dt)` for B instances (include/ilqr.h:30-44). */
int ilqr_create(const ilqr_desc *desc, ilqr_handle **out);
int ilqr_destroy(ilqr_handle *h);
const char *ilqr_last_error(const ilqr_handle *h); /* h may be NULL: last create error */

/* iLQR::init_traj(x_0, u_0) for every instance (src/ilqr_core.cpp:11-56): loads x0[B][n] and
 * u0[B][T][m], runs the open-loop rollout, sets cost_s, zeroes k/K, lambda = dlambda = 1. */
int ilqr_set_initial(ilqr_handle *h, const void *x0, const void *u0, int on_device);

/* Warm start, iLQR::generate_trajectory(x_0) (src/ilqr_core.cpp:65-76): keep us, K, xs, and
 * lambda/dlambda of the previous solve, re-roll from a new x0[B][n] WITH feedback (:316). */
int ilqr_warm_start(ilqr_handle *h, const void *x0, int on_device);

/* "Continue", iLQR::generate_trajectory() called again (src/ilqr_core.cpp:78-102): every instance re-enters the loop
 * with its counter at 0 and its derivatives refreshed, whatever its exit status was; lambda / dlambda carry over
 * (include/ilqr.h:17-18).  Follow with ilqr_solve / ilqr_iterate. */
int ilqr_resume(ilqr_handle *h);

/* Up to n_iters more trips of the loop body (src/ilqr_core.cpp:103-288) per instance;
 * instances that have terminated stay as they are. */
int ilqr_iterate(ilqr_handle *h, int n_iters);
/* iLQR::generate_trajectory(): iterate until every instance has terminated (<= max_iter trips). */
int ilqr_solve(ilqr_handle *h);

/* Single-phase test hooks.  backward_once: derivative sweep at the current (xs, us) and ONE
 * backward pass (src/ilqr_core.cpp:350-401) with lambda forced to `lambda` for every instance;
 * k, K, dV, Vx[0], Vxx[0], gnorm and DIVERGE are left readable.  rollout_once: the line search's
 * closed-loop rollout for one alpha (src/ilqr_core.cpp:188-197, 305-337), committing xs/us and
 * writing the new cost into COST. */
int ilqr_backward_once(ilqr_handle *h, double lambda);
int ilqr_rollout_once(ilqr_handle *h, double alpha);

int ilqr_get(ilqr_handle *h, int field, void *dst, int on_device);
/* Block until all work queued on the handle's stream has finished. */
int ilqr_sync(ilqr_handle *h);
/* The handle's CUDA stream as an opaque pointer (cudaStream_t), for event timing by the caller. */
void *ilqr_stream(ilqr_handle *h);
/* Kernel launches issued on behalf of this handle since creation. */
int64_t ilqr_launch_count(const ilqr_handle *h);

/* Synthetic instances (include/ilqr_synth.h).  Host buffers of doubles. */
int ilqr_make_inputs(uint64_t seed, int64_t B, int32_t T, int32_t n, int32_t m, double x_scale, double u_scale,
                     int canonical_first, double *x0, double *u0);

/* Measured fp64 issue rate of the device, in scalar operations per second (a fused multiply-add counts ONCE): the
 * compute-side ceiling bench.py reports next to the HBM roofline (SURVEY.md §8d "secondary ceiling").  The library is
 * built without FMA contraction, so its arithmetic runs at the mul / add rates. */
int ilqr_measure_fp64(int32_t device, double *fma_per_s, double *mul_per_s, double *add_per_s);

/* library build info: "ilqr_b200 <version> sm_100a ..." */
const char *ilqr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ILQR_B200_H_ */
