/* ilqr_phase_double_integrator.cu — the batch-lockstep phase kernels (ilqr_phases.cuh) of the built-in DoubleIntegrator
 * twin: f64 / f32, finite-difference / closed-form cost derivatives. */
#include "ilqr_variant.h"
#include "ilqr_phase_launch.cuh"

int ILQR_ENTRY(ilqr_phase_iterate_double_integrator)(ilqr_handle *h, int n_iters) {
  return ilqr::phase_iterate<ilqr::DoubleIntegrator>(h, n_iters);
}
