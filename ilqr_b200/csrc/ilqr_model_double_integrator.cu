/* ilqr_model_double_integrator.cu — the solver kernels of the built-in DoubleIntegrator twin (models.cuh): f64 / f32, finite-difference /
 * closed-form cost derivatives, one / two trajectories per warp. */
#include "ilqr_variant.h"
#include "ilqr_launch.cuh"

int ILQR_ENTRY(ilqr_launch_double_integrator)(ilqr_handle *h, int op, int n_iters, double scalar) {
  return ilqr::launch_s<ilqr::DoubleIntegrator>(h, op, n_iters, scalar);
}
