// Cycle counts of pieces of the solver for ONE warp on sm_100a (experiment; not part of the product).
// nvcc -O3 -fmad=false -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ilqr_b200/csrc -o /tmp/pieces tools/ubench_pieces.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "boxqp.cuh"
using namespace ilqr;
__device__ double sink[8];
__global__ void k_boxqp(QPParams<double> p, double Q, double c, double lo, double hi, int n, long long *out, int active) {
  if ((int)threadIdx.x >= active) return;
  double x0 = 0.1 * threadIdx.x * 0;
  double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    QPScalar<double> r = box_qp_scalar<double>(p, Q + acc * 1e-30, c, x0, lo, hi);
    x0 = r.x * 0.5;
    acc += r.Hinv + r.result;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[0] = acc + x0;
}
__global__ void k_dyn(double dt, int n, long long *out, int active) {
  if ((int)threadIdx.x >= active) return;
  double x[4] = {0.1 + 0.01 * threadIdx.x, -0.2, 0.3, 0.1}, u[1] = {0.3}, mp[4] = {3.1415, 0, 0, 0};
  double cost = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    double x1[4];
    cost += Acrobot::cost(x, u, mp);
    integrate<Acrobot, double>(x, u, mp, dt, x1);
    for (int j = 0; j < 4; j++) x[j] = x1[j];
    u[0] = 0.3 - 0.1 * x[0];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[1] = x[0] + x[1] + x[2] + x[3] + cost;
}
__global__ void k_sincos3(int n, long long *out) {
  double a = 0.1 + threadIdx.x * 0.01, b = 0.2, c = 0.3;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    double sn[3], cs[3];
    sincos_det3(a, b, c, sn, cs);
    a = sn[0] + cs[1];
    b = sn[1] + cs[2];
    c = sn[2] + cs[0];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[2] = a + b + c;
}
int main() {
  long long *d, h;
  cudaMalloc(&d, 8);
  QPParams<double> p;
  p.max_iter = 100; p.min_grad = 1e-8; p.min_rel_improve = 1e-8; p.step_dec = 0.6; p.min_step = 1e-22; p.armijo = 0.1; p.clamp_tol = 1e-4;
  const int n = 2000;
  struct { double Q, c, lo, hi; const char *name; } cases[] = {
      {2.0, -1.0, -5, 5, "interior (2 iterations, result 5)"},
      {2.0, -30.0, -5, 5, "hits bound (step clamped, then result 6)"},
      {2.0, -30.0, -5, 0.0, "clamped at once? x0=0=hi, grad<0 -> result 6 at iteration 0"},
  };
  for (auto &cs : cases)
    for (int active : {32, 1}) {
      k_boxqp<<<1, 32>>>(p, cs.Q, cs.c, cs.lo, cs.hi, n, d, active);
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("box_qp_scalar %-60s lanes %2d: %7.1f cycles/call\n", cs.name, active, (double)h / n);
    }
  for (int active : {32, 11, 1}) {
    k_dyn<<<1, 32>>>(0.02, n, d, active);
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("cost + integrate (one rollout step without feedback) lanes %2d: %7.1f cycles/step\n", active, (double)h / n);
  }
  k_sincos3<<<1, 32>>>(n, d);
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("sincos_det3: %7.1f cycles/call\n", (double)h / n);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
