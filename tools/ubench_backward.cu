// One warp, one trajectory: cycles of Core::derivative_sweep / backward_pass / rollout_candidates measured with
// clock64 inside the kernel (experiment; not part of the product).
// nvcc -O3 -fmad=false -std=c++17 -gencode arch=compute_100a,code=sm_100a [-DUB_MINB=4] -I ilqr_b200/csrc -I include -o /tmp/ubw tools/ubench_backward.cu
// UB_MINB = resident CTAs per SM in __launch_bounds__ (7: 72 registers, 4: 126); argv[1] = 0 disables the bulk (TMA) tile copy.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../include/ilqr_b200.h"
#include "ilqr_core.cuh"
#include "params.h"
using namespace ilqr;
#ifndef UB_G
#define UB_G 32
#endif
#ifndef UB_MINB
#define UB_MINB 7
#endif
using Ex = WarpExec<4, 1, double, UB_G>;
using CoreT = Core<Acrobot, double, kCostAnalytic, Ex>;
using Sc = CoreT::Sc;
struct Args {
  SolveParams<double> P;
  TrajPtrs<double> tr;
  SlotPtrs<double> sl;
  long long *out;
};
__global__ void __launch_bounds__(128, UB_MINB) k(const __grid_constant__ Args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  Sc &sc = *reinterpret_cast<Sc *>(smem);
  SlotPtrs<double> sl = a.sl;
  sl.gterm = reinterpret_cast<double *>(smem + sizeof(Sc));
  Ex ex;
  ex.lane = threadIdx.x & (UB_G - 1);
  ex.mask = UB_G == 32 ? 0xffffffffu : (0xffffu << ((threadIdx.x & 31) & ~(UB_G - 1)));
  ex.init_barrier(reinterpret_cast<unsigned long long *>(smem + ((sizeof(Sc) + a.P.T * 8 + 15) & ~(size_t)15)));
  CoreT core(a.P, sc, ex, a.tr, sl);
  core.load_state();
  long long t[8];
  for (int rep = 0; rep < 2; rep++) {
    t[0] = clock64();
    core.derivative_sweep();
    t[1] = clock64();
    core.backward_pass(1.0);
    t[2] = clock64();
    core.gradient_norm_from_terms();
    t[3] = clock64();
    core.rollout_candidates();
    t[4] = clock64();
    core.commit_candidate(3);
    t[5] = clock64();
  }
  if (threadIdx.x == 0)
    for (int i = 0; i < 5; i++) a.out[i] = t[i + 1] - t[i];
}
int main(int argc, char **argv) {
  const int T = 200, N = 4, M = 1;
  ilqr_desc d;
  memset(&d, 0, sizeof(d));
  default_params(&d.params);
  d.model_id = ILQR_MODEL_ACROBOT;
  d.T = T;
  d.B = 1;
  d.dt = 0.02;
  Args a;
  if (make_solve_params<double>(d, &a.P) != 0) return 1;
  std::vector<double> xs((T + 1) * N), us(T), K(T * N, 0.0), kk(T, 0.0);
  srand(1);
  double x[4] = {0.3, -0.2, 0.1, 0.05};
  for (int t = 0; t <= T; t++) { /* a plausible trajectory: an open-loop roll-out */
    for (int i = 0; i < 4; i++) xs[t * 4 + i] = x[i];
    if (t == T) break;
    us[t] = 0.5 * (2.0 * rand() / RAND_MAX - 1.0);
    double x1[4], mp[4] = {3.1415, 0, 0, 0};
    integrate<Acrobot, double>(x, &us[t], mp, 0.02, x1);
    for (int i = 0; i < 4; i++) x[i] = x1[i];
  }
  auto dev = [](const void *h, size_t bytes) { void *p; cudaMalloc(&p, bytes); if (h) cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice); else cudaMemset(p, 0, bytes); return p; };
  a.tr.x0 = (double *)dev(xs.data(), 32);
  a.tr.xs = (double *)dev(xs.data(), xs.size() * 8);
  a.tr.us = (double *)dev(us.data(), us.size() * 8);
  a.tr.K = (double *)dev(K.data(), K.size() * 8);
  a.tr.k = (double *)dev(kk.data(), kk.size() * 8);
  a.tr.Vx0 = (double *)dev(nullptr, 32);
  a.tr.Vxx0 = (double *)dev(nullptr, 128);
  a.tr.st = (TrajState<double> *)dev(nullptr, sizeof(TrajState<double>));
  a.sl.F = (double *)dev(nullptr, T * 5 * 4 * 8);
  a.sl.C = nullptr;
  a.sl.cand_x = (double *)dev(nullptr, 11 * T * 4 * 8);
  a.sl.cand_u = (double *)dev(nullptr, 11 * T * 8);
  a.P.bulk_f = argc > 1 ? atoi(argv[1]) : 1;
  a.out = (long long *)dev(nullptr, 64);
  const size_t smem = ((sizeof(Sc) + T * 8 + 15) & ~(size_t)15) + 16;
  k<<<1, 32, smem>>>(a);
  long long h[8];
  cudaMemcpy(h, a.out, 64, cudaMemcpyDeviceToHost);
  const char *names[] = {"derivative_sweep", "backward_pass", "gradient_norm", "rollout_candidates", "commit"};
  for (int i = 0; i < 5; i++) printf("%-20s %9lld cycles  (%7.1f per timestep)\n", names[i], h[i], (double)h[i] / T);
  std::vector<double> kout(T);
  cudaMemcpy(kout.data(), a.tr.k, T * 8, cudaMemcpyDeviceToHost);
  printf("k[0]=%.6g k[T-1]=%.6g  %s\n", kout[0], kout[T - 1], cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
