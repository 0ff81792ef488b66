// Latency microbenchmarks for one warp on sm_100a: dependent chains of fp64 ops, division, sqrt, shared-memory round trips.
// nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ double sink;
template <int OP>
__global__ void chain(double a, double b, int n, long long *out, int active) {
  __shared__ double sm[64];
  double x = a + threadIdx.x;
  if ((int)threadIdx.x >= active) return;
  sm[threadIdx.x] = x;
  __syncwarp(__activemask());
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    if (OP == 0) x = x + b;
    if (OP == 1) x = x * b;
    if (OP == 2) x = fma(x, b, a);
    if (OP == 3) x = b / x;
    if (OP == 4) x = sqrt(x) + a;
    if (OP == 5) { sm[threadIdx.x] = x; __syncwarp(__activemask()); x = sm[(threadIdx.x + 1) & (active - 1)]; __syncwarp(__activemask()); }
    if (OP == 6) { x = __shfl_sync(__activemask(), x, (threadIdx.x + 1) & (active - 1)); }
    if (OP == 7) { float f = (float)x; f = f * 1.0001f + 0.5f; x = f; }
    if (OP == 8) { x = (x < b) ? x + a : x - a; }
    if (OP == 9) { x = __drcp_rn(x) + a; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink = x;
}
// throughput: nw warps each with 8 independent chains
template <int OP>
__global__ void thru(double a, double b, int n, long long *out) {
  double x[8];
  for (int k = 0; k < 8; k++) x[k] = a + threadIdx.x + k;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (OP == 0) x[k] = x[k] + b;
      if (OP == 1) x[k] = x[k] * b;
      if (OP == 2) x[k] = fma(x[k], b, a);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < 8; k++) s += x[k];
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink = s;
}
int main() {
  long long *d, h;
  cudaMalloc(&d, 8);
  const char *names[] = {"dadd", "dmul", "dfma", "ddiv", "dsqrt+dadd", "sts+sync+lds+sync", "shfl", "d2f,ffma,f2d", "dsetp+sel+dadd", "drcp+dadd"};
  const int n = 4096;
  for (int active : {32, 16, 1}) {
    printf("active lanes %d\n", active);
#define RUN(OP) chain<OP><<<1, 32>>>(1.5, 1.0000001, n, d, active); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("  %-22s %7.1f cycles/iter\n", names[OP], (double)h / n);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
  }
  for (int threads : {32, 128, 256, 512, 1024}) {
#define RUNT(OP) thru<OP><<<1, threads>>>(1.5, 1.0000001, n, d); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("  thru %-6s threads %4d: %6.2f cycles per warp-instr per SM-subpartition\n", names[OP], threads, (double)h / n / 8 / ((threads + 127) / 128));
    RUNT(0) RUNT(1) RUNT(2)
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
