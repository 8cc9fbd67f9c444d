// DFMA latency / issue microbenchmark: W warps per SM sub-partition, C independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double* out, int iters, long long* cyc) {
  double a[C];
#pragma unroll
  for (int i = 0; i < C; ++i) a[i] = threadIdx.x * 1e-9 + i;
  const double m = 0.9999999, c = 1e-7;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < C; ++i) a[i] = fma(a[i], m, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < C; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int C>
void run(int warps_per_smsp) {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
  int iters = 2000, threads = 128 * warps_per_smsp;   // one CTA per SM: 4*W warps
  k<C><<<148, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
  k<C><<<148, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 8.0 * C);
  printf("warps/SMSP %d chains %d: %.2f cycles per DFMA per warp; SMSP DFMA/cycle %.3f\n", warps_per_smsp, C, per, warps_per_smsp / per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w = 1; w <= 4; ++w) { run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<6>(w); run<8>(w); }
  return 0;
}
