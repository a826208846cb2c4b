// Dependent-issue latencies that set the per-step critical path of the backward sweep (B200, sm_100a):
// DFMA chain, DADD chain, 1/x, rsqrt, sqrt, LDS->DFMA, SHFL.  One warp per SM, clock64() deltas.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *clk, double seed) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed + i * 1e-3;
  __syncthreads();
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0, t1;
  const int R = 256;
  // DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < R; ++i) x = fma(x, y, 1e-9);
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = (t1 - t0);
  // DADD
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < R; ++i) x = x + y;
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[1] = (t1 - t0);
  // reciprocal
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; ++i) x = 1.0 / (x + 1.5);
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[2] = (t1 - t0) * 4;
  // rsqrt
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; ++i) x = rsqrt(x + 1.5);
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[3] = (t1 - t0) * 4;
  // sqrt
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; ++i) x = sqrt(x + 1.5);
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[4] = (t1 - t0) * 4;
  // LDS (dependent address) -> value
  int idx = threadIdx.x & 7;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < R; ++i) idx = (int)sm[idx & 1023] & 1023;
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[5] = (t1 - t0);
  // SHFL double
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < R; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1e-9;
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[6] = (t1 - t0);
  // independent DFMA throughput for one warp (8 chains)
  double a[8];
  for (int j = 0; j < 8; ++j) a[j] = x + j;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fma(a[j], y, 1e-9);
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[7] = (t1 - t0) / 8;
  for (int j = 0; j < 8; ++j) x += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + idx;
}
int main() {
  double *out; long long *clk, h[8];
  cudaMalloc(&out, 148 * 128 * 8); cudaMalloc(&clk, 64);
  const char *names[8] = {"DFMA dependent", "DADD dependent", "1/x (+add)", "rsqrt (+add)", "sqrt (+add)", "LDS.64 dependent (+cvt)", "SHFL.f64 (+add)", "DFMA 8 indep chains (per instr)"};
  for (int warps = 1; warps <= 4; warps *= 2) {
    k<<<148, 32 * warps>>>(out, clk, 1.0); k<<<148, 32 * warps>>>(out, clk, 1.0);
    cudaDeviceSynchronize();
    cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
    printf("warps/SM=%d:", warps);
    for (int i = 0; i < 8; ++i) printf("  %s %.1f clk;", names[i], h[i] / 256.0);
    printf("\n");
  }
  return 0;
}
