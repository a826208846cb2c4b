// How many shared-memory wavefronts does a BROADCAST load cost when a warp holds k lane groups that read k
// different addresses (one per trajectory)?  Decides the lane-group size of the backward sweep.
// Throughput test: 16 warps/SM, each issuing back-to-back independent LDS; reports SM cycles per warp-level LDS.
#include <cstdio>
#include <cuda_runtime.h>
template <int G, int VEC>  // G lanes per group; VEC = 1 (LDS.64) or 2 (LDS.128)
__global__ void k(double *out, int stride_doubles, long long *clk) {
  extern __shared__ __align__(16) double sm[];
  for (int i = threadIdx.x; i < 6144; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double *base = sm + (lane / G) * stride_doubles + warp * 8;
  unsigned long long acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  const unsigned addr = (unsigned)__cvta_generic_to_shared(base);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      unsigned long long a0, a1, b0, b1, c0, c1, d0, d1;
      if (VEC == 2) {
        asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(a0), "=l"(a1) : "r"(addr + 16 * i));
        asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(b0), "=l"(b1) : "r"(addr + 16 * i + 16));
        asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(c0), "=l"(c1) : "r"(addr + 16 * i + 32));
        asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(d0), "=l"(d1) : "r"(addr + 16 * i + 48));
        acc0 ^= a0 ^ a1; acc1 ^= b0 ^ b1; acc2 ^= c0 ^ c1; acc3 ^= d0 ^ d1;
      } else {
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(a0) : "r"(addr + 8 * i));
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(b0) : "r"(addr + 8 * i + 8));
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(c0) : "r"(addr + 8 * i + 16));
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(d0) : "r"(addr + 8 * i + 24));
        acc0 ^= a0; acc1 ^= b0; acc2 ^= c0; acc3 ^= d0;
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = (double)(acc0 ^ acc1 ^ acc2 ^ acc3);
}
template <int G, int VEC>
void run(const char *name, int stride, double *out, long long *clk) {
  cudaFuncSetAttribute(k<G, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  k<G, VEC><<<148, 512, 49152>>>(out, stride, clk);
  k<G, VEC><<<148, 512, 49152>>>(out, stride, clk);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double loads_per_warp = 64.0 * 32 / (VEC == 2 ? 1 : 1);  // warp-level LDS instructions: 64*32 (VEC=1) or 64*32 (VEC=2, 4 per inner step of 4)
  const double n = (VEC == 2) ? 64.0 * 8 * 4 : 64.0 * 32;
  printf("%-34s stride %4d: %.2f SM-clk per warp-LDS per warp (16 warps => %.2f clk per LDS at SM level)\n", name, stride, h / n, h / n / 16.0);
}
int main() {
  double *out; long long *clk;
  cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&clk, 8);
  for (int stride : {614, 616, 624, 640}) {
    run<32, 1>("LDS.64  1 address/warp", stride, out, clk);
    run<16, 1>("LDS.64  2 addresses/warp (G=16)", stride, out, clk);
    run<8, 1>("LDS.64  4 addresses/warp (G=8)", stride, out, clk);
    run<4, 1>("LDS.64  8 addresses/warp (G=4)", stride, out, clk);
    run<32, 2>("LDS.128 1 address/warp", stride, out, clk);
    run<16, 2>("LDS.128 2 addresses/warp (G=16)", stride, out, clk);
    run<8, 2>("LDS.128 4 addresses/warp (G=8)", stride, out, clk);
    run<4, 2>("LDS.128 8 addresses/warp (G=4)", stride, out, clk);
  }
  return 0;
}
