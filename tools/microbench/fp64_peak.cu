// FP64 pipe microbenchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) issue rates.
// Used once to choose the backward-sweep inner-product strategy (DESIGN.md "FP64 ridge").
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void dmma1688_kernel(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// LDS-broadcast-fed DFMA: each lane does R*C fmas per k with C broadcast LDS.128 loads (pairs)
template <int R, int C>
__global__ void lds_dfma_kernel(double* out, int iters) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-6;
    __syncthreads();
    int warp = threadIdx.x >> 5;
    double v[R];
    double acc[R][C];
#pragma unroll
    for (int r = 0; r < R; ++r) { v[r] = threadIdx.x * 1e-3 + r;
#pragma unroll
        for (int c = 0; c < C; ++c) acc[r][c] = 0; }
    const double* base = sm + (warp & 7) * 256;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int c = 0; c < C; c += 2) {
                double2 g = *reinterpret_cast<const double2*>(base + ((it & 1) * 128 + k * C + c));
#pragma unroll
                for (int r = 0; r < R; ++r) { acc[r][c] = fma(v[r], g.x, acc[r][c]); acc[r][c + 1] = fma(v[r], g.y, acc[r][c + 1]); }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) s += acc[r][c];
    if (s == 123.456) out[0] = s;
}

template <typename F>
float time_it(F launch, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount; int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device %s sms %d clock_max_khz %d\n", p.name, sms, clk_khz);
    double* out; CK(cudaMalloc(&out, 64));
    const int iters = 20000;
    for (int warps_per_sm : {4, 8, 16, 32}) {
        int threads = 128; int blocks = sms * warps_per_sm * 32 / threads;
        {
            float ms = time_it([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double flops = 2.0 * blocks * threads * (double)iters * 8;
            printf("DFMA ilp8 warps/SM %2d: %.3f ms  %.2f TFLOP/s  (%.1f DFMA/clk/SM @max clk)\n", warps_per_sm, ms, flops / ms * 1e-9,
                   flops / 2 / (ms * 1e-3) / sms / (clk_khz * 1e3));
        }
        {
            float ms = time_it([&] { dmma884_kernel<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double flops = 2.0 * 256 * (blocks * threads / 32) * (double)iters * 4;
            printf("DMMA m8n8k4 ilp4 warps/SM %2d: %.3f ms  %.2f TFLOP/s  (%.1f MAC/clk/SM)\n", warps_per_sm, ms, flops / ms * 1e-9,
                   flops / 2 / (ms * 1e-3) / sms / (clk_khz * 1e3));
        }
        {
            float ms = time_it([&] { dmma1688_kernel<4><<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); });
            double flops = 2.0 * 1024 * (blocks * threads / 32) * (double)(iters / 4) * 4;
            printf("DMMA m16n8k8 ilp4 warps/SM %2d: %.3f ms  %.2f TFLOP/s  (%.1f MAC/clk/SM)\n", warps_per_sm, ms, flops / ms * 1e-9,
                   flops / 2 / (ms * 1e-3) / sms / (clk_khz * 1e3));
        }
    }
    for (int warps_per_sm : {8, 16, 28}) {
        int threads = 128; int blocks = sms * warps_per_sm * 32 / threads;
        const int it2 = 4000;
        {
            float ms = time_it([&] { lds_dfma_kernel<1, 8><<<blocks, threads>>>(out, it2); });
            double flops = 2.0 * blocks * threads * (double)it2 * 8 * 1 * 8;
            printf("LDS-fed DFMA R1xC8 warps/SM %2d: %.3f ms %.2f TFLOP/s\n", warps_per_sm, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_it([&] { lds_dfma_kernel<2, 4><<<blocks, threads>>>(out, it2); });
            double flops = 2.0 * blocks * threads * (double)it2 * 8 * 2 * 4;
            printf("LDS-fed DFMA R2xC4 warps/SM %2d: %.3f ms %.2f TFLOP/s\n", warps_per_sm, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_it([&] { lds_dfma_kernel<3, 4><<<blocks, threads>>>(out, it2); });
            double flops = 2.0 * blocks * threads * (double)it2 * 8 * 3 * 4;
            printf("LDS-fed DFMA R3xC4 warps/SM %2d: %.3f ms %.2f TFLOP/s\n", warps_per_sm, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_it([&] { lds_dfma_kernel<4, 4><<<blocks, threads>>>(out, it2); });
            double flops = 2.0 * blocks * threads * (double)it2 * 8 * 4 * 4;
            printf("LDS-fed DFMA R4xC4 warps/SM %2d: %.3f ms %.2f TFLOP/s\n", warps_per_sm, ms, flops / ms * 1e-9);
        }
    }
    return 0;
}
