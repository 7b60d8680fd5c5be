// Do DFMA and DMMA (mma.sync f64) share execution resources on B200?  Times the same kernel with
// (a) only a DFMA stream, (b) only a DMMA stream, (c) both interleaved in every warp.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_mix tools/fp64_mix.cu && tools/fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__device__ __forceinline__ void dfma(double& f, double a, double b)
{
    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f) : "d"(a), "d"(b));
}

template <int NF, int NM>     // per inner step: NF DFMA (2 issue cycles each per SMSP) and NM DMMA.8x8x4 (16 cycles each)
__global__ void __launch_bounds__(256) mix(double* out, int iters, double x)
{
    double f[8], m[8][2];
    for (int i = 0; i < 8; i++) { f[i] = threadIdx.x + i; m[i][0] = i; m[i][1] = -i; }
    const double a = x, b = 1.0 - x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
#pragma unroll
                for (int q = 0; q < NF; q++) dfma(f[(i + q) & 7], a, b);
                if (NM) dmma884(m[i], a, b);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += f[i] + m[i][0] + m[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NF, int NM>
float run(double* out, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<NF, NM><<<148 * 4, 256>>>(out, 10, 0.5);
    cudaEventRecord(e0);
    mix<NF, NM><<<148 * 4, 256>>>(out, iters, 0.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    double* out;
    cudaMalloc(&out, 148 * 4 * 256 * sizeof(double));
    const int iters = 4000;
    const double warps = 148.0 * 4 * 8, steps = (double)iters * 32;
    // 8 DFMA per DMMA.8x8x4: equal pipe time if each ran alone at its peak
    float tf = run<8, 0>(out, iters), tm = run<0, 1>(out, iters), tb = run<8, 1>(out, iters);
    printf("{\"dfma_only_ms\": %.3f, \"dmma_only_ms\": %.3f, \"both_ms\": %.3f, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, "
           "\"both_tflops\": %.2f, \"sum_of_parts_ms\": %.3f, \"verdict\": \"%s\"}\n",
           tf, tm, tb, warps * steps * 8 * 64 / tf / 1e9, warps * steps * 512 / tm / 1e9,
           warps * steps * (8 * 64 + 512) / tb / 1e9, tf + tm,
           tb > 0.85f * (tf + tm) ? "DFMA and DMMA share the FP64 pipe" : "DFMA and DMMA overlap");
    return 0;
}
