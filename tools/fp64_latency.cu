// Dependent-chain latencies of the FP64 pipe on this GPU (one warp, one block): DFMA, DMUL, DADD,
// DSETP predicate chains, MUFU.RCP64H, and a full 1.0/x division.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat(double* out, long long* cyc, double x0, double thr)
{
    const int N = 4096;
    double a = x0 + threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = fma(a, b, c);
    t1 = clock64();
    cyc[0] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = a * b;
    t1 = clock64();
    cyc[1] = t1 - t0;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = a + c;
    t1 = clock64();
    cyc[2] = t1 - t0;
    // 4 independent DFMA chains (throughput of one warp)
    double a1 = a + 1, a2 = a + 2, a3 = a + 3;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) { a = fma(a, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
    t1 = clock64();
    cyc[3] = t1 - t0;
    a += a1 + a2 + a3;
    // compare + select chain (DSETP -> FSEL x2 -> DSETP ...)
    double m = a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) m = (m > thr) ? m * b : m + c;
    t1 = clock64();
    cyc[4] = t1 - t0;
    // division chain
    double d = a;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 512; i++) d = 1.0 / (d + 1.5);
    t1 = clock64();
    cyc[5] = t1 - t0;
    out[threadIdx.x] = a + m + d;
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8);
    for (int r = 0; r < 2; r++) lat<<<1, 32>>>(out, cyc, 1.0, 0.5);
    long long h[8];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"dfma_chain_cycles\": %.2f, \"dmul_chain\": %.2f, \"dadd_chain\": %.2f, \"dfma_4chains_per_instr\": %.2f, \"cmp_select_mul_chain\": %.2f, \"div_plus_add_chain\": %.2f}\n",
           h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0, h[3] / (4.0 * 4096.0), h[4] / 4096.0, h[5] / 512.0);
    return cudaGetLastError() != cudaSuccess;
}
