// Micro-benchmark: FP64 peaks of the device (DFMA and mma.sync DMMA m8n8k4 / m16n8k8 / m16n8k16),
// the denominators BASELINE.md asks the builder to measure.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void dmma884_kernel(double* out, int iters)
{
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int i = 0; i < iters; i++) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

__global__ void dmma1688_kernel(double* out, int iters)
{
    double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c3[4] = {0, 0, 0, 0};
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1.0 + threadIdx.x * 1e-4, b1 = b0 + 1;
#define MMA1688(c) asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1))
    for (int i = 0; i < iters; i++) { MMA1688(c0); MMA1688(c1); MMA1688(c2); MMA1688(c3); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c1[1] + c2[2] + c3[3];
}

template <typename K>
double run(K kernel, double flops_per_thread_iter, int iters, int blocks, int threads, double* out)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kernel<<<blocks, threads>>>(out, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return flops_per_thread_iter * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    double dfma = run(dfma_kernel, 16.0, iters, blocks, threads, out);
    // m8n8k4: 8*8*4*2 = 512 flop per warp-instruction = 16 per thread; 4 per iteration
    double d884 = run(dmma884_kernel, 64.0, iters, blocks, threads, out);
    // m16n8k8: 16*8*8*2 = 2048 flop per warp = 64 per thread; 4 per iteration
    double d1688 = run(dmma1688_kernel, 256.0, iters / 4, blocks, threads, out);
    printf("{\"device\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"dmma_m16n8k8_tflops\": %.2f}\n",
           p.name, p.multiProcessorCount, dfma, d884, d1688);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
