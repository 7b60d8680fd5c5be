// FP64 tensor-core peak vs. independent accumulator chains, for mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k884(double* out, int iters)
{
    double c[CH][2];
    for (int i = 0; i < CH; i++) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k1684(double* out, int iters)
{
    double c[CH][4];
    for (int i = 0; i < CH; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    double a0 = threadIdx.x * 1e-3 + 1.0, a1 = a0 + 1, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k1688(double* out, int iters)
{
    double c[CH][4];
    for (int i = 0; i < CH; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    double a0 = threadIdx.x * 1e-3 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1.0 + threadIdx.x * 1e-4, b1 = b0 + 1;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double s = 0;
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k16816(double* out, int iters)
{
    double c[CH][4];
    for (int i = 0; i < CH; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; i++) b[i] = 1.0 + threadIdx.x * 1e-4 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
double run(K kernel, double flop_per_warp_instr, int chains, int iters, int blocks, int threads, double* out)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kernel<<<blocks, threads>>>(out, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return flop_per_warp_instr * chains * iters * (double)blocks * (threads / 32) / (best * 1e-3) / 1e12;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 8 * 512);
    const int it = 4000;
    for (int wps = 4; wps <= 16; wps *= 2) {       // warps per SM
        const int blocks = p.multiProcessorCount, threads = wps * 32;
        printf("{\"warps_per_sm\": %d", wps);
        printf(", \"m8n8k4_ch1\": %.1f", run(k884<1>, 512.0, 1, it * 4, blocks, threads, out));
        printf(", \"m8n8k4_ch4\": %.1f", run(k884<4>, 512.0, 4, it, blocks, threads, out));
        printf(", \"m8n8k4_ch16\": %.1f", run(k884<16>, 512.0, 16, it / 4, blocks, threads, out));
        printf(", \"m16n8k4_ch4\": %.1f", run(k1684<4>, 1024.0, 4, it, blocks, threads, out));
        printf(", \"m16n8k8_ch1\": %.1f", run(k1688<1>, 2048.0, 1, it, blocks, threads, out));
        printf(", \"m16n8k8_ch4\": %.1f", run(k1688<4>, 2048.0, 4, it / 2, blocks, threads, out));
        printf(", \"m16n8k16_ch4\": %.1f", run(k16816<4>, 4096.0, 4, it / 4, blocks, threads, out));
        printf("}\n");
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
