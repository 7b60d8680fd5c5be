// Compute-free kernels with level4's access pattern (NR planes read, NW planes written, 16 bytes per
// thread per plane): the HBM bandwidth this traffic mix can reach, as a yardstick for level4.
#include <cstdio>
#include <cuda_runtime.h>

template <int NR, int NW>
__global__ void __launch_bounds__(256) probe(const double* __restrict__ in, double* __restrict__ out, long long plane)
{
    const long long s = 2ll * (blockIdx.x * 256ll + threadIdx.x);
    if (s >= plane) return;
    double2 acc = make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < NR; i++) {
        const double2 v = __ldcs(reinterpret_cast<const double2*>(in + (long long)i * plane + s));
        acc.x += v.x; acc.y += v.y;
    }
#pragma unroll
    for (int i = 0; i < NW; i++) *reinterpret_cast<double2*>(out + (long long)i * plane + s) = make_double2(acc.x + i, acc.y - i);
}

template <int NR, int NW>
double run(const double* in, double* out, long long plane)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned grid = (unsigned)((plane / 2 + 255) / 256);
    probe<NR, NW><<<grid, 256>>>(in, out, plane);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        probe<NR, NW><<<grid, 256>>>(in, out, plane);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return (double)(NR + NW) * plane * 8 / (best * 1e-3) / 1e9;
}

int main()
{
    const long long plane = 16ll << 20;       // 16M doubles per plane = 128 MB (> L2)
    double *in, *out;
    cudaMalloc(&in, plane * 8 * 10);
    cudaMalloc(&out, plane * 8 * 5);
    cudaMemset(in, 0, plane * 8 * 10);
    printf("{\"copy_1r_1w_GBps\": %.0f, \"level4_inner_inner_10r_5w\": %.0f, \"level4_inner_tip_5r_5w\": %.0f, \"read_only_10r_0w\": %.0f, \"write_only_0r_5w\": %.0f}\n",
           run<1, 1>(in, out, plane), run<10, 5>(in, out, plane), run<5, 5>(in, out, plane), run<10, 0>(in, out, plane), run<0, 5>(in, out, plane));
    return cudaGetLastError() != cudaSuccess;
}
