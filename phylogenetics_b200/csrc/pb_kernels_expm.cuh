// Batched matrix exponential P = expm(t Q): one CTA per (branch, category) matrix, every
// intermediate in shared memory.
//
// Restates Matrix.expm (lib/linear_algebra.ml:308-404) step for step: 1-norm (dlange 'O'),
// Pade 3/5/7/9 below the thresholds of :320-330, otherwise Pade-13 with scaling by
// 2^-ceil(log2(norm/5.4)) and repeated squaring (:351-401); the linear solve is LU with partial
// pivoting (dgesv: reciprocal-scaled multipliers as dgetf2, forward/backward substitution as
// dtrsm).  Callers in the reference: Site_evolution_model.Make (:31-32), Mutsel (:96-97),
// Rate_matrix.transition_probability_matrix (lib/rate_matrix.ml:132-138).
//
// Matrices are row-major n x n in global memory; in shared memory they are zero-padded to
// NP x NP, NP = n rounded up to 4, so the GEMM can run on 4 x 4 register tiles without guards.
#pragma once
#include <cuda_runtime.h>

namespace pbk {

namespace expm_detail {

__device__ __constant__ double kPade3[4] = {120., 60., 12., 1.};
__device__ __constant__ double kPade5[6] = {30240., 15120., 3360., 420., 30., 1.};
__device__ __constant__ double kPade7[8] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
__device__ __constant__ double kPade9[10] = {17643225600., 8821612800., 2075673600., 302702400., 30270240.,
                                             2162160., 110880., 3960., 90., 1.};
__device__ __constant__ double kPade13[14] = {64764752532480000., 32382376266240000., 7771770303897600.,
                                              1187353796428800., 129060195264000., 10559470521600.,
                                              670442572800., 33522128640., 1323241920., 40840800., 960960.,
                                              16380., 182., 1.};

// C = A . B on NP x NP padded tiles; k accumulated in ascending order (dgemm).  Ends with a barrier.
__device__ __forceinline__ void gemm(const double* __restrict__ A, const double* __restrict__ B,
                                     double* __restrict__ C, int n, int NP)
{
    const int nt = NP >> 2;
    for (int tile = threadIdx.x; tile < nt * nt; tile += blockDim.x) {
        const int ti = tile / nt, tj = tile - ti * nt;
        double acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = 0.0;
        const double* a = A + (size_t)(4 * ti) * NP;
        const double* b = B + 4 * tj;
        for (int k = 0; k < n; k++) {
            double av[4], bv[4];
#pragma unroll
            for (int r = 0; r < 4; r++) av[r] = a[r * NP + k];
#pragma unroll
            for (int c = 0; c < 4; c++) bv[c] = b[(size_t)k * NP + c];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) C[(size_t)(4 * ti + r) * NP + 4 * tj + c] = acc[r][c];
    }
    __syncthreads();
}

// Solve A X = B in place (X overwrites B), LU with partial pivoting on the n x n blocks.
__device__ __forceinline__ void gesv(double* __restrict__ A, double* __restrict__ B, int n, int NP, int* piv_sh)
{
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int k = 0; k < n; k++) {
        if (tid < 32) {                       // idamax over column k, rows k..n-1 (first maximum wins)
            double best = -1.0;
            int at = k;
            for (int i = k + tid; i < n; i += 32) {
                const double v = fabs(A[(size_t)i * NP + k]);
                if (v > best) { best = v; at = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_down_sync(0xffffffffu, best, o);
                const int oa = __shfl_down_sync(0xffffffffu, at, o);
                if (ob > best || (ob == best && oa < at)) { best = ob; at = oa; }
            }
            if (tid == 0) *piv_sh = at;
        }
        __syncthreads();
        const int p = *piv_sh;
        if (p != k) {
            for (int j = tid; j < 2 * n; j += nth) {
                double* M = (j < n) ? A : B;
                const int jj = (j < n) ? j : j - n;
                const double x = M[(size_t)k * NP + jj];
                M[(size_t)k * NP + jj] = M[(size_t)p * NP + jj];
                M[(size_t)p * NP + jj] = x;
            }
        }
        __syncthreads();
        const double rp = 1.0 / A[(size_t)k * NP + k];
        for (int i = k + 1 + tid; i < n; i += nth) A[(size_t)i * NP + k] *= rp;
        __syncthreads();
        const int rows = n - k - 1, cols = (n - k - 1) + n;
        for (int idx = tid; idx < rows * cols; idx += nth) {
            const int i = k + 1 + idx / cols, jc = idx % cols;
            const double l = A[(size_t)i * NP + k];
            if (jc < n - k - 1) {
                const int j = k + 1 + jc;
                A[(size_t)i * NP + j] -= l * A[(size_t)k * NP + j];
            } else {
                const int j = jc - (n - k - 1);
                B[(size_t)i * NP + j] -= l * B[(size_t)k * NP + j];
            }
        }
        __syncthreads();
    }
    for (int k = n - 1; k >= 0; k--) {
        const double d = A[(size_t)k * NP + k];
        for (int j = tid; j < n; j += nth) B[(size_t)k * NP + j] /= d;
        __syncthreads();
        for (int idx = tid; idx < k * n; idx += nth) {
            const int i = idx / n, j = idx - i * n;
            B[(size_t)i * NP + j] -= B[(size_t)k * NP + j] * A[(size_t)i * NP + k];
        }
        __syncthreads();
    }
}

}  // namespace expm_detail

// grid = (nmat, ncat); t = blen[b] * (rates ? rates[c] : 1).  Output layout [c][b][n*n] row-major.
__global__ void __launch_bounds__(256)
expm_batched(int n, int NP, int nmat, int skip, const double* __restrict__ Q, const double* __restrict__ blen,
             const double* __restrict__ rates, double* __restrict__ out)
{
    using namespace expm_detail;
    extern __shared__ double sh[];
    const int b = blockIdx.x, c = blockIdx.y;
    if (b == skip) return;
    const int tid = threadIdx.x, nth = blockDim.x;
    const size_t SZ = (size_t)NP * NP;
    double* X = sh;            // t*Q (scaled)
    double* X2 = X + SZ;
    double* X4 = X2 + SZ;
    double* X6 = X4 + SZ;
    double* T1 = X6 + SZ;
    double* T2 = T1 + SZ;
    double* T3 = T2 + SZ;
    __shared__ double colsum[64];
    __shared__ double norm_sh;
    __shared__ int piv_sh;
    const double t = blen[b] * (rates ? rates[c] : 1.0);
    double* dst = out + ((size_t)c * nmat + b) * n * n;

    if (n == 1) {                                   // :313-314
        if (tid == 0) dst[0] = exp(t * Q[0]);
        return;
    }
    for (int idx = tid; idx < (int)SZ; idx += nth) {
        const int i = idx / NP, j = idx - i * NP;
        X[idx] = (i < n && j < n) ? t * Q[i * n + j] : 0.0;     // Matrix.scal_mul t q
    }
    __syncthreads();
    if (tid < n) {                                  // norm1 = max column abs-sum (:211)
        double s = 0.0;
        for (int i = 0; i < n; i++) s += fabs(X[(size_t)i * NP + tid]);
        colsum[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        double m = colsum[0];
        for (int j = 1; j < n; j++) m = fmax(m, colsum[j]);
        norm_sh = m;
    }
    __syncthreads();
    const double norm = norm_sh;
    auto is_diag = [&](int idx) { const int i = idx / NP; return i < n && idx - i * NP == i; };

    if (norm <= 2.097847961257068) {                // :320
        const double* cf;
        int len;
        if (norm > 0.9504178996162932) { cf = kPade9; len = 10; }
        else if (norm > 0.2539398330063230) { cf = kPade7; len = 8; }
        else if (norm > 0.01495585217958292) { cf = kPade5; len = 6; }
        else { cf = kPade3; len = 4; }
        gemm(X, X, X2, n, NP);
        // power p in X4/X6 (ping-pong), U in T1, V in T2
        for (int idx = tid; idx < (int)SZ; idx += nth) {
            const double id = is_diag(idx) ? 1.0 : 0.0;
            X4[idx] = id;
            T1[idx] = cf[1] * id;
            T2[idx] = cf[0] * id;
        }
        __syncthreads();
        double* pw = X4;
        double* pn = X6;
        for (int i = 1; i <= len / 2 - 1; i++) {    // :337-343
            gemm(pw, X2, pn, n, NP);
            for (int idx = tid; idx < (int)SZ; idx += nth) {
                T1[idx] = fma(cf[2 * i + 1], pn[idx], T1[idx]);
                T2[idx] = fma(cf[2 * i], pn[idx], T2[idx]);
            }
            __syncthreads();
            double* sw = pw; pw = pn; pn = sw;
        }
        gemm(X, T1, T3, n, NP);                      // u = x . u
        for (int idx = tid; idx < (int)SZ; idx += nth) {
            const double u = T3[idx], v = T2[idx];
            T3[idx] = v - u;
            T2[idx] = v + u;
        }
        __syncthreads();
        gesv(T3, T2, n, NP, &piv_sh);
        for (int idx = tid; idx < n * n; idx += nth) dst[idx] = T2[(size_t)(idx / n) * NP + idx % n];
        return;
    }

    const double s = log(norm / 5.4) / log(2.0);     // :305-306, :352
    const double ts = ceil(s);
    if (s > 0.0) {
        const double f = exp2(-ts);                  // 2. ** (-. t), exact
        for (int idx = tid; idx < (int)SZ; idx += nth) X[idx] *= f;
        __syncthreads();
    }
    const double* cf = kPade13;
    gemm(X, X, X2, n, NP);
    gemm(X2, X2, X4, n, NP);
    gemm(X2, X4, X6, n, NP);
    for (int idx = tid; idx < (int)SZ; idx += nth) {       // :368-372
        double m = cf[9] * X2[idx];
        m = fma(cf[11], X4[idx], m);
        m = fma(cf[13], X6[idx], m);
        T1[idx] = m;
    }
    __syncthreads();
    gemm(X6, T1, T2, n, NP);
    for (int idx = tid; idx < (int)SZ; idx += nth) {       // :374-377
        double m = T2[idx];
        if (is_diag(idx)) m = cf[1] + m;
        m = fma(cf[3], X2[idx], m);
        m = fma(cf[5], X4[idx], m);
        m = fma(cf[7], X6[idx], m);
        T2[idx] = m;
    }
    __syncthreads();
    gemm(X, T2, T1, n, NP);                                // U in T1 (:378)
    for (int idx = tid; idx < (int)SZ; idx += nth) {       // :380-384
        double m = cf[8] * X2[idx];
        m = fma(cf[10], X4[idx], m);
        m = fma(cf[12], X6[idx], m);
        T2[idx] = m;
    }
    __syncthreads();
    gemm(X6, T2, T3, n, NP);
    for (int idx = tid; idx < (int)SZ; idx += nth) {       // :386-389, then a = v - u, b = v + u
        double m = T3[idx];
        if (is_diag(idx)) m = cf[0] + m;
        m = fma(cf[2], X2[idx], m);
        m = fma(cf[4], X4[idx], m);
        m = fma(cf[6], X6[idx], m);
        const double u = T1[idx];
        T3[idx] = m - u;
        T2[idx] = m + u;
    }
    __syncthreads();
    gesv(T3, T2, n, NP, &piv_sh);
    double* cur = T2;
    double* oth = T1;
    if (s > 0.0) {
        for (int r = 0; r < (int)ts; r++) {                // :396-400
            gemm(cur, cur, oth, n, NP);
            double* sw = cur; cur = oth; oth = sw;
        }
    }
    for (int idx = tid; idx < n * n; idx += nth) dst[idx] = cur[(size_t)(idx / n) * NP + idx % n];
}

// Returns 0, or -1 when 7 padded matrices do not fit the shared memory of an SM.
inline int launch_expm_batched(int n, int nmat, int ncat, int skip, const double* Q, const double* blen,
                               const double* rates, double* out, size_t smem_optin, cudaStream_t stream)
{
    const int NP = (n + 3) / 4 * 4;
    const size_t smem = (size_t)7 * NP * NP * sizeof(double);
    if (smem + 1024 > smem_optin) return -1;
    if (cudaFuncSetAttribute(expm_batched, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -1;
    expm_batched<<<dim3(nmat, ncat), 256, smem, stream>>>(n, NP, nmat, skip, Q, blen, rates, out);
    return 0;
}

}  // namespace pbk
