// K2, level-scheduled, dense alphabets (20 amino acids, 61 sense codons) on the FP64 tensor
// cores: parent[i][site] = sum_j P[i][j] child[j][site] is a GEMM with M = parent states,
// K = child states, N = sites, issued as mma.sync.aligned.m16n8k8 .f64 tiles (SASS: DMMA.16x8x8;
// tcgen05.mma has no FP64 kind).  Measured on B200 (tools/fp64_peak2.cu,
// profiles/r01_fp64_peak_dmma_shapes.jsonl): every mma.sync .f64 shape lowers to DMMA.8x8x4 and
// sustains 36.8 TFLOP/s, the same as the DFMA pipe's 37 TFLOP/s peak — the tensor path saves
// issue slots and register-file bandwidth, not pipe time.
//
// Restates the node step of Phylo_ctmc.pruning (lib/phylo_ctmc.ml:91-95): one mat-vec per
// inner child (a column look-up for a leaf child), children folded left to right with SV.mul,
// i.e. element-wise product followed by SV.shift (:33-40, :76-78).
//
// One warp owns 8 adjacent sites (the N of the MMA) at a time:
//   A fragments  <- the child's P, zero-padded to MP x KP, staged in shared memory
//                   (row stride KP+4 doubles: conflict-free 8-byte fragment loads)
//   B fragments  <- the child's CLV planes straight from HBM: 8 sites x 8 bytes = 64-byte
//                   segments per state plane (two full sectors)
//   C fragments  -> the node's CLV, rescaled in registers; min/max over states by
//                   warp shuffles across the 8 row groups; stored as 16-byte pairs of sites.
#pragma once
#include "pb_kernels.cuh"

namespace pbk {

__device__ __forceinline__ void dmma_16x8x8(double (&c)[4], double a0, double a1, double a2, double a3,
                                            double b0, double b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
}

template <int N>
struct DmmaShape {
    static constexpr int MP = (N + 15) / 16 * 16;   // padded parent states
    static constexpr int KP = (N + 7) / 8 * 8;      // padded child states
    static constexpr int LD = KP + 4;               // smem row stride (doubles)
    static constexpr int MT = MP / 16;
    static constexpr int KS = KP / 8;
};

// SV.shift on a C-fragment tile: column j (j = 0, 1) of the thread's two sites; rows
// mt*16 + g (+8).  All 8 row groups of a column take the same decision.
template <int N>
__device__ __forceinline__ void shift_frag(double (&out)[DmmaShape<N>::MT][4], double (&car)[2], int g, double thr)
{
    constexpr int MT = DmmaShape<N>::MT;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        double mn = INFINITY, mx = -INFINITY;
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            if (mt * 16 + g < N) { mn = fmin(mn, out[mt][j]); mx = fmax(mx, out[mt][j]); }
            if (mt * 16 + g + 8 < N) { mn = fmin(mn, out[mt][2 + j]); mx = fmax(mx, out[mt][2 + j]); }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (!(mn > thr)) {
            const double inv = 1.0 / mx;
#pragma unroll
            for (int mt = 0; mt < MT; mt++) { out[mt][j] = inv * out[mt][j]; out[mt][2 + j] = inv * out[mt][2 + j]; }
            car[j] += log(mx);
        }
    }
}

template <int N>
__global__ void __launch_bounds__(256)
level_dmma(const LevelNode* __restrict__ nodes, const LevelChild* __restrict__ children, int node_base,
           const double* __restrict__ P, int nnodes, int ncat,
           const uint8_t* __restrict__ tips, int64_t tips_ld,
           double* __restrict__ pool, int64_t slot_stride, int64_t spad,
           const double* __restrict__ pi, double thr, int shift_at_root,
           double* __restrict__ site_cat, int store_root, int tiles_per_block, int seq_count)
{
    using S = DmmaShape<N>;
    constexpr int MP = S::MP, LD = S::LD, MT = S::MT, KS = S::KS;
    extern __shared__ double shraw[];        // pi[MP] then [nchild][MP*LD]
    double* shPi = shraw;
    double* sh = shraw + MP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int64_t ntiles = spad / 8;
    // seq_count == 0: one node per launch row (blockIdx.y), blocks split the sites once.
    // seq_count  > 0: persistent blocks; each takes groups of tiles_per_block site tiles and walks ALL
    //                 seq_count nodes (post-order) for a group before moving on, so the group's
    //                 conditional likelihoods are produced and consumed while still in L2.
    const bool seq = seq_count > 0;
    const int nloop = seq ? seq_count : 1;
    const int64_t ngroups = seq ? (ntiles + tiles_per_block - 1) / tiles_per_block : 1;
    if (tid < MP) shPi[tid] = (tid < N) ? pi[tid] : 0.0;

    for (int64_t grp = seq ? blockIdx.x : 0; grp < ngroups; grp += (seq ? gridDim.x : 1)) {
    const int64_t tile_begin = (seq ? grp : (int64_t)blockIdx.x) * tiles_per_block;
    const int64_t tile_end = min(tile_begin + tiles_per_block, ntiles);
    for (int ni = 0; ni < nloop; ni++) {
    const LevelNode nd = nodes[node_base + (seq ? ni : (int)blockIdx.y)];
    const LevelChild* ch = children + nd.child_off;

    for (int c = 0; c < ncat; c++) {
        __syncthreads();
        for (int idx = tid; idx < nd.nchild * MP * LD; idx += 256) {
            const int k = idx / (MP * LD), r = idx - k * (MP * LD);
            const int i = r / LD, j = r - i * LD;
            sh[idx] = (i < N && j < N) ? P[((size_t)c * nnodes + ch[k].pnode) * N * N + i * N + j] : 0.0;
        }
        __syncthreads();

        for (int64_t tile = tile_begin + warp; tile < tile_end; tile += 8) {
            const int64_t s0 = tile * 8;
            double out[MT][4];
            double car[2] = {0.0, 0.0};
            for (int k = 0; k < nd.nchild; k++) {
                const LevelChild cd = ch[k];
                const double* Pm = sh + (size_t)k * MP * LD;
                double cf[MT][4];
                double tc[2] = {0.0, 0.0};
                if (cd.kind == 0) {
                    const uchar2 st = *reinterpret_cast<const uchar2*>(tips + (int64_t)cd.src * tips_ld + s0 + tig * 2);
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        const double* r0 = Pm + (mt * 16 + g) * LD;
                        const double* r1 = r0 + 8 * LD;
                        cf[mt][0] = r0[st.x]; cf[mt][1] = r0[st.y];
                        cf[mt][2] = r1[st.x]; cf[mt][3] = r1[st.y];
                    }
                } else {
                    const double* src = pool + (int64_t)cd.src * slot_stride + (int64_t)c * (N + 1) * spad + s0;
                    double b[KS][2];
#pragma unroll
                    for (int ks = 0; ks < KS; ks++) {
                        const int k0 = ks * 8 + tig, k1 = k0 + 4;
                        b[ks][0] = (k0 < N) ? src[(int64_t)k0 * spad + g] : 0.0;
                        b[ks][1] = (k1 < N) ? src[(int64_t)k1 * spad + g] : 0.0;
                    }
                    const double2 q = *reinterpret_cast<const double2*>(src + (int64_t)N * spad + tig * 2);
                    tc[0] = q.x; tc[1] = q.y;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) cf[mt][0] = cf[mt][1] = cf[mt][2] = cf[mt][3] = 0.0;
                    // k-step outermost: the MT accumulator tiles are independent DMMA chains in flight
#pragma unroll
                    for (int ks = 0; ks < KS; ks++) {
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) {
                            const double* r0 = Pm + (mt * 16 + g) * LD + tig + ks * 8;
                            const double* r1 = r0 + 8 * LD;
                            dmma_16x8x8(cf[mt], r0[0], r1[0], r0[4], r1[4], b[ks][0], b[ks][1]);
                        }
                    }
                }
                if (k == 0) {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int e = 0; e < 4; e++) out[mt][e] = cf[mt][e];
                    car[0] = tc[0]; car[1] = tc[1];
                } else {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int e = 0; e < 4; e++) out[mt][e] = out[mt][e] * cf[mt][e];
                    car[0] = car[0] + tc[0]; car[1] = car[1] + tc[1];
                    shift_frag<N>(out, car, g, thr);
                }
            }
            if (!nd.is_root || store_root) {
                double* dst = pool + (int64_t)nd.out_slot * slot_stride + (int64_t)c * (N + 1) * spad + s0 + tig * 2;
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const int r0 = mt * 16 + g, r1 = r0 + 8;
                    if (r0 < N) *reinterpret_cast<double2*>(dst + (int64_t)r0 * spad) = make_double2(out[mt][0], out[mt][1]);
                    if (r1 < N) *reinterpret_cast<double2*>(dst + (int64_t)r1 * spad) = make_double2(out[mt][2], out[mt][3]);
                }
                if (g == 0) *reinterpret_cast<double2*>(dst + (int64_t)N * spad) = make_double2(car[0], car[1]);
            }
            if (nd.is_root) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const double p0 = shPi[mt * 16 + g], p1 = shPi[mt * 16 + g + 8];
                    out[mt][0] *= p0; out[mt][1] *= p0; out[mt][2] *= p1; out[mt][3] *= p1;
                }
                if (shift_at_root) shift_frag<N>(out, car, g, thr);
                double sum[2] = {0.0, 0.0};
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {      // padded rows hold exact zeros
                    sum[0] += out[mt][0] + out[mt][2];
                    sum[1] += out[mt][1] + out[mt][3];
                }
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    sum[0] += __shfl_xor_sync(0xffffffffu, sum[0], o);
                    sum[1] += __shfl_xor_sync(0xffffffffu, sum[1], o);
                }
                if (g == 0)
                    *reinterpret_cast<double2*>(site_cat + (int64_t)c * spad + s0 + tig * 2) =
                        make_double2(log(sum[0]) + car[0], log(sum[1]) + car[1]);
            }
        }
    }
    }   // nodes
    }   // groups
}

// --------------------------------------------------------------------------
// level_dmma2: the same node step, software-pipelined.  Every warp owns a private double buffer in
// shared memory and streams the CLV tiles of its NEXT 8 sites with cp.async (16-byte chunks, no
// register staging) while the tensor cores work on the current tile, so the DMMAs never wait on
// HBM.  Rows of a staged tile are 64 bytes (8 sites); 16-byte chunk c of row k is stored at chunk
// c ^ (k & 3), which makes the 8-byte B-fragment loads bank-conflict free.  Arity <= 2.
// --------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

template <int N>
__global__ void __launch_bounds__(256, (N <= 32 ? 2 : 1))
level_dmma2(const LevelNode* __restrict__ nodes, const LevelChild* __restrict__ children, int node_base,
            const double* __restrict__ P, int nnodes, int ncat,
            const uint8_t* __restrict__ tips, int64_t tips_ld,
            double* __restrict__ pool, int64_t slot_stride, int64_t spad,
            const double* __restrict__ pi, double thr, int shift_at_root,
            double* __restrict__ site_cat, int store_root, int tiles_per_block)
{
    using S = DmmaShape<N>;
    constexpr int MP = S::MP, LD = S::LD, MT = S::MT, KS = S::KS;
    constexpr int XROWS = N + 1;                 // state planes + the carry plane
    constexpr int XSZ = XROWS * 8;               // doubles per staged child tile
    extern __shared__ double shraw[];            // pi[MP] | P[2][MP*LD] | X[8 warps][2 stages][2 children][XSZ] | tips
    double* shPi = shraw;
    double* sh = shraw + MP;
    double* shX = sh + 2 * MP * LD;
    unsigned long long* shT = reinterpret_cast<unsigned long long*>(shX + (size_t)8 * 4 * XSZ);   // [8][2][2] x 8 tip bytes
    const LevelNode nd = nodes[node_base + blockIdx.y];
    const LevelChild* ch = children + nd.child_off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int64_t ntiles = spad / 8;
    const int64_t tile_begin = (int64_t)blockIdx.x * tiles_per_block;
    const int64_t tile_end = min(tile_begin + tiles_per_block, ntiles);
    double* myX = shX + (size_t)warp * 4 * XSZ;
    unsigned long long* myT = shT + warp * 4;
    if (tid < MP) shPi[tid] = (tid < N) ? pi[tid] : 0.0;
    const LevelChild cd0 = ch[0];
    const LevelChild cd1 = ch[nd.nchild > 1 ? 1 : 0];

    for (int c = 0; c < ncat; c++) {
        __syncthreads();
        for (int idx = tid; idx < nd.nchild * MP * LD; idx += 256) {
            const int k = idx / (MP * LD), r = idx - k * (MP * LD);
            const int i = r / LD, j = r - i * LD;
            sh[idx] = (i < N && j < N) ? P[((size_t)c * nnodes + ch[k].pnode) * N * N + i * N + j] : 0.0;
        }
        __syncthreads();

        // stage the CLV planes (and carry) of every inner child for one tile; tips go to registers
        auto prefetch = [&](int64_t tile, int stage) {
            const int64_t s0 = tile * 8;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                if (k >= nd.nchild) break;
                const LevelChild cd = k ? cd1 : cd0;
                if (cd.kind == 0) {            // the tile's 8 tip bytes, asynchronously like the CLV planes
                    if (lane == 0) {
                        const unsigned d = (unsigned)__cvta_generic_to_shared(myT + stage * 2 + k);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(tips + (int64_t)cd.src * tips_ld + s0) : "memory");
                    }
                } else {
                    const double* src = pool + (int64_t)cd.src * slot_stride + (int64_t)c * (N + 1) * spad + s0;
                    double* dst = myX + (size_t)(stage * 2 + k) * XSZ;
                    for (int idx = lane; idx < XROWS * 4; idx += 32) {
                        const int row = idx >> 2, q = idx & 3;
                        cp_async16(dst + row * 8 + ((q ^ (row & 3)) << 1), src + (int64_t)row * spad + (q << 1));
                    }
                }
            }
            cp_async_commit();
        };

        int64_t tile = tile_begin + warp;
        if (tile < tile_end) prefetch(tile, 0);
        int stage = 0;
        for (; tile < tile_end; tile += 8, stage ^= 1) {
            const bool more = tile + 8 < tile_end;
            if (more) { prefetch(tile + 8, stage ^ 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncwarp();
            const int64_t s0 = tile * 8;
            double out[MT][4];
            double car[2] = {0.0, 0.0};
#pragma unroll
            for (int k = 0; k < 2; k++) {
                if (k >= nd.nchild) break;
                const LevelChild cd = k ? cd1 : cd0;
                const double* Pm = sh + (size_t)k * MP * LD;
                double cf[MT][4];
                double tc[2] = {0.0, 0.0};
                if (cd.kind == 0) {
                    const unsigned char* tb = reinterpret_cast<const unsigned char*>(myT + stage * 2 + k) + tig * 2;
                    const uchar2 st = make_uchar2(tb[0], tb[1]);
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        const double* r0 = Pm + (mt * 16 + g) * LD;
                        const double* r1 = r0 + 8 * LD;
                        cf[mt][0] = r0[st.x]; cf[mt][1] = r0[st.y];
                        cf[mt][2] = r1[st.x]; cf[mt][3] = r1[st.y];
                    }
                } else {
                    const double* X = myX + (size_t)(stage * 2 + k) * XSZ;
                    double b[KS][2];
#pragma unroll
                    for (int ks = 0; ks < KS; ks++) {
                        const int k0 = ks * 8 + tig, k1 = k0 + 4;         // (k0 & 3) == (k1 & 3) == tig
                        const int off = (((g >> 1) ^ tig) << 1) + (g & 1);
                        b[ks][0] = (k0 < N) ? X[k0 * 8 + off] : 0.0;
                        b[ks][1] = (k1 < N) ? X[k1 * 8 + off] : 0.0;
                    }
                    const double2 q = *reinterpret_cast<const double2*>(X + N * 8 + ((tig ^ (N & 3)) << 1));
                    tc[0] = q.x; tc[1] = q.y;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) cf[mt][0] = cf[mt][1] = cf[mt][2] = cf[mt][3] = 0.0;
                    // k-step outermost: the MT accumulator tiles are independent DMMA chains in flight
#pragma unroll
                    for (int ks = 0; ks < KS; ks++) {
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) {
                            const double* r0 = Pm + (mt * 16 + g) * LD + tig + ks * 8;
                            const double* r1 = r0 + 8 * LD;
                            dmma_16x8x8(cf[mt], r0[0], r1[0], r0[4], r1[4], b[ks][0], b[ks][1]);
                        }
                    }
                }
                if (k == 0) {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int e = 0; e < 4; e++) out[mt][e] = cf[mt][e];
                    car[0] = tc[0]; car[1] = tc[1];
                } else {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int e = 0; e < 4; e++) out[mt][e] = out[mt][e] * cf[mt][e];
                    car[0] = car[0] + tc[0]; car[1] = car[1] + tc[1];
                    shift_frag<N>(out, car, g, thr);
                }
            }
            __syncwarp();                       // every lane is done reading this stage before it is refilled
            if (!nd.is_root || store_root) {
                double* dst = pool + (int64_t)nd.out_slot * slot_stride + (int64_t)c * (N + 1) * spad + s0 + tig * 2;
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const int r0 = mt * 16 + g, r1 = r0 + 8;
                    if (r0 < N) *reinterpret_cast<double2*>(dst + (int64_t)r0 * spad) = make_double2(out[mt][0], out[mt][1]);
                    if (r1 < N) *reinterpret_cast<double2*>(dst + (int64_t)r1 * spad) = make_double2(out[mt][2], out[mt][3]);
                }
                if (g == 0) *reinterpret_cast<double2*>(dst + (int64_t)N * spad) = make_double2(car[0], car[1]);
            }
            if (nd.is_root) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const double p0 = shPi[mt * 16 + g], p1 = shPi[mt * 16 + g + 8];
                    out[mt][0] *= p0; out[mt][1] *= p0; out[mt][2] *= p1; out[mt][3] *= p1;
                }
                if (shift_at_root) shift_frag<N>(out, car, g, thr);
                double sum[2] = {0.0, 0.0};
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    sum[0] += out[mt][0] + out[mt][2];
                    sum[1] += out[mt][1] + out[mt][3];
                }
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    sum[0] += __shfl_xor_sync(0xffffffffu, sum[0], o);
                    sum[1] += __shfl_xor_sync(0xffffffffu, sum[1], o);
                }
                if (g == 0)
                    *reinterpret_cast<double2*>(site_cat + (int64_t)c * spad + s0 + tig * 2) =
                        make_double2(log(sum[0]) + car[0], log(sum[1]) + car[1]);
            }
        }
    }
}

template <int N>
int launch_level_dmma2_n(const LevelNode* nodes, const LevelChild* children, int first, int count,
                         const double* P, int nnodes, int ncat, const uint8_t* tips, int64_t tips_ld, double* pool,
                         int64_t slot_stride, int64_t spad, const double* pi, double thr, int shift_at_root,
                         double* site_cat, int store_root, int sm_count, size_t smem_optin, cudaStream_t stream)
{
    using S = DmmaShape<N>;
    const size_t smem = ((size_t)S::MP + 2 * S::MP * S::LD + (size_t)8 * 4 * (N + 1) * 8 + 8 * 4) * sizeof(double);
    if (smem > smem_optin) return -2;
    auto kern = level_dmma2<N>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem) != cudaSuccess || occ < 1) return -1;
    const int64_t ntiles = spad / 8;
    int64_t want_blocks = std::max<int64_t>(1, (int64_t)sm_count * occ / count);
    int64_t tpb = (ntiles + want_blocks - 1) / want_blocks;
    tpb = std::max<int64_t>(8, (tpb + 7) / 8 * 8);
    const int64_t gx = (ntiles + tpb - 1) / tpb;
    kern<<<dim3((unsigned)gx, count), 256, smem, stream>>>(nodes, children, first, P, nnodes, ncat, tips, tips_ld, pool,
                                                           slot_stride, spad, pi, thr, shift_at_root, site_cat,
                                                           store_root, (int)tpb);
    return 0;
}

inline int launch_level_dmma2(int n, const LevelNode* nodes, const LevelChild* children, int first, int count,
                              const double* P, int nnodes, int ncat, const uint8_t* tips, int64_t tips_ld,
                              double* pool, int64_t slot_stride, int64_t spad, const double* pi, double thr,
                              int shift_at_root, double* site_cat, int store_root, int sm_count, size_t smem_optin,
                              cudaStream_t stream)
{
    if (n == 20)
        return launch_level_dmma2_n<20>(nodes, children, first, count, P, nnodes, ncat, tips, tips_ld, pool, slot_stride,
                                        spad, pi, thr, shift_at_root, site_cat, store_root, sm_count, smem_optin, stream);
    if (n == 61)
        return launch_level_dmma2_n<61>(nodes, children, first, count, P, nnodes, ncat, tips, tips_ld, pool, slot_stride,
                                        spad, pi, thr, shift_at_root, site_cat, store_root, sm_count, smem_optin, stream);
    return -1;
}

inline bool dmma_supported(int n) { return n == 20 || n == 61; }

template <int N>
int launch_level_dmma_n(const LevelNode* nodes, const LevelChild* children, int first, int count, int max_arity,
                        const double* P, int nnodes, int ncat, const uint8_t* tips, int64_t tips_ld, double* pool,
                        int64_t slot_stride, int64_t spad, const double* pi, double thr, int shift_at_root,
                        double* site_cat, int store_root, int sm_count, size_t smem_optin, bool sequential,
                        int seq_tiles, cudaStream_t stream)
{
    using S = DmmaShape<N>;
    const size_t smem = ((size_t)max_arity * S::MP * S::LD + S::MP) * sizeof(double);
    if (smem > smem_optin) return -2;
    auto kern = level_dmma<N>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem) != cudaSuccess || occ < 1) return -1;
    const int64_t ntiles = spad / 8;
    if (sequential) {
        const int64_t tpb = seq_tiles;                       // tiles per group (8 warps x tiles each)
        const int64_t ngroups = (ntiles + tpb - 1) / tpb;
        const int64_t gx = std::min<int64_t>(ngroups, (int64_t)sm_count * occ);
        kern<<<dim3((unsigned)gx, 1), 256, smem, stream>>>(nodes, children, first, P, nnodes, ncat, tips, tips_ld, pool,
                                                            slot_stride, spad, pi, thr, shift_at_root, site_cat,
                                                            store_root, (int)tpb, count);
        return 0;
    }
    int64_t want_blocks = std::max<int64_t>(1, (int64_t)sm_count * occ / count);
    int64_t tpb = (ntiles + want_blocks - 1) / want_blocks;
    tpb = std::max<int64_t>(8, (tpb + 7) / 8 * 8);
    const int64_t gx = (ntiles + tpb - 1) / tpb;
    kern<<<dim3((unsigned)gx, count), 256, smem, stream>>>(nodes, children, first, P, nnodes, ncat, tips, tips_ld, pool,
                                                           slot_stride, spad, pi, thr, shift_at_root, site_cat,
                                                           store_root, (int)tpb, 0);
    return 0;
}

// Returns 0 on success, -2 when the node arity needs more shared memory than the SM has
// (the caller falls back to level_dense), -1 on a launch-configuration error.
// sequential: one persistent launch walks nodes [first, first+count) in order for each site group.
inline int launch_level_dmma(int n, const LevelNode* nodes, const LevelChild* children, int first, int count,
                             int max_arity, const double* P, int nnodes, int ncat, const uint8_t* tips,
                             int64_t tips_ld, double* pool, int64_t slot_stride, int64_t spad, const double* pi,
                             double thr, int shift_at_root, double* site_cat, int store_root, int sm_count,
                             size_t smem_optin, bool sequential, int seq_tiles, cudaStream_t stream)
{
    if (n == 20)
        return launch_level_dmma_n<20>(nodes, children, first, count, max_arity, P, nnodes, ncat, tips, tips_ld, pool,
                                       slot_stride, spad, pi, thr, shift_at_root, site_cat, store_root, sm_count,
                                       smem_optin, sequential, seq_tiles, stream);
    if (n == 61)
        return launch_level_dmma_n<61>(nodes, children, first, count, max_arity, P, nnodes, ncat, tips, tips_ld, pool,
                                       slot_stride, spad, pi, thr, shift_at_root, site_cat, store_root, sm_count,
                                       smem_optin, sequential, seq_tiles, stream);
    return -1;
}

}  // namespace pbk
