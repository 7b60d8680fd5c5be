// K2, fused site-resident form for dense alphabets (20 amino acids, 61 sense codons) on the FP64
// tensor cores: a CTA of 8 warps walks the WHOLE tree for 64 sites (8 per warp); conditional
// likelihoods never leave the SM.
//
// Restates Phylo_ctmc.pruning (lib/phylo_ctmc.ml:83-98) with SV.mul / SV.shift (:33-40, :76-78), on
// the same host-compiled accumulator/stack program as the 4-state fused kernel (pb_api.cu).
//   * ACC (the running SV of the subtree being folded) lives in registers in the DMMA C-fragment
//     layout: rows = states, columns = the warp's 8 sites.
//   * A mat-vec P.ACC is a GEMM [states x states] x [states x 8 sites]: ACC is written to a per-warp
//     shared-memory tile (XOR-swizzled 64-byte rows), read back as B fragments, and multiplied by
//     the branch's P, zero-padded to MP x (KP+4) and staged in shared memory for the whole CTA.
//     The matrices arrive by cp.async, double buffered: matrix m+1 streams in from L2 while the
//     tensor cores work on matrix m (one CTA barrier per matrix).
//   * A leaf child is a column look-up in the leaf branch's matrix, read from a transposed,
//     padded copy in global memory (a column is contiguous; the 1-2 MB of leaf matrices live in L2).
//   * The stack holds C fragments lane-privately in shared memory: no layout change on push/pop.
//   * Rescaling decisions and rescaled vectors are the reference's; the log-scale factors of a site
//     are accumulated as (mantissa, binary exponent) and logged once at the root, as in fused4.
// HBM traffic per site: ntaxa bytes in, 8 bytes out.  Bound: the FP64 tensor pipe.
#pragma once
#include "pb_kernels_dmma.cuh"
#include "pb_kernels_fused.cuh"

namespace pbk {

template <int N>
struct FdmmaShape {
    using S = DmmaShape<N>;
    static constexpr int MATSZ = S::MP * S::LD;          // doubles per staged (padded) inner matrix
    static constexpr int XSZ = S::KP * 8;                // doubles per warp conversion tile
    static constexpr int FRAG = S::MT * 4;               // doubles per lane per stack entry
    static constexpr int NBUF = (N <= 32) ? 4 : 2;       // ring of staged inner matrices
    static constexpr int CTAS = (N <= 32) ? 2 : 1;       // resident CTAs per SM the shared-memory budget is cut for
    static constexpr size_t fixed_bytes(int ntaxa, int nops, int nmat)
    {
        return ((size_t)S::MP + (size_t)NBUF * MATSZ + 8 * XSZ) * sizeof(double) + (size_t)ntaxa * 64 +
               (size_t)nops * sizeof(FusedOp2) + (size_t)nmat * sizeof(int) + 16;
    }
    static constexpr size_t level_bytes = (size_t)256 * FRAG * sizeof(double);
};

// Ppad[c][k] = P[c][inner_nodes[k]] zero-padded to MP x LD (row-major);
// PTleaf[c][k] = transpose of P[c][leaf_nodes[k]] zero-padded to N columns x MP rows.
template <int N>
__global__ void fdmma_prepare(const double* __restrict__ P, const int* __restrict__ inner_nodes, int ninner,
                              const int* __restrict__ leaf_nodes, int nleaf, int nnodes,
                              double* __restrict__ Ppad, double* __restrict__ PTleaf)
{
    using S = DmmaShape<N>;
    const int c = blockIdx.y, k = blockIdx.x;
    if (k < ninner) {
        const double* src = P + ((size_t)c * nnodes + inner_nodes[k]) * N * N;
        double* dst = Ppad + ((size_t)c * ninner + k) * S::MP * S::LD;
        for (int idx = threadIdx.x; idx < S::MP * S::LD; idx += blockDim.x) {
            const int i = idx / S::LD, j = idx - i * S::LD;
            dst[idx] = (i < N && j < N) ? src[i * N + j] : 0.0;
        }
    } else {
        const int l = k - ninner;
        const double* src = P + ((size_t)c * nnodes + leaf_nodes[l]) * N * N;
        double* dst = PTleaf + ((size_t)c * nleaf + l) * N * S::MP;
        for (int idx = threadIdx.x; idx < N * S::MP; idx += blockDim.x) {
            const int s = idx / S::MP, i = idx - s * S::MP;
            dst[idx] = (i < N) ? src[i * N + s] : 0.0;
        }
    }
}

// SV.shift on a C-fragment tile with the (mantissa, exponent) scale accumulators of the lane's two sites.
template <int N>
__device__ __forceinline__ void shift_frag_scaled(double (&v)[DmmaShape<N>::MT][4], Scale<false> (&sc)[2], int g, double thr)
{
    constexpr int MT = DmmaShape<N>::MT;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        double mn = INFINITY, mx = -INFINITY;
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            if (mt * 16 + g < N) { mn = fmin(mn, v[mt][j]); mx = fmax(mx, v[mt][j]); }
            if (mt * 16 + g + 8 < N) { mn = fmin(mn, v[mt][2 + j]); mx = fmax(mx, v[mt][2 + j]); }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (!(mn > thr)) {
            const double inv = 1.0 / mx;
#pragma unroll
            for (int mt = 0; mt < MT; mt++) { v[mt][j] = inv * v[mt][j]; v[mt][2 + j] = inv * v[mt][2 + j]; }
            sc[j].add(mx);
        }
    }
}

// The same SV.shift with fewer cross-lane steps: the `min v > thr` test is a per-lane predicate combined
// by one ballot per site, and the maxima are reduced only when some site of the warp rescales, with the
// two sites of a lane folded into one butterfly (even row groups carry site 0, odd ones site 1).
template <int N>
__device__ __forceinline__ void shift_frag_ballot(double (&v)[DmmaShape<N>::MT][4], Scale<false> (&sc)[2], int lane, double thr)
{
    constexpr int MT = DmmaShape<N>::MT;
    const int g = lane >> 2, tig = lane & 3;
    bool ok[2] = {true, true};
    double mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            if (mt * 16 + g < N) { ok[j] = ok[j] && (v[mt][j] > thr); mx[j] = fmax(mx[j], v[mt][j]); }
            if (mt * 16 + g + 8 < N) { ok[j] = ok[j] && (v[mt][2 + j] > thr); mx[j] = fmax(mx[j], v[mt][2 + j]); }
        }
    const unsigned b0 = __ballot_sync(0xffffffffu, ok[0]), b1 = __ballot_sync(0xffffffffu, ok[1]);
    if ((b0 & b1) == 0xffffffffu) return;                   // warp-uniform: nothing to rescale
    const bool odd = g & 1;
    double keep = odd ? mx[1] : mx[0];
    keep = fmax(keep, __shfl_xor_sync(0xffffffffu, odd ? mx[0] : mx[1], 4));
    keep = fmax(keep, __shfl_xor_sync(0xffffffffu, keep, 8));
    keep = fmax(keep, __shfl_xor_sync(0xffffffffu, keep, 16));
    const double other = __shfl_xor_sync(0xffffffffu, keep, 4);
    const double m[2] = {odd ? other : keep, odd ? keep : other};
    const unsigned grp = 0x11111111u << tig;                // the 8 row groups holding this lane's two sites
    const unsigned bits[2] = {b0, b1};
#pragma unroll
    for (int j = 0; j < 2; j++) {
        if ((bits[j] & grp) != grp) {
            const double inv = 1.0 / m[j];
#pragma unroll
            for (int mt = 0; mt < MT; mt++) { v[mt][j] = inv * v[mt][j]; v[mt][2 + j] = inv * v[mt][2 + j]; }
            sc[j].add(m[j]);
        }
    }
}

// prog: FusedOp2 with offA / offB = compact inner (APPLY, APPLY_MUL_TIP.A, ACC_POP) or leaf index,
// sh = tip row A | tip row B << 16.  mlist: compact inner indices in the order the program consumes them.
// The first stack_smem stack levels live in shared memory, deeper ones in gstack (per CTA, L2-resident).
template <int N>
__global__ void __launch_bounds__(256, FdmmaShape<N>::CTAS)
fused_dmma(const FusedOp2* __restrict__ prog, int nops, const int* __restrict__ mlist, int nmat, int stack_smem,
           double* __restrict__ gstack, int gdepth,
           const double* __restrict__ Ppad, const double* __restrict__ PTleaf, int ninner, int nleaf, int ncat,
           const uint8_t* __restrict__ tips, int64_t tips_ld, int ntaxa, int64_t spad,
           const double* __restrict__ pi, double thr, int shift_at_root, double* __restrict__ site_cat)
{
    using S = DmmaShape<N>;
    using F = FdmmaShape<N>;
    constexpr int MP = S::MP, LD = S::LD, MT = S::MT, KS = S::KS, NBUF = F::NBUF;
    extern __shared__ double shraw[];
    double* shPi = shraw;                                   // [MP]
    double* shM = shPi + MP;                                // [NBUF][MATSZ] ring of staged inner matrices
    double* shX = shM + NBUF * F::MATSZ;                    // [8 warps][XSZ] C -> B conversion tiles
    double* shS = shX + 8 * F::XSZ;                         // [stack_smem][FRAG][256 threads]
    uint8_t* shT = reinterpret_cast<uint8_t*>(shS + (size_t)stack_smem * 256 * F::FRAG);   // [ntaxa][64]
    FusedOp2* shOps = reinterpret_cast<FusedOp2*>(shT + (size_t)ntaxa * 64);
    int* shMl = reinterpret_cast<int*>(shOps + nops);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    if (tid < MP) shPi[tid] = (tid < N) ? pi[tid] : 0.0;
    {
        const int4* src = reinterpret_cast<const int4*>(prog);
        int4* dst = reinterpret_cast<int4*>(shOps);
        for (int idx = tid; idx < nops; idx += 256) dst[idx] = src[idx];
        for (int idx = tid; idx < nmat; idx += 256) shMl[idx] = mlist[idx];
    }
    double* myX = shX + warp * F::XSZ;
    gstack += (size_t)blockIdx.x * gdepth * 256 * F::FRAG;
    auto slot = [&](int level) -> double* {                 // element k of the entry is slot[k * 256]
        return (level < stack_smem ? shS + (size_t)level * 256 * F::FRAG
                                   : gstack + (size_t)(level - stack_smem) * 256 * F::FRAG) + tid;
    };
    __syncthreads();
    const int64_t ntiles = spad / 64;
    const int64_t work = ntiles * ncat;

    for (int64_t w = blockIdx.x; w < work; w += gridDim.x) {
        const int c = (int)(w / ntiles);
        const int64_t tile0 = (w - (int64_t)c * ntiles) * 64;             // the CTA's 64 sites
        const int64_t s0 = tile0 + warp * 8;                             // this warp's 8 sites
        const double* Pc = Ppad + (size_t)c * ninner * F::MATSZ;
        const double* PTc = PTleaf + (size_t)c * nleaf * N * MP;
        __syncthreads();                                     // the previous tile is done with shT and shM
        // tips of the CTA's 64 sites, every taxon: 4 x 16-byte chunks per row
        for (int idx = tid; idx < ntaxa * 4; idx += 256) {
            const int row = idx >> 2, q = idx & 3;
            cp_async16(shT + row * 64 + q * 16, tips + (int64_t)row * tips_ld + tile0 + q * 16);
        }
        cp_async_commit();
        auto load_matrix = [&](int m) {                      // inner matrix #m of the program -> ring slot m % NBUF
            if (m < nmat) {
                const double* src = Pc + (size_t)shMl[m] * F::MATSZ;
                double* dst = shM + (m % NBUF) * F::MATSZ;
                for (int idx = tid; idx < F::MATSZ / 2; idx += 256) cp_async16(dst + idx * 2, src + idx * 2);
            }
            cp_async_commit();                               // (an empty group keeps the group count uniform)
        };
#pragma unroll
        for (int m = 0; m < NBUF - 1; m++) load_matrix(m);
        int m_next = 0;                                      // next matrix the program will consume
        auto acquire = [&]() -> const double* {             // matrix m_next has landed; refill the slot of m_next - 1
            cp_async_wait<NBUF - 2>();
            __syncthreads();
            load_matrix(m_next + NBUF - 1);
            const double* Pm = shM + (m_next % NBUF) * F::MATSZ;
            m_next++;
            return Pm;
        };
        cp_async_wait<NBUF - 1>();
        __syncthreads();                                     // tips landed

        double acc[MT][4];
        Scale<false> sc[2];
        sc[0].init(); sc[1].init();
        int sp = 0;

        // y = P . x   (x in C-fragment registers -> swizzled smem tile -> B fragments -> DMMA)
        auto apply = [&](const double* Pm, const double (&x)[MT][4], double (&y)[MT][4]) {
            __syncwarp();
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const int r0 = mt * 16 + g, r1 = r0 + 8;     // rows >= N hold zeros (P is zero padded)
                if (r0 < S::KP) *reinterpret_cast<double2*>(myX + r0 * 8 + ((tig ^ (r0 & 3)) << 1)) = make_double2(x[mt][0], x[mt][1]);
                if (r1 < S::KP) *reinterpret_cast<double2*>(myX + r1 * 8 + ((tig ^ (r1 & 3)) << 1)) = make_double2(x[mt][2], x[mt][3]);
            }
            __syncwarp();
            double b[KS][2];
            const int off = (((g >> 1) ^ tig) << 1) + (g & 1);
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                b[ks][0] = myX[(ks * 8 + tig) * 8 + off];
                b[ks][1] = myX[(ks * 8 + tig + 4) * 8 + off];
            }
#pragma unroll
            for (int mt = 0; mt < MT; mt++) y[mt][0] = y[mt][1] = y[mt][2] = y[mt][3] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const double* r0 = Pm + (mt * 16 + g) * LD + tig + ks * 8;
                    const double* r1 = r0 + 8 * LD;
                    dmma_16x8x8(y[mt], r0[0], r1[0], r0[4], r1[4], b[ks][0], b[ks][1]);
                }
            }
        };
        // column look-up of a leaf branch: the lane's rows for its two sites
        auto leaf_cols = [&](int leaf, int tiprow, double (&t)[MT][4]) {
            const uint8_t* tb = shT + tiprow * 64 + warp * 8 + tig * 2;
            const double* ca = PTc + ((size_t)leaf * N + min((int)tb[0], N - 1)) * MP;
            const double* cb = PTc + ((size_t)leaf * N + min((int)tb[1], N - 1)) * MP;
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                t[mt][0] = ca[mt * 16 + g]; t[mt][1] = cb[mt * 16 + g];
                t[mt][2] = ca[mt * 16 + g + 8]; t[mt][3] = cb[mt * 16 + g + 8];
            }
        };

        for (int ip = 0; ip < nops; ip++) {
            const int4 opw = reinterpret_cast<const int4*>(shOps)[ip];
            const int code = opw.x & 0xff, rowA = opw.w & 0xffff, rowB = (opw.w >> 16) & 0xffff;
            if (code == FOP_APPLY_MUL_TIP) {
                double y[MT][4], t[MT][4];
                leaf_cols(opw.z, rowB, t);
                apply(acquire(), acc, y);
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[mt][e] = y[mt][e] * t[mt][e];
                shift_frag_scaled<N>(acc, sc, g, thr);
            } else if (code == FOP_ACC_POP) {
                const double* top = slot(--sp);
                double y[MT][4], x[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int e = 0; e < 4; e++) x[mt][e] = top[(mt * 4 + e) * 256];
                apply(acquire(), acc, y);
                apply(acquire(), x, acc);
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[mt][e] = y[mt][e] * acc[mt][e];
                shift_frag_scaled<N>(acc, sc, g, thr);
            } else if (code == FOP_TIP_TIP) {
                double t[MT][4];
                leaf_cols(opw.y, rowA, acc);
                leaf_cols(opw.z, rowB, t);
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[mt][e] = acc[mt][e] * t[mt][e];
                shift_frag_scaled<N>(acc, sc, g, thr);
            } else if (code == FOP_TIP) {
                leaf_cols(opw.y, rowA, acc);
            } else if (code == FOP_APPLY) {
                double y[MT][4];
                apply(acquire(), acc, y);
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[mt][e] = y[mt][e];
            } else if (code == FOP_MUL_TIP) {
                double t[MT][4];
                leaf_cols(opw.y, rowA, t);
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[mt][e] = acc[mt][e] * t[mt][e];
                shift_frag_scaled<N>(acc, sc, g, thr);
            } else if (code == FOP_POP_MUL) {
                const double* top = slot(--sp);
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[mt][e] = top[(mt * 4 + e) * 256] * acc[mt][e];
                shift_frag_scaled<N>(acc, sc, g, thr);
            } else if (code == FOP_ROOT) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const double p0 = shPi[mt * 16 + g], p1 = shPi[mt * 16 + g + 8];
                    acc[mt][0] *= p0; acc[mt][1] *= p0; acc[mt][2] *= p1; acc[mt][3] *= p1;
                }
                if (shift_at_root) shift_frag_scaled<N>(acc, sc, g, thr);
                double sum[2] = {0.0, 0.0};
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {            // padded rows hold exact zeros
                    sum[0] += acc[mt][0] + acc[mt][2];
                    sum[1] += acc[mt][1] + acc[mt][3];
                }
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    sum[0] += __shfl_xor_sync(0xffffffffu, sum[0], o);
                    sum[1] += __shfl_xor_sync(0xffffffffu, sum[1], o);
                }
                if (g == 0)
                    *reinterpret_cast<double2*>(site_cat + (int64_t)c * spad + s0 + tig * 2) =
                        make_double2(log(sum[0]) + sc[0].value(), log(sum[1]) + sc[1].value());
            }
            if (opw.x & FOP_PUSH_AFTER) {
                double* top = slot(sp++);
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int e = 0; e < 4; e++) top[(mt * 4 + e) * 256] = acc[mt][e];
            }
        }
        cp_async_wait<0>();                                  // drain the trailing (empty) groups
    }
}

// --------------------------------------------------------------------------
// fused_dmma2: the same program with the warps DECOUPLED.  In fused_dmma the CTA barrier per branch
// matrix keeps all warps in the same phase, so the tensor pipe idles whenever they rescale, convert
// layouts or gather tip columns together (ncu: 52 % DMMA pipe at 61 states, 34 % at 20).  Here
//   * a producer warp streams the inner branch matrices, cut into k-panels, through a ring in
//     shared memory with 1-D bulk copies (cp.async.bulk, one elected lane) completing on "full"
//     mbarriers; every compute warp releases a panel with one arrive on its "empty" mbarrier;
//   * compute warps never meet at a CTA barrier: each runs the program for its own 8 sites at its
//     own pace (up to a ring's worth apart), so one warp's rescaling overlaps another's DMMAs;
//   * a panel holds, per parent-state row, the 8 k-columns of one k-step reordered as
//     (k, k+4) pairs: one 16-byte load per row gives both A-fragment halves, conflict free;
//   * stack entries are stored in the B-fragment tile layout, so the popped operand of a
//     two-inner-children node feeds the tensor cores without passing through registers.
// Sites are split over CTAs in 8-site units (balanced to one unit), every CTA walks all categories.
// --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// Release of a ring slot whose contents were read with ld.shared: the arrive must not be issued before
// the loads have RETURNED — an mbarrier arrive does not queue behind earlier shared-memory loads, and
// with 15 warps bursting 16-byte fragment loads the load queue can be deeper than the refill latency
// of the slot (seen on B200 as rare stale A fragments).  `witness` is the AND of the high words of the
// loaded values: the arrive is predicated on it not being all ones (never true for transition
// probabilities), which makes the instruction wait on the loads' scoreboards.
__device__ __forceinline__ void mbar_arrive_after(uint64_t* b, uint32_t witness)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0xffffffff;\n @p mbarrier.arrive.shared::cta.b64 _, [%0];\n}"
                 ::"r"(smem_u32(b)), "r"(witness) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

template <int N>
struct Fd2Shape {
    using S = DmmaShape<N>;
    static constexpr int NW = 15;                               // compute warps per CTA (+1 producer warp = 512 threads, 128 registers each; 23 warps x 80 registers measured slower)
    static constexpr int KPP = (N <= 32) ? S::KS : 1;           // k-steps per ring panel
    static constexpr int NPAN = S::KS / KPP;                    // panels per matrix
    static constexpr int PSZ = S::MP * 8 * KPP;                 // doubles per panel
    static constexpr int MATSZ = S::MP * 8 * S::KS;             // doubles per panelled matrix
    static constexpr int XSZ = S::KP * 8;                       // doubles per 8-site tile [KP states][8 sites]
    static constexpr int NSLOT_MAX = (N <= 32) ? 8 : 14;        // ring slots: as many as fit once the stack is on chip,
    static constexpr int NSLOT_MIN = (N <= 32) ? 3 : 6;         // between these bounds (16 mbarrier pairs are reserved)
    static constexpr size_t BAR_BYTES = 256;
    static constexpr size_t fixed_bytes(int nslot, int ntaxa, int nops, int nmat)
    {
        return BAR_BYTES + ((size_t)nslot * PSZ + S::MP) * sizeof(double) + (((size_t)NW * ntaxa * 8 + 15) & ~(size_t)15) +
               (size_t)nops * sizeof(FusedOp2) + (size_t)nmat * sizeof(int) + 16;
    }
    static constexpr size_t level_bytes = (size_t)NW * XSZ * sizeof(double);   // one stack level (or the scratch tile) of all warps
};

// Apanel[c][k][ks][row][tig][h] = P[c][inner_nodes[k]][row][ks*8 + tig + 4h], zero padded.
template <int N>
__global__ void fdmma2_prepare(const double* __restrict__ P, const int* __restrict__ inner_nodes, int ninner,
                               const int* __restrict__ leaf_nodes, int nleaf, int nnodes,
                               double* __restrict__ Apanel, double* __restrict__ PTleaf)
{
    using S = DmmaShape<N>;
    using F = Fd2Shape<N>;
    const int c = blockIdx.y, k = blockIdx.x;
    if (k < ninner) {
        const double* src = P + ((size_t)c * nnodes + inner_nodes[k]) * N * N;
        double* dst = Apanel + ((size_t)c * ninner + k) * F::MATSZ;
        for (int idx = threadIdx.x; idx < F::MATSZ; idx += blockDim.x) {
            const int h = idx & 1, tig = (idx >> 1) & 3, row = (idx >> 3) % S::MP, ks = idx / (8 * S::MP);
            const int col = ks * 8 + tig + 4 * h;
            dst[idx] = (row < N && col < N) ? src[row * N + col] : 0.0;
        }
    } else {
        const int l = k - ninner;
        const double* src = P + ((size_t)c * nnodes + leaf_nodes[l]) * N * N;
        double* dst = PTleaf + ((size_t)c * nleaf + l) * N * S::MP;
        for (int idx = threadIdx.x; idx < N * S::MP; idx += blockDim.x) {
            const int s = idx / S::MP, i = idx - s * S::MP;
            dst[idx] = (i < N) ? src[i * N + s] : 0.0;
        }
    }
}

template <int N>
__global__ void __launch_bounds__((Fd2Shape<N>::NW + 1) * 32, 1)
fused_dmma2(const FusedOp2* __restrict__ prog, int nops, const int* __restrict__ mlist, int nmat, int nslot, int stack_smem,
            double* __restrict__ gstack, int gdepth,
            const double* __restrict__ Apanel, const double* __restrict__ PTleaf, int ninner, int nleaf, int ncat,
            const uint8_t* __restrict__ tips, int64_t tips_ld, int ntaxa, int64_t spad,
            const double* __restrict__ pi, double thr, int shift_at_root, int stagger, double* __restrict__ site_cat)
{
    using S = DmmaShape<N>;
    using F = Fd2Shape<N>;
    constexpr int MP = S::MP, MT = S::MT, NW = F::NW, KPP = F::KPP, NPAN = F::NPAN;
    extern __shared__ __align__(128) unsigned char shbytes[];
    uint64_t* full = reinterpret_cast<uint64_t*>(shbytes);              // [nslot] (16 reserved)
    uint64_t* empty = full + 16;                                          // [nslot]
    double* ring = reinterpret_cast<double*>(shbytes + F::BAR_BYTES);   // [nslot][PSZ]
    double* shPi = ring + nslot * F::PSZ;                                 // [MP]
    double* shS = shPi + MP;                                              // [stack_smem][NW][XSZ]: stack levels; the level above the top is scratch
    uint8_t* shT = reinterpret_cast<uint8_t*>(shS + (size_t)stack_smem * NW * F::XSZ);   // [NW][ntaxa][8]
    FusedOp2* shOps = reinterpret_cast<FusedOp2*>(shT + (((size_t)NW * ntaxa * 8 + 15) & ~(size_t)15));
    int* shMl = reinterpret_cast<int*>(shOps + nops);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    if (tid == 0) {
        for (int i = 0; i < nslot; i++) { mbar_init(full + i, 1); mbar_init(empty + i, NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < MP) shPi[tid] = (tid < N) ? pi[tid] : 0.0;
    {
        const int4* src = reinterpret_cast<const int4*>(prog);
        int4* dst = reinterpret_cast<int4*>(shOps);
        for (int idx = tid; idx < nops; idx += blockDim.x) dst[idx] = src[idx];
        for (int idx = tid; idx < nmat; idx += blockDim.x) shMl[idx] = mlist[idx];
    }
    __syncthreads();                                         // the only CTA-wide barrier

    // this CTA's 8-site units [lo, hi), NW per round
    const int64_t units = spad / 8;
    const int64_t lo = units * blockIdx.x / gridDim.x, hi = units * (blockIdx.x + 1) / gridDim.x;
    const int rounds = (int)((hi - lo + NW - 1) / NW);
    int slot = 0;
    uint32_t phase = 0;
    auto advance = [&]() { if (++slot == nslot) { slot = 0; phase ^= 1; } };

    if (warp == NW) {                                        // ---- producer: one lane feeds the ring
        if (lane == 0) {
            for (int c = 0; c < ncat; c++)
                for (int r = 0; r < rounds; r++)
                    for (int m = 0; m < nmat; m++) {
                        const double* src = Apanel + ((size_t)c * ninner + shMl[m]) * F::MATSZ;
                        for (int p = 0; p < NPAN; p++) {
                            mbar_wait(empty + slot, phase ^ 1);
                            mbar_expect_tx(full + slot, F::PSZ * sizeof(double));
                            bulk_g2s(ring + slot * F::PSZ, src + p * F::PSZ, F::PSZ * sizeof(double), full + slot);
                            advance();
                        }
                    }
        }
        return;
    }

    // ---- compute warps
    uint8_t* myT = shT + (size_t)warp * ntaxa * 8;
    gstack += ((size_t)blockIdx.x * gdepth * NW + warp) * F::XSZ;
    auto level_tile = [&](int level) -> double* {
        return level < stack_smem ? shS + ((size_t)level * NW + warp) * F::XSZ
                                  : gstack + (size_t)(level - stack_smem) * NW * F::XSZ;
    };
    const int boff = (((g >> 1) ^ tig) << 1) + (g & 1);     // B-fragment column of this lane in a swizzled row

    // registers (C-fragment layout) -> 8-site tile [state][site], 16-byte chunk q of row r at q ^ (r & 3)
    auto to_tile = [&](double* T, const double (&x)[MT][4]) {
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const int r0 = mt * 16 + g, r1 = r0 + 8;         // rows >= N hold zeros
            if (r0 < S::KP) *reinterpret_cast<double2*>(T + r0 * 8 + ((tig ^ (r0 & 3)) << 1)) = make_double2(x[mt][0], x[mt][1]);
            if (r1 < S::KP) *reinterpret_cast<double2*>(T + r1 * 8 + ((tig ^ (r1 & 3)) << 1)) = make_double2(x[mt][2], x[mt][3]);
        }
        __syncwarp();
    };
    // y = P . x, x read as B fragments from a tile, P from the next NPAN panels of the ring
    auto apply = [&](const double* T, double (&y)[MT][4]) {
#pragma unroll
        for (int mt = 0; mt < MT; mt++) y[mt][0] = y[mt][1] = y[mt][2] = y[mt][3] = 0.0;
#pragma unroll 1
        for (int p = 0; p < NPAN; p++) {
            double b[KPP][2];
#pragma unroll
            for (int kk = 0; kk < KPP; kk++) {
                const int ks = p * KPP + kk;
                b[kk][0] = T[(ks * 8 + tig) * 8 + boff];
                b[kk][1] = T[(ks * 8 + tig + 4) * 8 + boff];
            }
            mbar_wait(full + slot, phase);
            const double* A = ring + slot * F::PSZ + g * 8 + tig * 2;
            uint32_t witness = 0xffffffffu;
#pragma unroll
            for (int kk = 0; kk < KPP; kk++) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const double2 lo2 = *reinterpret_cast<const double2*>(A + (kk * MP + mt * 16) * 8);
                    const double2 hi2 = *reinterpret_cast<const double2*>(A + (kk * MP + mt * 16 + 8) * 8);
                    witness &= (uint32_t)__double2hiint(lo2.x) & (uint32_t)__double2hiint(hi2.x);
                    dmma_16x8x8(y[mt], lo2.x, hi2.x, lo2.y, hi2.y, b[kk][0], b[kk][1]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_after(empty + slot, witness);
            advance();
        }
    };
    auto leaf_ptrs = [&](const double* PTc, int leaf, int tiprow, const double*& ca, const double*& cb) {
        const uint8_t* tb = myT + tiprow * 8 + tig * 2;
        ca = PTc + ((size_t)leaf * N + min((int)tb[0], N - 1)) * MP + g;
        cb = PTc + ((size_t)leaf * N + min((int)tb[1], N - 1)) * MP + g;
    };

    if (stagger > 0 && (warp & 4)) __nanosleep(stagger);     // start half of the warps out of phase

    for (int c = 0; c < ncat; c++) {
        const double* PTc = PTleaf + (size_t)c * nleaf * N * MP;
        for (int r = 0; r < rounds; r++) {
            const int64_t unit = lo + (int64_t)r * NW + warp;
            if (unit >= hi) {                                // no sites left for this warp: keep the ring moving
                for (int p = 0; p < nmat * NPAN; p++) {
                    mbar_wait(full + slot, phase);
                    if (lane == 0) mbar_arrive(empty + slot);
                    advance();
                }
                continue;
            }
            const int64_t s0 = unit * 8;
            __syncwarp();
            for (int row = lane; row < ntaxa; row += 32)
                *reinterpret_cast<uint2*>(myT + row * 8) = *reinterpret_cast<const uint2*>(tips + (int64_t)row * tips_ld + s0);
            __syncwarp();

            double acc[MT][4];
            Scale<false> sc[2];
            sc[0].init(); sc[1].init();
            int sp = 0;
#pragma unroll 1
            for (int ip = 0; ip < nops; ip++) {
                const int4 opw = reinterpret_cast<const int4*>(shOps)[ip];
                const int code = opw.x & 0xff, rowA = opw.w & 0xffff, rowB = (opw.w >> 16) & 0xffff;
                if (code == FOP_APPLY_MUL_TIP) {
                    double y[MT][4];
                    const double *ca, *cb;
                    leaf_ptrs(PTc, opw.z, rowB, ca, cb);
                    double* X = level_tile(sp);              // scratch: the free level above the stack top
                    to_tile(X, acc);
                    apply(X, y);
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        acc[mt][0] = y[mt][0] * ca[mt * 16]; acc[mt][1] = y[mt][1] * cb[mt * 16];
                        acc[mt][2] = y[mt][2] * ca[mt * 16 + 8]; acc[mt][3] = y[mt][3] * cb[mt * 16 + 8];
                    }
                    shift_frag_ballot<N>(acc, sc, lane, thr);
                } else if (code == FOP_ACC_POP) {
                    double y[MT][4];
                    double* X = level_tile(sp);
                    to_tile(X, acc);
                    apply(X, y);
                    apply(level_tile(--sp), acc);
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int e = 0; e < 4; e++) acc[mt][e] = y[mt][e] * acc[mt][e];
                    shift_frag_ballot<N>(acc, sc, lane, thr);
                } else if (code == FOP_TIP_TIP) {
                    const double *ca, *cb, *da, *db;
                    leaf_ptrs(PTc, opw.y, rowA, ca, cb);
                    leaf_ptrs(PTc, opw.z, rowB, da, db);
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        acc[mt][0] = ca[mt * 16] * da[mt * 16]; acc[mt][1] = cb[mt * 16] * db[mt * 16];
                        acc[mt][2] = ca[mt * 16 + 8] * da[mt * 16 + 8]; acc[mt][3] = cb[mt * 16 + 8] * db[mt * 16 + 8];
                    }
                    shift_frag_ballot<N>(acc, sc, lane, thr);
                } else if (code == FOP_TIP) {
                    const double *ca, *cb;
                    leaf_ptrs(PTc, opw.y, rowA, ca, cb);
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        acc[mt][0] = ca[mt * 16]; acc[mt][1] = cb[mt * 16];
                        acc[mt][2] = ca[mt * 16 + 8]; acc[mt][3] = cb[mt * 16 + 8];
                    }
                } else if (code == FOP_APPLY) {
                    double* X = level_tile(sp);
                    to_tile(X, acc);
                    apply(X, acc);
                } else if (code == FOP_MUL_TIP) {
                    const double *ca, *cb;
                    leaf_ptrs(PTc, opw.y, rowA, ca, cb);
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        acc[mt][0] *= ca[mt * 16]; acc[mt][1] *= cb[mt * 16];
                        acc[mt][2] *= ca[mt * 16 + 8]; acc[mt][3] *= cb[mt * 16 + 8];
                    }
                    shift_frag_ballot<N>(acc, sc, lane, thr);
                } else if (code == FOP_POP_MUL) {
                    const double* T = level_tile(--sp);
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        const int r0 = mt * 16 + g, r1 = r0 + 8;
                        if (r0 < S::KP) {
                            const double2 v = *reinterpret_cast<const double2*>(T + r0 * 8 + ((tig ^ (r0 & 3)) << 1));
                            acc[mt][0] *= v.x; acc[mt][1] *= v.y;
                        }
                        if (r1 < S::KP) {
                            const double2 v = *reinterpret_cast<const double2*>(T + r1 * 8 + ((tig ^ (r1 & 3)) << 1));
                            acc[mt][2] *= v.x; acc[mt][3] *= v.y;
                        }
                    }
                    shift_frag_ballot<N>(acc, sc, lane, thr);
                } else if (code == FOP_ROOT) {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        const double p0 = shPi[mt * 16 + g], p1 = shPi[mt * 16 + g + 8];
                        acc[mt][0] *= p0; acc[mt][1] *= p0; acc[mt][2] *= p1; acc[mt][3] *= p1;
                    }
                    if (shift_at_root) shift_frag_ballot<N>(acc, sc, lane, thr);
                    double sum[2] = {0.0, 0.0};
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {        // padded rows hold exact zeros
                        sum[0] += acc[mt][0] + acc[mt][2];
                        sum[1] += acc[mt][1] + acc[mt][3];
                    }
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        sum[0] += __shfl_xor_sync(0xffffffffu, sum[0], o);
                        sum[1] += __shfl_xor_sync(0xffffffffu, sum[1], o);
                    }
                    if (g == 0)
                        *reinterpret_cast<double2*>(site_cat + (int64_t)c * spad + s0 + tig * 2) =
                            make_double2(log(sum[0]) + sc[0].value(), log(sum[1]) + sc[1].value());
                }
                if (opw.x & FOP_PUSH_AFTER) to_tile(level_tile(sp++), acc);
            }
        }
    }
}

// Shared-memory plan: how many stack levels stay on chip (the rest spill to an L2-resident scratch).
// Returns false when even the fixed part (matrix ring, conversion tiles, the tile's tips, the program)
// does not fit — the caller then takes the level-scheduled path.
template <int N>
bool fused_dmma_plan_n(size_t smem_optin, int stack_depth, int ntaxa, int nops, int nmat, int* stack_smem, size_t* smem)
{
    using F = FdmmaShape<N>;
    const size_t budget = F::CTAS == 1 ? smem_optin : (smem_optin + 1024) / F::CTAS - 1024;
    const size_t fixed = F::fixed_bytes(ntaxa, nops, nmat);
    if (fixed > budget) return false;
    const int fit = (int)std::min<size_t>((budget - fixed) / F::level_bytes, (size_t)stack_depth);
    *stack_smem = fit;
    *smem = fixed + (size_t)fit * F::level_bytes;
    return true;
}

inline bool fused_dmma_plan(int n, size_t smem_optin, int stack_depth, int ntaxa, int nops, int nmat, int* stack_smem, size_t* smem)
{
    if (n == 20) return fused_dmma_plan_n<20>(smem_optin, stack_depth, ntaxa, nops, nmat, stack_smem, smem);
    if (n == 61) return fused_dmma_plan_n<61>(smem_optin, stack_depth, ntaxa, nops, nmat, stack_smem, smem);
    return false;
}

inline size_t fused_dmma_matsz(int n) { return n == 20 ? FdmmaShape<20>::MATSZ : FdmmaShape<61>::MATSZ; }
inline size_t fused_dmma_leafsz(int n) { return n == 20 ? 20 * DmmaShape<20>::MP : 61 * DmmaShape<61>::MP; }
inline size_t fused_dmma_frag(int n) { return n == 20 ? FdmmaShape<20>::FRAG : FdmmaShape<61>::FRAG; }

template <int N>
int fused_dmma_grid_n(int sm_count, size_t smem, int64_t work)
{
    auto kern = fused_dmma<N>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem) != cudaSuccess || occ < 1) return -1;
    return (int)std::max<int64_t>(1, std::min<int64_t>(work, (int64_t)sm_count * occ));
}

inline int fused_dmma_grid(int n, int sm_count, size_t smem, int64_t work)
{
    return n == 20 ? fused_dmma_grid_n<20>(sm_count, smem, work) : fused_dmma_grid_n<61>(sm_count, smem, work);
}

struct FusedDmmaArgs {
    const FusedOp2* prog; int nops; const int* mlist; int nmat; int stack_smem; double* gstack; int gdepth;
    const double* Ppad; const double* PTleaf; int ninner, nleaf, ncat;
    const uint8_t* tips; int64_t tips_ld; int ntaxa; int64_t spad;
    const double* pi; double thr; int shift_at_root; double* site_cat;
    // fused_dmma3 only: tabulated cherries (reduced program; see DenseTab)
    const double* tabV = nullptr; const int* tabK = nullptr; const int4* tabs = nullptr; int ntab = 0;
    int64_t tab_entries = 0;          // entries of all tables of one category (N^2 per cherry, N^3 per three-leaf clade)
    int64_t cat_stride = 0;           // fused_dmma3: stride of site_cat between categories (0: spad); a launch over a column
                                      // chunk passes tips / site_cat at its first column and spad = its width
};

template <int N>
void launch_fused_dmma_n(const FusedDmmaArgs& a, int grid, size_t smem, cudaStream_t stream)
{
    fused_dmma<N><<<grid, 256, smem, stream>>>(a.prog, a.nops, a.mlist, a.nmat, a.stack_smem, a.gstack, a.gdepth, a.Ppad,
                                               a.PTleaf, a.ninner, a.nleaf, a.ncat, a.tips, a.tips_ld, a.ntaxa, a.spad,
                                               a.pi, a.thr, a.shift_at_root, a.site_cat);
}

inline void launch_fused_dmma(int n, const FusedDmmaArgs& a, int grid, size_t smem, cudaStream_t stream)
{
    if (n == 20) launch_fused_dmma_n<20>(a, grid, smem, stream);
    else launch_fused_dmma_n<61>(a, grid, smem, stream);
}

inline void launch_fdmma_prepare(int n, const double* P, const int* inner_nodes, int ninner, const int* leaf_nodes, int nleaf,
                                 int nnodes, int ncat, double* Ppad, double* PTleaf, cudaStream_t stream)
{
    dim3 grid(ninner + nleaf, ncat);
    if (n == 20) fdmma_prepare<20><<<grid, 256, 0, stream>>>(P, inner_nodes, ninner, leaf_nodes, nleaf, nnodes, Ppad, PTleaf);
    else fdmma_prepare<61><<<grid, 256, 0, stream>>>(P, inner_nodes, ninner, leaf_nodes, nleaf, nnodes, Ppad, PTleaf);
}

// --------------------------------------------------------------------------
// fused_dmma3: fused_dmma2 re-tiled on the native DMMA shape, mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; the
// m16n8k8 form is four of these).  Parent states are padded to a multiple of 8 and child states to a
// multiple of 4 instead of 16 and 8: a 20-state mat-vec takes 3 x 5 = 15 DMMAs instead of 24 (the FP64
// tensor and vector instructions share one pipe on B200 — profiles/r01i_fp64_dfma_dmma_share_pipe.json —
// so padding is paid in full), fragments are 2 doubles per 8-row tile, and every per-row loop (leaf
// columns, rescaling, layout conversion) shrinks with the padding.  61 states: 8 x 16 = 128, as before.
// Ring panels hold, per 8-row tile and PAIR of k-steps, one 16-byte (k, k+4) entry per lane:
// panel[kk][mt][lane] = { P[mt*8+g][(2kk)*4+tig], P[mt*8+g][(2kk+1)*4+tig] }.
// --------------------------------------------------------------------------
__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

#ifndef FD3_NW20
#define FD3_NW20 11   // compute warps per CTA at 20 states (see Fd3Shape::NW)
#endif
#ifndef FD3_CTAS20
#define FD3_CTAS20 2  // CTAs per SM at 20 states (see Fd3Shape::CTAS)
#endif
template <int N>
struct Fd3Shape {
    static constexpr int NW = (N <= 32) ? FD3_NW20 : 15;              // compute warps per CTA (+1 producer warp): 768 threads x 80 registers at 20 states (measured: 15 warps 1.02 ms, 23 0.915, 31 0.908 at c3 in round 1; with cherry tables 23 warps 0.575 ms, 27 warps x 72 registers 0.576: -DFD3_NW20=27, profiles/r03a_sweep_dense_nw27_fold.jsonl), 512 x 128 at 61
    static constexpr int CTAS = (N <= 32) ? FD3_CTAS20 : 1;     // CTAs per SM: two independent CTAs of 11 + 1 warps at 20 states (their rounds interleave)
    static constexpr int M8 = (N + 7) / 8;                      // 8-row tiles of parent states (20: 3, 61: 8)
    static constexpr int K4 = (N + 3) / 4;                      // k-steps of 4 child states (20: 5, 61: 16)
    static constexpr int KP2 = (K4 + 1) / 2;                    // pairs of k-steps (20: 3, 61: 8)
    static constexpr int KR = K4 * 4;                           // rows of an 8-site tile (20, 64)
    static constexpr int MP8 = M8 * 8;                          // padded parent states (24, 64)
    static constexpr int KPP = (N <= 32) ? KP2 : 1;             // k-step pairs per ring panel
    static constexpr int NPAN = KP2 / KPP;                      // panels per matrix
    static constexpr int PSZ = KPP * M8 * 64;                   // doubles per panel (20: 576, 61: 512)
    static constexpr int MATSZ = KP2 * M8 * 64;
    static constexpr int XSZ = KR * 8;                          // doubles per 8-site tile
    static constexpr int NSLOT_MAX = (N <= 32) ? 8 : 14;
    static constexpr int NSLOT_MIN = (N <= 32) ? 3 : 6;
    static constexpr size_t BAR_BYTES = 256;
    static constexpr size_t fixed_bytes(int nslot, int ntaxa, int nops, int nmat)
    {
        return BAR_BYTES + ((size_t)nslot * PSZ + MP8) * sizeof(double) + (((size_t)NW * ntaxa * 8 + 15) & ~(size_t)15) +
               (size_t)nops * sizeof(FusedOp2) + (size_t)nmat * sizeof(int) + 16;
    }
    static constexpr size_t level_bytes = (size_t)NW * XSZ * sizeof(double);
};

template <int N>
__global__ void fdmma3_prepare(const double* __restrict__ P, const int* __restrict__ inner_nodes, int ninner,
                               const int* __restrict__ leaf_nodes, int nleaf, int nnodes,
                               double* __restrict__ Apanel, double* __restrict__ PTleaf)
{
    using F = Fd3Shape<N>;
    const int c = blockIdx.y, k = blockIdx.x;
    if (k < ninner) {
        const double* src = P + ((size_t)c * nnodes + inner_nodes[k]) * N * N;
        double* dst = Apanel + ((size_t)c * ninner + k) * F::MATSZ;
        for (int idx = threadIdx.x; idx < F::MATSZ; idx += blockDim.x) {
            const int h = idx & 1, lane = (idx >> 1) & 31, mt = (idx >> 6) % F::M8, pr = idx / (64 * F::M8);
            const int row = mt * 8 + (lane >> 2), col = (2 * pr + h) * 4 + (lane & 3);
            dst[idx] = (row < N && col < N) ? src[row * N + col] : 0.0;
        }
    } else {
        const int l = k - ninner;
        const double* src = P + ((size_t)c * nnodes + leaf_nodes[l]) * N * N;
        double* dst = PTleaf + ((size_t)c * nleaf + l) * N * F::MP8;
        for (int idx = threadIdx.x; idx < N * F::MP8; idx += blockDim.x) {
            const int s = idx / F::MP8, i = idx - s * F::MP8;
            dst[idx] = (i < N) ? src[i * N + s] : 0.0;
        }
    }
}

// SV.shift on 8-row C tiles: v[mt][j] = state mt*8+g of the lane's site j (same scheme as shift_frag_ballot).
template <int N>
__device__ __forceinline__ void shift_tiles8(double (&v)[Fd3Shape<N>::M8][2], Scale<false> (&sc)[2], int lane, double thr)
{
    constexpr int M8 = Fd3Shape<N>::M8;
    const int g = lane >> 2, tig = lane & 3;
    bool ok[2] = {true, true};
    double mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int mt = 0; mt < M8; mt++)
            if (mt * 8 + 7 < N || mt * 8 + g < N) { ok[j] = ok[j] && (v[mt][j] > thr); mx[j] = fmax(mx[j], v[mt][j]); }
    const unsigned b0 = __ballot_sync(0xffffffffu, ok[0]), b1 = __ballot_sync(0xffffffffu, ok[1]);
    if ((b0 & b1) == 0xffffffffu) return;
    const bool odd = g & 1;
    double keep = odd ? mx[1] : mx[0];
    keep = fmax(keep, __shfl_xor_sync(0xffffffffu, odd ? mx[0] : mx[1], 4));
    keep = fmax(keep, __shfl_xor_sync(0xffffffffu, keep, 8));
    keep = fmax(keep, __shfl_xor_sync(0xffffffffu, keep, 16));
    const double other = __shfl_xor_sync(0xffffffffu, keep, 4);
    const double m[2] = {odd ? other : keep, odd ? keep : other};
    const unsigned grp = 0x11111111u << tig;
    const unsigned bits[2] = {b0, b1};
#pragma unroll
    for (int j = 0; j < 2; j++) {
        if ((bits[j] & grp) != grp) {
            const double inv = 1.0 / m[j];
#pragma unroll
            for (int mt = 0; mt < M8; mt++) v[mt][j] = inv * v[mt][j];
            sc[j].add(m[j]);
        }
    }
}

// Power-of-two form (scaling mode 2 of pb_kernels_fused.cuh): the test and the exponent of the largest
// component are taken on the high words, so the cross-lane steps are one ballot per site and 32-bit
// shuffles, and there is no division; v is multiplied by an exact power of two.
template <int N>
__device__ __forceinline__ void shift_tiles8(double (&v)[Fd3Shape<N>::M8][2], Scale<2> (&sc)[2], int lane, double thr)
{
    constexpr int M8 = Fd3Shape<N>::M8;
    const int g = lane >> 2, tig = lane & 3;
    const int thr_hi = __double2hiint(thr);
    int hmin[2] = {0x7fffffff, 0x7fffffff}, hmax[2] = {(int)0x80000000, (int)0x80000000};
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int mt = 0; mt < M8; mt++)
            if (mt * 8 + 7 < N || mt * 8 + g < N) {
                const int h = __double2hiint(v[mt][j]);
                hmin[j] = min(hmin[j], h);
                hmax[j] = max(hmax[j], h);
            }
    const unsigned b0 = __ballot_sync(0xffffffffu, hmin[0] > thr_hi), b1 = __ballot_sync(0xffffffffu, hmin[1] > thr_hi);
    if ((b0 & b1) == 0xffffffffu) return;                   // warp-uniform: nothing to rescale
    const bool odd = g & 1;
    int keep = odd ? hmax[1] : hmax[0];
    keep = max(keep, __shfl_xor_sync(0xffffffffu, odd ? hmax[0] : hmax[1], 4));
    keep = max(keep, __shfl_xor_sync(0xffffffffu, keep, 8));
    keep = max(keep, __shfl_xor_sync(0xffffffffu, keep, 16));
    const int other = __shfl_xor_sync(0xffffffffu, keep, 4);
    const int m[2] = {odd ? other : keep, odd ? keep : other};
    const unsigned grp = 0x11111111u << tig;
    const unsigned bits[2] = {b0, b1};
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int d = ((bits[j] & grp) == grp) ? 1023 : (m[j] >> 20);      // biased exponent to take out (1023: none)
        const double f = __hiloint2double((2046 - d) << 20, 0);
#pragma unroll
        for (int mt = 0; mt < M8; mt++) v[mt][j] *= f;
        sc[j].e += d - 1023;
    }
}

// Tabulated cherries (dense alphabets).  A node whose two children are leaves can take only N^2 values per rate
// category, whatever the site (400 at 20 states, 3 721 at 61): once per evaluation every such node gets a table of
//     P[v] . shift(col_a(P[leaf A]) * col_b(P[leaf B]))        and the shift's exponent,
// P[v] the matrix of the branch ABOVE the cherry, so that the kernel replaces two column gathers, a product, a
// rescaling AND its parent's matrix-vector product (15 / 128 DMMAs, a layout conversion and a trip through the matrix
// ring) by one gather of the same shape as a leaf-column gather: a third of the matrix-vector products of a
// birth-death tree.  The program is compiled over the reduced tree (pb_api.cu: build_dense_reduced_program) with
// the table as an operand of its parent's instruction, exactly as in the 4-state kernel fused4t.
// A table descriptor: x = first entry, y / z = tip-matrix rows of the two leaves (entry index = state y * N + state z).
// What a table is made of: leaves A, B under node `cherry`; optionally (three-leaf clade) a third leaf C joining the
// cherry under `top`.  The table holds P[top] . clv(top) (without the matrix when top is the root of the tree).
struct DenseTabNodes { int leafA, leafB, cherry, leafC, top, is_root, pad0, pad1; };

template <int N>
__global__ void __launch_bounds__(256)
fdmma3_build_cherries(const double* __restrict__ P, int nnodes, const int4* __restrict__ tabs, const DenseTabNodes* __restrict__ tab_nodes,
                      int64_t tab_entries, double thr, double* __restrict__ tabV, int* __restrict__ tabK)
{
    using F = Fd3Shape<N>;
    // A block builds a slice of ONE table, four entries per step (one per group of 64 threads).  Thread l of a group
    // owns output state l and keeps ROW l of the matrix applied last in registers for the whole slice; the entry's
    // vector is broadcast from shared memory, so a step costs N broadcast loads and N FMAs per thread.  (Two earlier
    // forms were bound elsewhere: rows of P read from global memory, 64 lines per step - 193 us at c4; the matrix
    // transposed in shared memory, two loads per FMA - 131 us; profiles/r02j_*, r02n_*.)
    // Three-leaf clades (20 states): the inner step shift(col_a * col_b) -> P[cherry] . -> * col_c -> shift runs from
    // shared memory (P[cherry] transposed), then the same final product.
    extern __shared__ double sh[];
    double* PAs = sh;                                    // [N][N]  leaf A: P[i][a]
    double* PBs = PAs + N * N;                           // [N][N]  leaf B
    double* svs = PBs + N * N;                           // [4][64] the entry's vector, per group
    int* reds = reinterpret_cast<int*>(svs + 4 * 64);    // [4][2 warps][min, max]
    double* PCs = svs + 4 * 64 + 8;                      // [N][N]  leaf C          (three-leaf clades only)
    double* PMt = PCs + N * N;                           // [N][65] P[cherry] transposed
    const int t = blockIdx.x, c = blockIdx.y, tid = threadIdx.x, grp = tid >> 6, l = tid & 63;
    const DenseTabNodes nd = tab_nodes[t];
    const bool three = nd.leafC >= 0;
    const int nent = three ? N * N * N : N * N;
    const double* Pc = P + (size_t)c * nnodes * N * N;
    const bool pre = !nd.is_root;
    double pv[N];
#pragma unroll
    for (int j = 0; j < N; j++) pv[j] = (pre && l < N) ? Pc[(size_t)nd.top * N * N + l * N + j] : 0.0;
    for (int idx = tid; idx < N * N; idx += 256) {
        PAs[idx] = Pc[(size_t)nd.leafA * N * N + idx];
        PBs[idx] = Pc[(size_t)nd.leafB * N * N + idx];
        if (three) {
            PCs[idx] = Pc[(size_t)nd.leafC * N * N + idx];
            PMt[(idx % N) * 65 + idx / N] = Pc[(size_t)nd.cherry * N * N + idx];
        }
    }
    const int thr_hi = __double2hiint(thr);
    const size_t base = (size_t)c * tab_entries + (size_t)tabs[t].x;
    int* fold_flag = tabK + (size_t)gridDim.y * tab_entries;
    const bool fold = *reinterpret_cast<volatile int*>(fold_flag) == 0;    // (a mix of folded and unfolded entries is valid)
    double* sv = svs + grp * 64;
    int* red = reds + grp * 4;
    auto shift_group = [&](double v, int& d_out) {       // SV.shift over the group's N values, power-of-two form; barrier inside
        const int h = __double2hiint(v);
        const int wmin = __reduce_min_sync(0xffffffffu, l < N ? h : 0x7fffffff);
        const int wmax = __reduce_max_sync(0xffffffffu, l < N ? h : (int)0x80000000);
        if ((l & 31) == 0) { red[(l >> 5) * 2] = wmin; red[(l >> 5) * 2 + 1] = wmax; }
        sv[l] = v;
        __syncthreads();
        const int hmin = min(red[0], red[2]), hmax = max(red[1], red[3]);
        d_out = (hmin > thr_hi) ? 1023 : (hmax >> 20);
        return __hiloint2double((2046 - d_out) << 20, 0);
    };
    __syncthreads();
    for (int e0 = blockIdx.z * 4; e0 < nent; e0 += gridDim.z * 4) {
        const int e = e0 + grp;
        const bool valid = e < nent;
        const int ev = valid ? e : 0;
        const int cc = three ? ev % N : 0, ab = three ? ev / N : ev;
        const int a = ab / N, b = ab - a * N;
        double v = (l < N) ? PAs[l * N + a] * PBs[l * N + b] : 0.0;      // column a of leaf A times column b of leaf B
        int d = 1023, d1 = 1023;
        double f = shift_group(v, d);
        if (three) {                                     // (uniform over the block: a block builds one table)
            double w = 0.0;
            if (l < N) {
#pragma unroll 4
                for (int j = 0; j < N; j++) w = fma(PMt[j * 65 + l], sv[j] * f, w);
                w *= PCs[l * N + cc];
            }
            d1 = d;
            __syncthreads();                             // everyone has read sv
            v = w;
            f = shift_group(v, d);
        }
        double y = v;
        if (pre) {                                       // P[top] . vector: four partial sums (the order is not the kernel's anyway)
            double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
            for (int j = 0; j + 3 < N; j += 4) {
                const double2 s01 = *reinterpret_cast<const double2*>(sv + j), s23 = *reinterpret_cast<const double2*>(sv + j + 2);
                y0 = fma(pv[j], s01.x, y0);
                y1 = fma(pv[j + 1], s01.y, y1);
                y2 = fma(pv[j + 2], s23.x, y2);
                y3 = fma(pv[j + 3], s23.y, y3);
            }
#pragma unroll
            for (int j = N / 4 * 4; j < N; j++) y0 = fma(pv[j], sv[j], y0);
            y = (y0 + y1) + (y2 + y3);
        }
        // (scaling by a power of two commutes with every rounding above: f * (P . v) = P . (f * v) exactly)
        // The entry's exponent k is folded into its vector when it is small (see fold_exponent, pb_kernels_fused.cuh):
        // y f 2^k = y 2^(d1 - 1023), and the kernel then skips the exponent gathers.  tabK[ncat * tab_entries] is the
        // flag "some exponent was not folded" (preset to non-zero when folding is switched off).
        int k = (d - 1023) + (d1 - 1023);
        if (fold && k >= -480) {
            f = __hiloint2double(d1 << 20, 0);
            k = 0;
        } else if (fold && valid && l == 0)
            atomicOr(fold_flag, 1);
        if (valid && l < F::MP8) tabV[(base + e) * F::MP8 + l] = (l < N) ? y * f : 0.0;
        if (valid && l == 0) tabK[base + e] = k;
        __syncthreads();
    }
}

template <int N, int SCALING>
__global__ void __launch_bounds__((Fd3Shape<N>::NW + 1) * 32, Fd3Shape<N>::CTAS)
fused_dmma3(const FusedOp2* __restrict__ prog, int nops, const int* __restrict__ mlist, int nmat, int nslot, int stack_smem,
            double* __restrict__ gstack, int gdepth,
            const double* __restrict__ Apanel, const double* __restrict__ PTleaf, int ninner, int nleaf, int ncat,
            const uint8_t* __restrict__ tips, int64_t tips_ld, int ntaxa, int64_t spad,
            const double* __restrict__ pi, double thr, int shift_at_root, double* __restrict__ site_cat,
            const double* __restrict__ tabV, const int* __restrict__ tabK, const int4* __restrict__ tabs, int ntab,
            int64_t tab_entries, int64_t cat_stride)
{
    // (spad: the site columns of THIS launch — tips and site_cat point at its first column; tips_ld and cat_stride are the
    // strides of the whole arrays: a pipelined host evaluation launches the kernel per column chunk)
    using F = Fd3Shape<N>;
    constexpr int M8 = F::M8, K4 = F::K4, MP8 = F::MP8, NW = F::NW, KPP = F::KPP, NPAN = F::NPAN;
    extern __shared__ __align__(128) unsigned char shbytes[];
    uint64_t* full = reinterpret_cast<uint64_t*>(shbytes);              // [nslot] (16 reserved)
    uint64_t* empty = full + 16;
    double* ring = reinterpret_cast<double*>(shbytes + F::BAR_BYTES);   // [nslot][PSZ]
    double* shPi = ring + nslot * F::PSZ;                                 // [MP8]
    double* shS = shPi + MP8;                                             // [stack_smem][NW][XSZ]
    uint8_t* shT = reinterpret_cast<uint8_t*>(shS + (size_t)stack_smem * NW * F::XSZ);   // [NW][ntaxa][8]
    FusedOp2* shOps = reinterpret_cast<FusedOp2*>(shT + (((size_t)NW * ntaxa * 8 + 15) & ~(size_t)15));
    int4* shTabs = reinterpret_cast<int4*>(shOps + nops);               // [ntab] cherry descriptors (16 bytes each, like an op)
    int* shMl = reinterpret_cast<int*>(shTabs + ntab);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    if (tid == 0) {
        for (int i = 0; i < nslot; i++) { mbar_init(full + i, 1); mbar_init(empty + i, NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < MP8) shPi[tid] = (tid < N) ? pi[tid] : 0.0;
    {
        const int4* src = reinterpret_cast<const int4*>(prog);
        int4* dst = reinterpret_cast<int4*>(shOps);
        for (int idx = tid; idx < nops; idx += blockDim.x) dst[idx] = src[idx];
        for (int idx = tid; idx < ntab; idx += blockDim.x) shTabs[idx] = tabs[idx];
        for (int idx = tid; idx < nmat; idx += blockDim.x) shMl[idx] = mlist[idx];
    }
    __syncthreads();                                         // the only CTA-wide barrier

    const int64_t units = spad / 8;
    const int64_t lo = units * blockIdx.x / gridDim.x, hi = units * (blockIdx.x + 1) / gridDim.x;
    const int rounds = (int)((hi - lo + NW - 1) / NW);
    int slot = 0;
    uint32_t phase = 0;
    auto advance = [&]() { if (++slot == nslot) { slot = 0; phase ^= 1; } };

    if (warp == NW) {                                        // ---- producer: one lane feeds the ring
        if (lane == 0) {
            for (int c = 0; c < ncat; c++)
                for (int r = 0; r < rounds; r++)
                    for (int m = 0; m < nmat; m++) {
                        const double* src = Apanel + ((size_t)c * ninner + shMl[m]) * F::MATSZ;
                        for (int p = 0; p < NPAN; p++) {
                            mbar_wait(empty + slot, phase ^ 1);
                            mbar_expect_tx(full + slot, F::PSZ * sizeof(double));
                            bulk_g2s(ring + slot * F::PSZ, src + p * F::PSZ, F::PSZ * sizeof(double), full + slot);
                            advance();
                        }
                    }
        }
        return;
    }

    // ---- compute warps
    uint8_t* myT = shT + (size_t)warp * ntaxa * 8;
    gstack += ((size_t)blockIdx.x * gdepth * NW + warp) * F::XSZ;
    auto level_tile = [&](int level) -> double* {
        return level < stack_smem ? shS + ((size_t)level * NW + warp) * F::XSZ
                                  : gstack + (size_t)(level - stack_smem) * NW * F::XSZ;
    };
    const int boff = (((g >> 1) ^ tig) << 1) + (g & 1);     // B-fragment column of this lane in a swizzled row

    // registers (C tiles) -> 8-site tile [state][site], 16-byte chunk q of row r at q ^ (r & 3)
    auto to_tile = [&](double* T, const double (&x)[M8][2]) {
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < M8; mt++) {
            const int r = mt * 8 + g;                        // rows >= N hold zeros
            if (mt * 8 + 7 < F::KR || r < F::KR)
                *reinterpret_cast<double2*>(T + r * 8 + ((tig ^ (r & 3)) << 1)) = make_double2(x[mt][0], x[mt][1]);
        }
        __syncwarp();
    };
    // y = P . x, x read as B fragments from a tile, P from the next NPAN panels of the ring
    auto apply = [&](const double* T, double (&y)[M8][2]) {
#pragma unroll
        for (int mt = 0; mt < M8; mt++) y[mt][0] = y[mt][1] = 0.0;
#pragma unroll 1
        for (int p = 0; p < NPAN; p++) {
            double b[KPP][2];
#pragma unroll
            for (int kk = 0; kk < KPP; kk++) {
                const int ks = 2 * (p * KPP + kk);
                b[kk][0] = T[(ks * 4 + tig) * 8 + boff];
                b[kk][1] = (ks + 1 < K4) ? T[((ks + 1) * 4 + tig) * 8 + boff] : 0.0;
            }
            mbar_wait(full + slot, phase);
            const double* A = ring + slot * F::PSZ + lane * 2;
            uint32_t witness = 0xffffffffu;
#pragma unroll
            for (int kk = 0; kk < KPP; kk++) {
#pragma unroll
                for (int mt = 0; mt < M8; mt++) {
                    const double2 a = *reinterpret_cast<const double2*>(A + (kk * M8 + mt) * 64);
                    witness &= (uint32_t)__double2hiint(a.x);
                    dmma_8x8x4(y[mt], a.x, b[kk][0]);
                    if (2 * (p * KPP + kk) + 1 < K4) dmma_8x8x4(y[mt], a.y, b[kk][1]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_after(empty + slot, witness);
            advance();
        }
    };
    auto leaf_ptrs = [&](const double* PTc, int leaf, int tiprow, const double*& ca, const double*& cb) {
        const uint8_t* tb = myT + tiprow * 8 + tig * 2;
        ca = PTc + ((size_t)leaf * N + min((int)tb[0], N - 1)) * MP8 + g;
        cb = PTc + ((size_t)leaf * N + min((int)tb[1], N - 1)) * MP8 + g;
    };

    // a tabulated cherry's value for the lane's two sites: a gather of the same shape as a leaf column, plus the exponents
    const double* Tc = nullptr;
    const int* Kc = nullptr;
    auto tab_ptrs = [&](int t, const double*& ca, const double*& cb, Scale<SCALING> (&sc)[2]) {
        const int4 d = shTabs[t];                          // first entry, tip rows A, B and (clades of three leaves) C or -1
        const uint8_t* t0 = myT + d.y * 8 + tig * 2;
        const uint8_t* t1 = myT + d.z * 8 + tig * 2;
        int ia = min((int)t0[0], N - 1) * N + min((int)t1[0], N - 1);
        int ib = min((int)t0[1], N - 1) * N + min((int)t1[1], N - 1);
        if (d.w >= 0) {
            const uint8_t* t2 = myT + d.w * 8 + tig * 2;
            ia = ia * N + min((int)t2[0], N - 1);
            ib = ib * N + min((int)t2[1], N - 1);
        }
        ia += d.x;
        ib += d.x;
        ca = Tc + (size_t)ia * MP8 + g;
        cb = Tc + (size_t)ib * MP8 + g;
        if (Kc) {                                          // (null: every exponent is folded into its vector)
            sc[0].e += Kc[ia];
            sc[1].e += Kc[ib];
        }
    };

    for (int c = 0; c < ncat; c++) {
        const double* PTc = PTleaf + (size_t)c * nleaf * N * MP8;
        Tc = tabV + (size_t)c * tab_entries * MP8;
        Kc = (tabK && tabK[(size_t)ncat * tab_entries] != 0) ? tabK + (size_t)c * tab_entries : nullptr;
        for (int r = 0; r < rounds; r++) {
            const int64_t unit = lo + (int64_t)r * NW + warp;
            if (unit >= hi) {                                // no sites left for this warp: keep the ring moving
                for (int p = 0; p < nmat * NPAN; p++) {
                    mbar_wait(full + slot, phase);
                    if (lane == 0) mbar_arrive(empty + slot);
                    advance();
                }
                continue;
            }
            const int64_t s0 = unit * 8;
            __syncwarp();
            for (int row = lane; row < ntaxa; row += 32)
                *reinterpret_cast<uint2*>(myT + row * 8) = *reinterpret_cast<const uint2*>(tips + (int64_t)row * tips_ld + s0);
            __syncwarp();

            double acc[M8][2];
            Scale<SCALING> sc[2];
            sc[0].init(); sc[1].init();
            int sp = 0;
#pragma unroll 1
            for (int ip = 0; ip < nops; ip++) {
                const int4 opw = reinterpret_cast<const int4*>(shOps)[ip];
                const int code = opw.x & 0xff, rowA = opw.w & 0xffff, rowB = (opw.w >> 16) & 0xffff;
                if (code == FOP_APPLY_MUL_TIP) {
                    double y[M8][2];
                    const double *ca, *cb;
                    leaf_ptrs(PTc, opw.z, rowB, ca, cb);
                    double* X = level_tile(sp);              // scratch: the free level above the stack top
                    to_tile(X, acc);
                    apply(X, y);
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) { acc[mt][0] = y[mt][0] * ca[mt * 8]; acc[mt][1] = y[mt][1] * cb[mt * 8]; }
                    shift_tiles8<N>(acc, sc, lane, thr);
                } else if (code == FOP_ACC_POP) {
                    double y[M8][2];
                    double* X = level_tile(sp);
                    to_tile(X, acc);
                    apply(X, y);
                    apply(level_tile(--sp), acc);
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) { acc[mt][0] *= y[mt][0]; acc[mt][1] *= y[mt][1]; }
                    shift_tiles8<N>(acc, sc, lane, thr);
                } else if (SCALING == 2 && code == FOP_APPLY_MUL_TAB) {     // (tables exist in the power-of-two mode only)
                    double y[M8][2];
                    const double *ca, *cb;
                    tab_ptrs(opw.z, ca, cb, sc);
                    double* X = level_tile(sp);
                    to_tile(X, acc);
                    apply(X, y);
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) { acc[mt][0] = y[mt][0] * ca[mt * 8]; acc[mt][1] = y[mt][1] * cb[mt * 8]; }
                    shift_tiles8<N>(acc, sc, lane, thr);
                } else if (SCALING == 2 && code == FOP_TAB_TAB) {
                    const double *ca, *cb, *da, *db;
                    tab_ptrs(opw.y, ca, cb, sc);
                    tab_ptrs(opw.z, da, db, sc);
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) { acc[mt][0] = ca[mt * 8] * da[mt * 8]; acc[mt][1] = cb[mt * 8] * db[mt * 8]; }
                    shift_tiles8<N>(acc, sc, lane, thr);
                } else if (SCALING == 2 && code == FOP_TAB_TIP) {
                    const double *ca, *cb, *da, *db;
                    tab_ptrs(opw.y, ca, cb, sc);
                    leaf_ptrs(PTc, opw.z, rowB, da, db);
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) { acc[mt][0] = ca[mt * 8] * da[mt * 8]; acc[mt][1] = cb[mt * 8] * db[mt * 8]; }
                    shift_tiles8<N>(acc, sc, lane, thr);
                } else if (SCALING == 2 && code == FOP_TAB) {
                    const double *ca, *cb;
                    tab_ptrs(opw.y, ca, cb, sc);
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) { acc[mt][0] = ca[mt * 8]; acc[mt][1] = cb[mt * 8]; }
                } else if (SCALING == 2 && code == FOP_MUL_TAB) {
                    const double *ca, *cb;
                    tab_ptrs(opw.y, ca, cb, sc);
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) { acc[mt][0] *= ca[mt * 8]; acc[mt][1] *= cb[mt * 8]; }
                    shift_tiles8<N>(acc, sc, lane, thr);
                } else if (code == FOP_TIP_TIP) {
                    const double *ca, *cb, *da, *db;
                    leaf_ptrs(PTc, opw.y, rowA, ca, cb);
                    leaf_ptrs(PTc, opw.z, rowB, da, db);
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) { acc[mt][0] = ca[mt * 8] * da[mt * 8]; acc[mt][1] = cb[mt * 8] * db[mt * 8]; }
                    shift_tiles8<N>(acc, sc, lane, thr);
                } else if (code == FOP_TIP) {
                    const double *ca, *cb;
                    leaf_ptrs(PTc, opw.y, rowA, ca, cb);
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) { acc[mt][0] = ca[mt * 8]; acc[mt][1] = cb[mt * 8]; }
                } else if (code == FOP_APPLY) {
                    double* X = level_tile(sp);
                    to_tile(X, acc);
                    apply(X, acc);
                } else if (code == FOP_MUL_TIP) {
                    const double *ca, *cb;
                    leaf_ptrs(PTc, opw.y, rowA, ca, cb);
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) { acc[mt][0] *= ca[mt * 8]; acc[mt][1] *= cb[mt * 8]; }
                    shift_tiles8<N>(acc, sc, lane, thr);
                } else if (code == FOP_POP_MUL) {
                    const double* T = level_tile(--sp);
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) {
                        const int r = mt * 8 + g;
                        if (mt * 8 + 7 < F::KR || r < F::KR) {
                            const double2 v = *reinterpret_cast<const double2*>(T + r * 8 + ((tig ^ (r & 3)) << 1));
                            acc[mt][0] *= v.x; acc[mt][1] *= v.y;
                        }
                    }
                    shift_tiles8<N>(acc, sc, lane, thr);
                } else if (code == FOP_ROOT) {
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) {
                        const double p0 = shPi[mt * 8 + g];
                        acc[mt][0] *= p0; acc[mt][1] *= p0;
                    }
                    if (shift_at_root) shift_tiles8<N>(acc, sc, lane, thr);
                    double sum[2] = {0.0, 0.0};
#pragma unroll
                    for (int mt = 0; mt < M8; mt++) { sum[0] += acc[mt][0]; sum[1] += acc[mt][1]; }   // padded rows are exact zeros
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        sum[0] += __shfl_xor_sync(0xffffffffu, sum[0], o);
                        sum[1] += __shfl_xor_sync(0xffffffffu, sum[1], o);
                    }
                    if (g == 0)
                        *reinterpret_cast<double2*>(site_cat + (int64_t)c * cat_stride + s0 + tig * 2) =
                            make_double2(log(sum[0]) + sc[0].value(), log(sum[1]) + sc[1].value());
                }
                if (opw.x & FOP_PUSH_AFTER) to_tile(level_tile(sp++), acc);
            }
        }
    }
}

// ---- fused_dmma2 host side
// Shared-memory plan: stack_depth + 1 tiles per warp are needed (the level above the top is the layout-
// conversion scratch).  The ring gives up slots (down to NSLOT_MIN) before a level leaves the chip;
// levels that still do not fit live in an L2-resident scratch (correct, slower).
template <int N>
bool fused_dmma2_plan_n(size_t smem_optin, int stack_depth, int ntaxa, int nops, int nmat, int* nslot, int* stack_smem, size_t* smem)
{
    using F = Fd2Shape<N>;
    const int levels = stack_depth + 1;
    if (F::fixed_bytes(F::NSLOT_MIN, ntaxa, nops, nmat) > smem_optin) return false;
    int ns = F::NSLOT_MAX;
    while (ns > F::NSLOT_MIN && F::fixed_bytes(ns, ntaxa, nops, nmat) + (size_t)levels * F::level_bytes > smem_optin) ns--;
    const size_t fixed = F::fixed_bytes(ns, ntaxa, nops, nmat);
    const int fit = (int)std::min<size_t>((smem_optin - fixed) / F::level_bytes, (size_t)levels);
    *nslot = ns;
    *stack_smem = fit;
    *smem = fixed + (size_t)fit * F::level_bytes;
    return true;
}

inline bool fused_dmma2_plan(int n, size_t smem_optin, int stack_depth, int ntaxa, int nops, int nmat, int* nslot, int* stack_smem, size_t* smem)
{
    if (n == 20) return fused_dmma2_plan_n<20>(smem_optin, stack_depth, ntaxa, nops, nmat, nslot, stack_smem, smem);
    if (n == 61) return fused_dmma2_plan_n<61>(smem_optin, stack_depth, ntaxa, nops, nmat, nslot, stack_smem, smem);
    return false;
}

inline size_t fused_dmma2_panelsz(int n) { return n == 20 ? Fd2Shape<20>::PSZ : Fd2Shape<61>::PSZ; }
inline size_t fused_dmma2_matsz(int n) { return n == 20 ? Fd2Shape<20>::MATSZ : Fd2Shape<61>::MATSZ; }
inline size_t fused_dmma2_tilesz(int n) { return n == 20 ? Fd2Shape<20>::XSZ * Fd2Shape<20>::NW : Fd2Shape<61>::XSZ * Fd2Shape<61>::NW; }

template <int N>
int launch_fused_dmma2_n(const FusedDmmaArgs& a, int nslot, int stagger, int sm_count, size_t smem, cudaStream_t stream)
{
    using F = Fd2Shape<N>;
    auto kern = fused_dmma2<N>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    const int64_t units = a.spad / 8;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((units + F::NW - 1) / F::NW, sm_count));
    kern<<<grid, (F::NW + 1) * 32, smem, stream>>>(a.prog, a.nops, a.mlist, a.nmat, nslot, a.stack_smem, a.gstack, a.gdepth, a.Ppad,
                                                   a.PTleaf, a.ninner, a.nleaf, a.ncat, a.tips, a.tips_ld, a.ntaxa, a.spad,
                                                   a.pi, a.thr, a.shift_at_root, stagger, a.site_cat);
    return grid;
}

inline int launch_fused_dmma2(int n, const FusedDmmaArgs& a, int nslot, int stagger, int sm_count, size_t smem, cudaStream_t stream)
{
    return n == 20 ? launch_fused_dmma2_n<20>(a, nslot, stagger, sm_count, smem, stream)
                   : launch_fused_dmma2_n<61>(a, nslot, stagger, sm_count, smem, stream);
}

inline int fused_dmma2_grid(int n, int sm_count, int64_t spad)
{
    const int nw = n == 20 ? Fd2Shape<20>::NW : Fd2Shape<61>::NW;
    return (int)std::max<int64_t>(1, std::min<int64_t>((spad / 8 + nw - 1) / nw, sm_count));
}

inline void launch_fdmma2_prepare(int n, const double* P, const int* inner_nodes, int ninner, const int* leaf_nodes, int nleaf,
                                  int nnodes, int ncat, double* Apanel, double* PTleaf, cudaStream_t stream)
{
    dim3 grid(ninner + nleaf, ncat);
    if (n == 20) fdmma2_prepare<20><<<grid, 256, 0, stream>>>(P, inner_nodes, ninner, leaf_nodes, nleaf, nnodes, Apanel, PTleaf);
    else fdmma2_prepare<61><<<grid, 256, 0, stream>>>(P, inner_nodes, ninner, leaf_nodes, nleaf, nnodes, Apanel, PTleaf);
}

// ---- fused_dmma3 host side (same plan rules as fused_dmma2)
template <int N>
bool fused_dmma3_plan_n(size_t smem_optin, int stack_depth, int ntaxa, int nops, int nmat, int* nslot, int* stack_smem, size_t* smem)
{
    using F = Fd3Shape<N>;
    const int levels = stack_depth + 1;
    smem_optin = smem_optin / F::CTAS - (F::CTAS > 1 ? 1024 : 0);        // (per CTA; 1 KB per resident CTA is the system's)
    if (F::fixed_bytes(F::NSLOT_MIN, ntaxa, nops, nmat) > smem_optin) return false;
    int ns = F::NSLOT_MAX;
    while (ns > F::NSLOT_MIN && F::fixed_bytes(ns, ntaxa, nops, nmat) + (size_t)levels * F::level_bytes > smem_optin) ns--;
    const size_t fixed = F::fixed_bytes(ns, ntaxa, nops, nmat);
    const int fit = (int)std::min<size_t>((smem_optin - fixed) / F::level_bytes, (size_t)levels);
    *nslot = ns;
    *stack_smem = fit;
    *smem = fixed + (size_t)fit * F::level_bytes;
    return true;
}

inline bool fused_dmma3_plan(int n, size_t smem_optin, int stack_depth, int ntaxa, int nops, int nmat, int* nslot, int* stack_smem, size_t* smem)
{
    if (n == 20) return fused_dmma3_plan_n<20>(smem_optin, stack_depth, ntaxa, nops, nmat, nslot, stack_smem, smem);
    if (n == 61) return fused_dmma3_plan_n<61>(smem_optin, stack_depth, ntaxa, nops, nmat, nslot, stack_smem, smem);
    return false;
}

inline size_t fused_dmma3_panelsz(int n) { return n == 20 ? Fd3Shape<20>::PSZ : Fd3Shape<61>::PSZ; }
inline size_t fused_dmma3_matsz(int n) { return n == 20 ? Fd3Shape<20>::MATSZ : Fd3Shape<61>::MATSZ; }
inline size_t fused_dmma3_leafsz(int n) { return n == 20 ? 20 * Fd3Shape<20>::MP8 : 61 * Fd3Shape<61>::MP8; }
inline size_t fused_dmma3_tilesz(int n) { return n == 20 ? Fd3Shape<20>::XSZ * Fd3Shape<20>::NW : Fd3Shape<61>::XSZ * Fd3Shape<61>::NW; }
inline int fused_dmma3_grid(int n, int sm_count, int64_t spad)
{
    const int nw = n == 20 ? Fd3Shape<20>::NW : Fd3Shape<61>::NW;
    const int ctas = n == 20 ? Fd3Shape<20>::CTAS : Fd3Shape<61>::CTAS;
    return (int)std::max<int64_t>(1, std::min<int64_t>((spad / 8 + nw - 1) / nw, (int64_t)sm_count * ctas));
}

// Site columns one ROUND of the kernel covers when every CTA is resident (each compute warp one 8-site unit).
inline int64_t fused_dmma3_round_sites(int n, int sm_count)
{
    const int nw = n == 20 ? Fd3Shape<20>::NW : Fd3Shape<61>::NW;
    const int ctas = n == 20 ? Fd3Shape<20>::CTAS : Fd3Shape<61>::CTAS;
    return (int64_t)sm_count * ctas * nw * 8;
}

template <int N, int SCALING>
int launch_fused_dmma3_n(const FusedDmmaArgs& a, int nslot, int sm_count, size_t smem, cudaStream_t stream)
{
    using F = Fd3Shape<N>;
    auto kern = fused_dmma3<N, SCALING>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    const int grid = fused_dmma3_grid(N, sm_count, a.spad);
    kern<<<grid, (F::NW + 1) * 32, smem, stream>>>(a.prog, a.nops, a.mlist, a.nmat, nslot, a.stack_smem, a.gstack, a.gdepth, a.Ppad,
                                                   a.PTleaf, a.ninner, a.nleaf, a.ncat, a.tips, a.tips_ld, a.ntaxa, a.spad,
                                                   a.pi, a.thr, a.shift_at_root, a.site_cat, a.tabV, a.tabK, a.tabs, a.ntab, a.tab_entries,
                                                   a.cat_stride ? a.cat_stride : a.spad);
    return grid;
}

inline void launch_fdmma3_build_cherries(int n, const double* P, int nnodes, const int4* tabs, const DenseTabNodes* tab_nodes, int ntab,
                                         int64_t tab_entries, bool any_three, int ncat, double thr, double* tabV, int* tabK,
                                         cudaStream_t stream, int sm_count = 148, int waves = 0)
{
    if (ntab < 1) return;
    // A block per (table, category, slice), 4 entries per step.  The slices fill WHOLE waves of the resident blocks and
    // never spill into another one: at 61 states a block holds a matrix row per thread (164 registers, one block per
    // SM), and 300 blocks on 148 SMs ran as three waves of which the last held four blocks (profiles/r02o_*).
    const int per_sm = n == 20 ? 3 : 1;
    if (waves < 1) waves = 1;                                   // (c4: one wave 0.737 ms, two 0.744, three 0.750: profiles/r03l_*)
    const int slices = std::max(1, std::min(64, waves * per_sm * sm_count / (ntab * ncat)));
    const dim3 grid(ntab, ncat, slices);
    const size_t smem = ((any_three ? 3 : 2) * (size_t)n * n + (any_three ? (size_t)n * 65 : 0) + 4 * 64 + 8) * sizeof(double);
    if (n == 20) {
        fdmma3_build_cherries<20><<<grid, 256, smem, stream>>>(P, nnodes, tabs, tab_nodes, tab_entries, thr, tabV, tabK);
    } else {
        cudaFuncSetAttribute(fdmma3_build_cherries<61>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // 62 KB
        fdmma3_build_cherries<61><<<grid, 256, smem, stream>>>(P, nnodes, tabs, tab_nodes, tab_entries, thr, tabV, tabK);
    }
}

inline int launch_fused_dmma3(int n, bool pow2, const FusedDmmaArgs& a, int nslot, int sm_count, size_t smem, cudaStream_t stream)
{
    if (pow2)
        return n == 20 ? launch_fused_dmma3_n<20, 2>(a, nslot, sm_count, smem, stream) : launch_fused_dmma3_n<61, 2>(a, nslot, sm_count, smem, stream);
    return n == 20 ? launch_fused_dmma3_n<20, 0>(a, nslot, sm_count, smem, stream) : launch_fused_dmma3_n<61, 0>(a, nslot, sm_count, smem, stream);
}

inline void launch_fdmma3_prepare(int n, const double* P, const int* inner_nodes, int ninner, const int* leaf_nodes, int nleaf,
                                  int nnodes, int ncat, double* Apanel, double* PTleaf, cudaStream_t stream)
{
    dim3 grid(ninner + nleaf, ncat);
    if (n == 20) fdmma3_prepare<20><<<grid, 256, 0, stream>>>(P, inner_nodes, ninner, leaf_nodes, nleaf, nnodes, Apanel, PTleaf);
    else fdmma3_prepare<61><<<grid, 256, 0, stream>>>(P, inner_nodes, ninner, leaf_nodes, nleaf, nnodes, Apanel, PTleaf);
}

}  // namespace pbk
