// K2, fused site-resident form (4 states): the whole post-order traversal of one
// (site, category) runs in one thread; conditional likelihoods never leave the SM.
//
// Restates Phylo_ctmc.pruning (lib/phylo_ctmc.ml:83-98) with SV.mul / SV.shift
// (:33-40, :76-78).  The tree is compiled on the host (pb_api.cu) into a short
// accumulator/stack program that every thread executes in lock step:
//   * binary nodes evaluate their deeper subtree first (x*y and c1+c2 commute, so the
//     result is bit-identical) which bounds the stack by log2(ntaxa);
//   * nodes of arity >= 3 keep the reference's left-fold order.
// Tip states are re-packed once per (alignment, tree) into 2-bit codes IN PROGRAM ORDER,
// 32 tips per 64-bit word, word-major / site-contiguous: a thread streams its site's tips
// with one coalesced 8-byte load per 32 tips, prefetched two words ahead.
//
// Rescaling: the rule (min v > threshold ? keep : v := (1/max v) v) and therefore every
// vector are the reference's; the log-scale accumulator sums log(max v) over the rescaling
// events of the whole traversal.  With EXACT_LOG it adds log(mx) at every event like the
// reference's carry; otherwise it accumulates the product of the mx as (mantissa, binary
// exponent) and takes one log at the root: the same real number, rounded differently
// (<= ~1e-15 relative on the log-likelihood, inside the 1e-9 per-site tolerance).
//
// HBM traffic per site: ntaxa/4 bytes in, 8*C bytes out.  Bound: the FP64 pipe.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pb_kernels.cuh"

namespace pbk {

struct FusedOp2 {          // 16 bytes: one uniform 128-bit shared-memory load per instruction
    int code;              // FusedOpCode, FOP_LOADW or FOP_ACC_POP
    int offA, offB;        // offsets (in doubles) of P[pA], P[pB] in the staged matrices: node id * 16
    int sh;                // shA | shB << 8: bit offsets of tip A / B in the current packed word
};
constexpr int FOP_LOADW = 8;     // advance the packed-tip window by one 64-bit word
constexpr int FOP_ACC_POP = 9;   // ACC = shiftmul(P[pA].ACC, P[pB].pop())   (both children inner)
constexpr int FOP_CHERRY = 10;   // ACC = table[offA + ((window >> shA) & offB)]: the 4^k possible values of a clade of k leaves
                                 // (16 for a node with two leaf children, ...; see build_clade_tables)

// Tip states uint8 [ntaxa][ld] -> 2-bit codes in program order: packed[w][site], bits
// 2j..2j+1 of word w = state of the (32 w + j)-th tip the program consumes.
// Also flags any state >= 4.
// (The 32 loads of a word are issued before any of them is used, from row indices staged in shared memory: a
// first form that loaded order[k] and then the byte inside a guarded loop ran its 2 x ntaxa loads per thread back
// to back, 70 us per 250k sites.)
constexpr int PACK_ORDER_SMEM = 8192;     // tips whose row indices are staged in shared memory (else read from global)

// `xorder` (nullable): slot k holds state(order[k]) XOR state(xorder[k]) when xorder[k] >= 0 — the reduced program
// stores every tip of a tabulated clade but the first as its DIFFERENCE from the first, so that the clade's most
// frequent patterns (all tips equal) index four ADJACENT table entries, one 128-byte line (see fused4t).
__global__ void __launch_bounds__(256)
pack_tips4(const uint8_t* __restrict__ tips, int64_t ld, const int* __restrict__ order, const int* __restrict__ xorder, int ntips,
           int nwords, uint64_t* __restrict__ packed, int64_t spad, int64_t site0, int64_t site1,
           int* __restrict__ bad_flag)
{
    __shared__ int s_order[PACK_ORDER_SMEM];
    const bool staged = ntips <= PACK_ORDER_SMEM;
    if (staged) {
        for (int i = threadIdx.x; i < ntips; i += blockDim.x) s_order[i] = order[i];
        __syncthreads();
    }
    const int* ord = staged ? s_order : order;
    const int64_t s = site0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // columns [site0, site1)
    if (s >= site1) return;
    unsigned bad = 0;
    for (int w = 0; w < nwords; w++) {
        unsigned st[32];
        int row[32];
#pragma unroll
        for (int j = 0; j < 32; j++) row[j] = ord[min(w * 32 + j, ntips - 1)];      // a negative row is a padding slot
#pragma unroll
        for (int j = 0; j < 32; j++) st[j] = tips[(int64_t)max(row[j], 0) * ld + s];
        if (xorder) {                                  // (every partner row is also a slot of its own: range-checked there)
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int xr = xorder[min(w * 32 + j, ntips - 1)];
                bad |= (row[j] >= 0 && w * 32 + j < ntips) ? st[j] : 0u;
                if (xr >= 0) st[j] ^= tips[(int64_t)xr * ld + s];
            }
        }
        uint64_t word = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            if (w * 32 + j < ntips && row[j] >= 0) {
                if (!xorder) bad |= st[j];
                word |= (uint64_t)(st[j] & 3u) << (2 * j);
            }
        }
        packed[(int64_t)w * spad + s] = word;
    }
    if (bad > 3u) atomicOr(bad_flag, 1);
}

// Nucleotide alignments that cross PCIe 2-bit packed (pb_set_tips_packed2 / pb_loglik_host_packed2): byte b of a
// row holds sites 4b..4b+3, site 4b+j in bits 2j..2j+1.  One thread per packed byte writes its four states
// as one 32-bit store into the uint8 tip matrix every kernel reads; columns [byte0, byte1) of every row.
__global__ void __launch_bounds__(256)
expand_tips2(const uint8_t* __restrict__ in, int64_t ld_in, uint8_t* __restrict__ tips, int64_t ld_out,
             int64_t byte0, int64_t byte1)
{
    const int row = blockIdx.x;          // rows in gridDim.x (no 65535 limit on the taxa), column blocks in gridDim.y
    for (int64_t b = byte0 + (int64_t)blockIdx.y * blockDim.x + threadIdx.x; b < byte1; b += (int64_t)gridDim.y * blockDim.x) {
        const unsigned v = in[(int64_t)row * ld_in + b];
        const unsigned out = (v & 3u) | ((v & 12u) << 6) | ((v & 48u) << 12) | ((v & 192u) << 18);
        *reinterpret_cast<unsigned*>(tips + (int64_t)row * ld_out + 4 * b) = out;
    }
}

// The two steps above in one, for the pipelined host evaluation (pb_loglik_host_packed2): 2-bit codes in the
// HOST layout (in[row][byte], 4 sites per byte) -> 2-bit codes in PROGRAM order (packed[word][site], 32 tips per
// word).  This is a transpose of 2-bit elements: a thread takes one 32-bit column of the upload (16 sites; a warp reads
// 128 consecutive bytes of a row), gathers the rows of 16 slots, transposes the 16 x 16 matrix of 2-bit elements in
// registers (four masked-swap stages, 0.75 instructions per element instead of 3 for shift-and-or) and, after both
// halves of a word, writes its 16 sites' words as four 32-byte stores.  Moves ntaxa/4 bytes in and ntaxa/4 (rounded
// up to whole words) bytes out per site.  grid (32-bit columns / 128, words): small blocks (128 threads, the 32 row
// indices of their word in shared memory) that run beside the resident blocks of the pruning kernel on a side stream.
template <int J>
__device__ __forceinline__ void transpose_stage(uint32_t (&A)[16])
{
    constexpr uint32_t m = J == 8 ? 0x0000FFFFu : J == 4 ? 0x00FF00FFu : J == 2 ? 0x0F0F0F0Fu : 0x33333333u;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k & J) continue;
        const uint32_t t = ((A[k] >> (2 * J)) ^ A[k + J]) & m;
        A[k] ^= t << (2 * J);
        A[k + J] ^= t;
    }
}

__device__ __forceinline__ void transpose16x16x2(uint32_t (&A)[16])
{
    // A[k] bits 2i..2i+1 = element (k, i)  ->  A[i] bits 2k..2k+1 = element (k, i)
    transpose_stage<8>(A);
    transpose_stage<4>(A);
    transpose_stage<2>(A);
    transpose_stage<1>(A);
}

__global__ void __launch_bounds__(128)
pack_tips4_from2(const uint8_t* __restrict__ in, int64_t ld_in, const int* __restrict__ order, const int* __restrict__ xorder,
                 int ntips, int nwords, uint64_t* __restrict__ packed, int64_t spad, int64_t byte0, int64_t byte1)
{
    // rows of the upload are 256-byte aligned and byte0, byte1, ld_in multiples of 256 (chunk bounds are multiples of
    // 1024 sites): whole 32-bit columns
    __shared__ int s_row[32], s_xrow[32];
    const uint32_t* __restrict__ in32 = reinterpret_cast<const uint32_t*>(in);
    const int64_t ld32 = ld_in >> 2;
    const int64_t col = (byte0 >> 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = col < (byte1 >> 2);
    for (int w = blockIdx.y; w < nwords; w += gridDim.y) {
        __syncthreads();
        if (threadIdx.x < 32) {
            const int slot = w * 32 + threadIdx.x;
            s_row[threadIdx.x] = slot < ntips ? order[slot] : -1;                   // a negative row is a padding slot
            s_xrow[threadIdx.x] = (xorder && slot < ntips) ? xorder[slot] : -1;     // difference coding: XOR with this row
        }
        __syncthreads();
        if (!live) continue;
        uint32_t A[16], B[16];                         // slots 0..15 and 16..31 of the word
#pragma unroll
        for (int k = 0; k < 32; k++) {
            // branch-free: padding slots and tips without a difference row read row 0 and drop it (2-bit fields XOR
            // independently: 16 sites at once; the tips of a clade share their first tip's row, an L1 hit)
            const int r = s_row[k], xr = s_xrow[k];
            const uint32_t x = in32[(int64_t)max(r, 0) * ld32 + col];
            const uint32_t base = in32[(int64_t)max(xr, 0) * ld32 + col];
            const uint32_t v = r >= 0 ? (x ^ (xr >= 0 ? base : 0u)) : 0u;
            if (k < 16) A[k] = v; else B[k - 16] = v;
        }
        transpose16x16x2(A);
        transpose16x16x2(B);
        uint64_t* out = packed + (int64_t)w * spad + 16 * col;               // 128-byte aligned
#pragma unroll
        for (int i = 0; i < 16; i += 4)
            asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(out + i),
                         "l"((uint64_t)A[i] | (uint64_t)B[i] << 32), "l"((uint64_t)A[i + 1] | (uint64_t)B[i + 1] << 32),
                         "l"((uint64_t)A[i + 2] | (uint64_t)B[i + 2] << 32), "l"((uint64_t)A[i + 3] | (uint64_t)B[i + 3] << 32)
                         : "memory");
    }
}

// Flags tip states >= nstates (the level kernels index P with them).
__global__ void __launch_bounds__(256)
validate_tips(const uint8_t* __restrict__ tips, int64_t total, int nstates, int* __restrict__ bad_flag)
{
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        bad |= (tips[i] >= nstates);
    if (bad) atomicOr(bad_flag, 1);
}

// f4: alignment text -> tip states on the device.  text[row*ld + site*width .. + width) is the symbol
// of taxon `row` at `site`; width 1 goes through a 256-entry table (nucleotides, amino acids),
// width 3 (codons) through the nucleotide table and then a 64-entry sense-codon table.  255 marks
// a character outside the alphabet (flagged, like Nucleotide.of_char_exn's invalid_arg).
__global__ void __launch_bounds__(256)
encode_text(const unsigned char* __restrict__ text, int64_t ld, int width, const unsigned char* __restrict__ lut,
            uint8_t* __restrict__ tips, int64_t tips_ld, int64_t nsites, int* __restrict__ bad_flag)
{
    const int row = blockIdx.x;          // rows in gridDim.x, column blocks in gridDim.y
    int bad = 0;
    for (int64_t s = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; s < nsites; s += (int64_t)gridDim.y * blockDim.x) {
        const unsigned char* p = text + (int64_t)row * ld + s * width;
        unsigned st;
        if (width == 1) {
            st = lut[p[0]];
        } else {
            const unsigned a = lut[p[0]], b = lut[p[1]], c = lut[p[2]];
            st = (a > 3u || b > 3u || c > 3u) ? 255u : lut[256 + a * 16 + b * 4 + c];
        }
        bad |= (st == 255u);
        tips[(int64_t)row * tips_ld + s] = (uint8_t)st;
    }
    if (bad) atomicOr(bad_flag, 1);
}

// Scaling modes of the site-resident kernels (template parameter EXACT_LOG, historically a bool):
//   0  reference rule, v := (1/max v) v; the maxima are multiplied into (mantissa m, exponent e), one log at the root
//   1  reference rule, carry += log(max v) at every event, literally lib/phylo_ctmc.ml:33-40
//   2  same test (min v > thr keeps the vector), but the scale factor is the power of two 2^-k with
//      k = exponent(max v): exact (no rounding at all), branch-free, no division; the carry is k ln 2
constexpr int SCALE_PRODUCT = 0, SCALE_EXACT_LOG = 1, SCALE_POW2 = 2;
#ifndef POW2_BRANCH_FREE
#define POW2_BRANCH_FREE 1
#endif

template <int EXACT_LOG>
struct Scale {             // log-scale accumulator of one (site, category)
    double m;              // mode 1: running sum of logs; mode 0: mantissa product in [2^-400, 1]; mode 2: unused
    int e;
    __device__ __forceinline__ void init() { m = EXACT_LOG == 1 ? 0.0 : 1.0; e = 0; }
    __device__ __forceinline__ void add(double mx)
    {
        if (EXACT_LOG == 1) {
            m += log(mx);
        } else {
            m *= mx;
            if (__double2hiint(m) < 0x26F00000) {      // m < 2^-400 (also catches 0 / negatives)
                if (m > 0.0) { m *= 0x1p+400; e -= 400; }
            }
        }
    }
    __device__ __forceinline__ double value() const
    {
        if (EXACT_LOG == 2) return (double)e * 0.693147180559945309417232121458;
        return EXACT_LOG == 1 ? m : (log(m) + (double)e * 0.693147180559945309417232121458);
    }
};

// max of non-negative doubles through their bit patterns (monotone for x >= 0; a negative
// rounding residue compares below every positive value, as it should).
__device__ __forceinline__ double max_nonneg(double a, double b)
{
    return __longlong_as_double(max(__double_as_longlong(a), __double_as_longlong(b)));
}

// min v > thr  <=>  every component > thr (a NaN fails the test, like the reference's compare)
__device__ __forceinline__ bool keep4(const double (&v)[4], double thr)
{
    return (v[0] > thr) & (v[1] > thr) & (v[2] > thr) & (v[3] > thr);
}

template <int EXACT_LOG>
__device__ __forceinline__ void rescale4(double (&v)[4], Scale<EXACT_LOG>& sc)
{
    if (EXACT_LOG == 2) {
        const int h = max(max(__double2hiint(v[0]), __double2hiint(v[1])), max(__double2hiint(v[2]), __double2hiint(v[3])));
        const int k = (h >> 20) - 1023;
        const double f = __hiloint2double((1023 - k) << 20, 0);
        v[0] *= f; v[1] *= f; v[2] *= f; v[3] *= f;
        sc.e += k;
        return;
    }
    const double mx = max_nonneg(max_nonneg(v[0], v[1]), max_nonneg(v[2], v[3]));
    const double inv = 1.0 / mx;
    v[0] = inv * v[0]; v[1] = inv * v[1]; v[2] = inv * v[2]; v[3] = inv * v[3];
    sc.add(mx);
}

template <int EXACT_LOG>
__device__ __forceinline__ void shift4s(double (&v)[4], Scale<EXACT_LOG>& sc, double thr)
{
    if (EXACT_LOG == 2 && POW2_BRANCH_FREE) {                     // the same rule as in shift_all
        const int h0 = __double2hiint(v[0]), h1 = __double2hiint(v[1]), h2 = __double2hiint(v[2]), h3 = __double2hiint(v[3]);
        const int hmin = min(min(h0, h1), min(h2, h3)), hmax = max(max(h0, h1), max(h2, h3));
        const int d = (hmin > __double2hiint(thr)) ? 1023 : (hmax >> 20);
        const double f = __hiloint2double((2046 - d) << 20, 0);
        v[0] *= f; v[1] *= f; v[2] *= f; v[3] *= f;
        sc.e += d - 1023;
        return;
    }
    if (!keep4(v, thr)) rescale4(v, sc);
}

// SV.shift for the thread's SPT sites with ONE branch on the common path (no site rescales).
template <int SPT, int EXACT_LOG>
__device__ __forceinline__ void shift_all(double (&v)[SPT][4], Scale<EXACT_LOG> (&sc)[SPT], double thr)
{
    if (EXACT_LOG == 2 && POW2_BRANCH_FREE) {
        // Power-of-two scaling without a branch and without FP64 instructions in the test: non-negative doubles
        // order like their bit patterns, so the smallest / largest component are found on the high words (a
        // negative rounding residue counts as small).  keep <=> min v > thr, compared on the upper 32 bits
        // (v whose high word equals the threshold's is rescaled too: a sliver of 2^-20 relative width).
        const int thr_hi = __double2hiint(thr);
#pragma unroll
        for (int q = 0; q < SPT; q++) {
            const int h0 = __double2hiint(v[q][0]), h1 = __double2hiint(v[q][1]);
            const int h2 = __double2hiint(v[q][2]), h3 = __double2hiint(v[q][3]);
            const int hmin = min(min(h0, h1), min(h2, h3)), hmax = max(max(h0, h1), max(h2, h3));
            const int d = (hmin > thr_hi) ? 1023 : (hmax >> 20);        // biased exponent to take out (1023: none)
            const double f = __hiloint2double((2046 - d) << 20, 0);     // 2^(1023 - d)
            v[q][0] *= f; v[q][1] *= f; v[q][2] *= f; v[q][3] *= f;
            sc[q].e += d - 1023;
        }
        return;
    }
    bool keep[SPT];
    bool all = true;
#pragma unroll
    for (int q = 0; q < SPT; q++) { keep[q] = keep4(v[q], thr); all &= keep[q]; }
    if (!all) {
#pragma unroll
        for (int q = 0; q < SPT; q++)
            if (!keep[q]) rescale4(v[q], sc[q]);
    }
}

// Shared-memory loads by 32-bit shared address: the P matrices are indexed by an offset read from
// the instruction stream, and plain pointer arithmetic makes the compiler go through generic
// addressing (S2R SR_CgaCtaId + LEA per access).
__device__ __forceinline__ double lds_f64(unsigned addr)
{
    double x;
    asm("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(addr));
    return x;
}
__device__ __forceinline__ void lds_mat16(unsigned addr, double (&m)[16])
{
#pragma unroll
    for (int i = 0; i < 8; i++)
        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(m[2 * i]), "=d"(m[2 * i + 1]) : "r"(addr + 16u * i));
}
// column s of a leaf branch's matrix: leaf matrices are staged TRANSPOSED, so the column is
// 32 contiguous bytes
__device__ __forceinline__ void lds_col4(unsigned addr, int s, double (&c)[4])
{
    const unsigned a = addr + 32u * (unsigned)s;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c[0]), "=d"(c[1]) : "r"(a));
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c[2]), "=d"(c[3]) : "r"(a + 16u));
}

constexpr int FOP_PUSH_AFTER = 0x100;   // flag in the code word: push ACC after the instruction
// FUSED_PREAPPLY (off by default; host side checked on the CPU, kernel side not measured yet): a tabulated clade is
// consumed by exactly one parent instruction, which first applies the clade's branch matrix to it; the table can
// hold P . v instead of v and the parent skips that product.  The flags say which operand of FOP_ACC_POP /
// FOP_APPLY_MUL_TIP already carries its matrix.
constexpr int FOP_PRE_A = 0x200;        // operand A (the accumulator) is a pre-applied table value
constexpr int FOP_PRE_B = 0x400;        // operand B (the popped value) is a pre-applied table value
#ifndef FUSED_PREAPPLY
#define FUSED_PREAPPLY 0
#endif

// Persistent kernel: grid = (blocks_x, ncat); each block stages the P matrices of its category
// once and then loops over tiles of BLOCK*SPT sites.
template <int BLOCK, int SPT, int EXACT_LOG>
__global__ void __launch_bounds__(BLOCK)
fused4(const FusedOp2* __restrict__ prog, int nops, int stack_depth,
       const double* __restrict__ P, const uint8_t* __restrict__ is_leaf, int nnodes,
       const uint64_t* __restrict__ packed, int64_t spad, int64_t tile0, int64_t ntiles,
       const double* __restrict__ pi, double thr, int shift_at_root,
       double* __restrict__ site_cat, int* __restrict__ site_exp)
{
    extern __shared__ double sh[];
    double* shP = sh;                                       // [nnodes][16] of this category
    double* shS = sh + (size_t)nnodes * 16;                 // [stack_depth][4][BLOCK*SPT]
    FusedOp2* shOps = reinterpret_cast<FusedOp2*>(shS + (size_t)stack_depth * 4 * BLOCK * SPT);
    const int c = blockIdx.y, tid = threadIdx.x;
    {
        const double* Pg = P + (size_t)c * nnodes * 16;
        for (int idx = tid; idx < nnodes * 16; idx += BLOCK) {
            const int node = idx >> 4, e = idx & 15;           // leaf branches: transposed (column look-ups)
            shP[is_leaf[node] ? (node * 16 + (e & 3) * 4 + (e >> 2)) : idx] = Pg[idx];
        }
        const int4* src = reinterpret_cast<const int4*>(prog);
        int4* dst = reinterpret_cast<int4*>(shOps);
        for (int idx = tid; idx < nops; idx += BLOCK) dst[idx] = src[idx];
    }
    __syncthreads();
    const unsigned pbase = (unsigned)__cvta_generic_to_shared(shP);
    double pir[4];
#pragma unroll
    for (int i = 0; i < 4; i++) pir[i] = pi[i];

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {     // tiles [tile0, tile0 + ntiles)
        const int64_t s0 = (tile0 + tile) * (BLOCK * SPT) + tid;              // site of slot q: s0 + q*BLOCK
        double v[SPT][4];
        Scale<EXACT_LOG> sc[SPT];
        uint64_t cur[SPT], nxt[SPT];
        const uint64_t* pk = packed + s0;
#pragma unroll
        for (int q = 0; q < SPT; q++) {
            sc[q].init();
            cur[q] = pk[q * BLOCK];
            nxt[q] = pk[spad + q * BLOCK];
#pragma unroll
            for (int i = 0; i < 4; i++) v[q][i] = 1.0;
        }
        pk += 2 * spad;
        double* st = shS + tid;                               // stack pointer (this thread's column)
        const int4* opp = reinterpret_cast<const int4*>(shOps);
        for (int ip = 0; ip < nops; ip++) {
            const int4 opw = opp[ip];
            const unsigned PA = pbase + 8u * (unsigned)opw.y;
            const unsigned PB = pbase + 8u * (unsigned)opw.z;
            const int shA = opw.w & 0xff, shB = opw.w >> 8;
            const int code = opw.x & 0xff;
            if (code == FOP_APPLY_MUL_TIP) {
                double pm[16];
                lds_mat16(PA, pm);
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double y[4], t[4];
                    lds_col4(PB, (int)(cur[q] >> shB) & 3, t);
                    matvec4(pm, v[q], y);
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = y[i] * t[i];
                }
                shift_all<SPT, EXACT_LOG>(v, sc, thr);
            } else if (code == FOP_ACC_POP) {
                st -= 4 * BLOCK * SPT;
                double pm[16];
                lds_mat16(PA, pm);
                double y[SPT][4];
#pragma unroll
                for (int q = 0; q < SPT; q++) matvec4(pm, v[q], y[q]);
                lds_mat16(PB, pm);
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double x[4], z[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) x[i] = st[(i * SPT + q) * BLOCK];
                    matvec4(pm, x, z);
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = y[q][i] * z[i];
                }
                shift_all<SPT, EXACT_LOG>(v, sc, thr);
            } else if (code == FOP_TIP_TIP) {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double a[4], b[4];
                    lds_col4(PA, (int)(cur[q] >> shA) & 3, a);
                    lds_col4(PB, (int)(cur[q] >> shB) & 3, b);
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = a[i] * b[i];
                }
                shift_all<SPT, EXACT_LOG>(v, sc, thr);
            } else if (code == FOP_LOADW) {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    cur[q] = nxt[q];
                    nxt[q] = pk[q * BLOCK];                   // the array is padded by 2 words
                }
                pk += spad;
            } else if (code == FOP_TIP) {
#pragma unroll
                for (int q = 0; q < SPT; q++) lds_col4(PA, (int)(cur[q] >> shA) & 3, v[q]);
            } else if (code == FOP_APPLY) {
                double pm[16];
                lds_mat16(PA, pm);
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double y[4];
                    matvec4(pm, v[q], y);
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = y[i];
                }
            } else if (code == FOP_MUL_TIP) {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double t[4];
                    lds_col4(PA, (int)(cur[q] >> shA) & 3, t);
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = v[q][i] * t[i];
                    shift4s(v[q], sc[q], thr);
                }
            } else if (code == FOP_POP_MUL) {
                st -= 4 * BLOCK * SPT;
#pragma unroll
                for (int q = 0; q < SPT; q++) {
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = st[(i * SPT + q) * BLOCK] * v[q][i];
                    shift4s(v[q], sc[q], thr);
                }
            } else if (code == FOP_ROOT) {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = v[q][i] * pir[i];
                    if (shift_at_root) shift4s(v[q], sc[q], thr);
                    const double sum = ((v[q][0] + v[q][1]) + v[q][2]) + v[q][3];
                    // power-of-two mode: the likelihood leaves as (value, binary exponent); the logarithm is taken once
                    // per site, of the category mixture, by root_combine (no log here, no exp there)
                    if (EXACT_LOG == SCALE_POW2) {
                        site_cat[(int64_t)c * spad + s0 + q * BLOCK] = sum;
                        site_exp[(int64_t)c * spad + s0 + q * BLOCK] = sc[q].e;
                    } else
                        site_cat[(int64_t)c * spad + s0 + q * BLOCK] = log(sum) + sc[q].value();
                }
            }
            if (opw.x & FOP_PUSH_AFTER) {                     // also the stand-alone FOP_PUSH
#pragma unroll
                for (int q = 0; q < SPT; q++)
#pragma unroll
                    for (int i = 0; i < 4; i++) st[(i * SPT + q) * BLOCK] = v[q][i];
                st += 4 * BLOCK * SPT;
            }
        }
    }
}

// --------------------------------------------------------------------------
// fused4c: the same traversal with the INNER-branch matrices in constant memory.
//
// ncu on fused4 (profiles/r01c_ncu_full_fused4_v5_variants.txt) shows the LSU data pipe at 92 % of
// peak: broadcasting a 128-byte matrix from shared memory to the 32 lanes of a warp costs 32 lanes x
// 128 B of register-write bandwidth, whatever the bank pattern.  The instruction stream is uniform,
// so the matrix address is warp-uniform: from the constant bank the 16 doubles go to UNIFORM
// registers once per warp (LDCU) and the DFMA/DMUL take them as uniform operands - no per-lane
// traffic at all.  Leaf-branch matrices are looked up with a per-lane column index, which the
// constant path would serialise, so they stay in shared memory (transposed).
// The constant bank (64 KB) holds the inner matrices of as many rate categories as fit (4 categories
// of a 128-taxon tree: 4 x 127 x 128 B = 65 024 B); one launch covers those categories
// (blockIdx.y), further categories take further launches.  The matrices travel device -> constant
// bank in stream order (cudaMemcpyToSymbolAsync, device to device).
// --------------------------------------------------------------------------
// The instruction stream sits in the constant bank as well: the interpreter's fetch is then a uniform load
// (LDCU) and the matrix offsets are uniform BY CONSTRUCTION.  (With the program in shared memory the
// compiler has to prove the offsets warp-uniform to use uniform loads for the matrices; ptxas did for one
// form of the rescaling code and did not for another, a 20 % difference.)
constexpr int FUSED_CONST_MATS = 380;              // 48 640 bytes of inner matrices
constexpr int FUSED_CONST_OPS = 960;               // 15 360 bytes of program; together 64 000 of the 65 536 bytes
__constant__ double cPinner[FUSED_CONST_MATS * 16];
__constant__ int4 cProg[FUSED_CONST_OPS];

// d_Pc[c][k][16] = P[c][inner_nodes[k]][16]   (categories packed back to back: ninner matrices each)
__global__ void gather_inner_p(const double* __restrict__ P, const int* __restrict__ inner_nodes, int ninner,
                               int nnodes, double* __restrict__ Pc)
{
    const int c = blockIdx.y;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < ninner * 16; idx += gridDim.x * blockDim.x)
        Pc[((size_t)c * ninner + (idx >> 4)) * 16 + (idx & 15)] =
            P[((size_t)c * nnodes + inner_nodes[idx >> 4]) * 16 + (idx & 15)];
}

__device__ __forceinline__ void matvec4c(int off, const double (&x)[4], double (&y)[4])
{
    const double* m = cPinner + off;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double a = m[i * 4 + 0] * x[0];
        a = fma(m[i * 4 + 1], x[1], a);
        a = fma(m[i * 4 + 2], x[2], a);
        a = fma(m[i * 4 + 3], x[3], a);
        y[i] = a;
    }
}

// Small clades are tabulated.  A subtree with k leaves can only take 4^k values per rate category, whatever the
// site: a node with two leaf children (a "cherry", a third of the internal nodes of a birth-death tree) 16, the
// parent of a cherry and a leaf 64, a clade of four leaves 256, ...  Once per evaluation every such clade (up to
// `fused_clade_leaves` leaves, all of them in one packed-tip word) gets a table of its possible conditional
// likelihood vectors and scale exponents, computed by INTERPRETING THE CLADE'S OWN INSTRUCTIONS of the program with
// the kernel's arithmetic — same operation order, same FMAs, same power-of-two rescaling, so the kernel's results
// do not change by a bit — and the kernel replaces the whole clade (column look-ups, matrix-vector products,
// rescalings, stack traffic) by ONE 32-byte look-up and an integer add (FOP_CHERRY).  The leaves of a clade are
// consumed back to back, so the table index is one shift and mask of the packed-tip window.
// One table per category, `ntab` entries of (4 doubles, 1 exponent); clade t owns entries [ent0, ent0 + 4^k).
// fold_flag != nullptr (the reduced program's tables): an entry's scale exponent is folded into its vector — v 2^k, an
// exact multiplication — whenever |k| <= 480, so that a look-up is ONE 32-byte gather instead of a gather of the vector
// and a gather of the exponent from another cache line (the exponent gather costs the L1 as many wavefronts as the
// vector's).  The consumer's rescaling rule restores the range; every factor is a power of two, so the likelihood is
// the same number whichever way it is split into value and exponent (root_combine normalises the split).  An entry
// that does not fit keeps (v, k) and raises *fold_flag: the kernel then gathers the exponents too.
__device__ __forceinline__ void fold_exponent(double (&v)[4], int& k, int* fold_flag)
{
    if (!fold_flag || k == 0) return;
    if (k >= -480 && k <= 480) {
        const double f = __hiloint2double((1023 + k) << 20, 0);
        v[0] *= f; v[1] *= f; v[2] *= f; v[3] *= f;
        k = 0;
    } else
        atomicOr(fold_flag, 1);
}

struct CladeTab {
    int op0, op1;      // the clade's instructions in the untabulated program (inclusive)
    int sh0, k;        // bit offset of its first tip in the window, number of leaves
    int ent0;
    int pre;           // >= 0: the table stores P[inner pre / 16] . v instead of v (FUSED_PREAPPLY), else -1
    int xor_coded;     // the index holds tip 0 and the other tips XOR tip 0 (reduced program, pack_tips4)
    int radix;         // entries = radix^k: 4 states, or the alignment's distinct state sets (build_clade_tables_masks)
};

template <bool DIRECT>
__global__ void __launch_bounds__(256)
build_clade_tables(const FusedOp2* __restrict__ prog, const CladeTab* __restrict__ tabs, const double* __restrict__ P,
                   const int* __restrict__ inner_nodes, const int* __restrict__ leaf_nodes, int nnodes, int ntab, double thr,
                   double* __restrict__ tabV, int* __restrict__ tabK, int* __restrict__ fold_flag = nullptr)
{
    const CladeTab tb = tabs[blockIdx.x];
    const int c = blockIdx.y;
    const double* Pc = P + (size_t)c * nnodes * 16;
    for (int idx = blockIdx.z * blockDim.x + threadIdx.x; idx < (1 << (2 * tb.k)); idx += gridDim.z * blockDim.x) {
        uint64_t cur = (uint64_t)idx << tb.sh0;                // the window as the kernel would see it
        if (tb.xor_coded) {   // the reduced program's fields hold tip 0 and the other tips XOR tip 0 (pack_tips4): undo
            const uint64_t rep = (uint64_t)(idx & 3) * 0x5555555555555555ull;
            cur ^= rep & ~3ull & ((1ull << (2 * tb.k)) - 1ull);
        }
        double v[4] = {1.0, 1.0, 1.0, 1.0}, stk[8][4];
        int sp = 0;
        Scale<SCALE_POW2> sc;
        sc.init();
        for (int ip = tb.op0; ip <= tb.op1; ip++) {
            const FusedOp2 o = prog[ip];
            const int code = o.code & 0xff, shA = o.sh & 0xff, shB = o.sh >> 8;
            // inner matrices are row-major (matvec4, the expression of matvec4c); a leaf's column s is P[i*4 + s]
            // (DIRECT: the clade mini-programs of the reduced program address matrices by node id)
            auto inner = [&](int off) { return Pc + (DIRECT ? (size_t)off : (size_t)inner_nodes[off / 16] * 16); };
            auto leaf = [&](int off) { return Pc + (DIRECT ? (size_t)off : (size_t)leaf_nodes[off / 16] * 16); };
            const int tA = (int)(cur >> shA) & 3, tB = (int)(cur >> shB) & 3;
            if (code == FOP_APPLY_MUL_TIP) {
                const double* lf = leaf(o.offB);
                double y[4];
                matvec4(inner(o.offA), v, y);
#pragma unroll
                for (int i = 0; i < 4; i++) v[i] = y[i] * lf[i * 4 + tB];
                shift4s(v, sc, thr);
            } else if (code == FOP_ACC_POP) {
                double y[4], z[4];
                sp--;
                matvec4(inner(o.offA), v, y);
                matvec4(inner(o.offB), stk[sp], z);
#pragma unroll
                for (int i = 0; i < 4; i++) v[i] = y[i] * z[i];
                shift4s(v, sc, thr);
            } else if (code == FOP_TIP_TIP) {
                const double* la = leaf(o.offA);
                const double* lb = leaf(o.offB);
#pragma unroll
                for (int i = 0; i < 4; i++) v[i] = la[i * 4 + tA] * lb[i * 4 + tB];
                shift4s(v, sc, thr);
            } else if (code == FOP_TIP) {
                const double* la = leaf(o.offA);
#pragma unroll
                for (int i = 0; i < 4; i++) v[i] = la[i * 4 + tA];
            } else if (code == FOP_APPLY) {
                double y[4];
                matvec4(inner(o.offA), v, y);
#pragma unroll
                for (int i = 0; i < 4; i++) v[i] = y[i];
            } else if (code == FOP_MUL_TIP) {
                const double* la = leaf(o.offA);
#pragma unroll
                for (int i = 0; i < 4; i++) v[i] = v[i] * la[i * 4 + tA];
                shift4s(v, sc, thr);
            } else if (code == FOP_POP_MUL) {
                sp--;
#pragma unroll
                for (int i = 0; i < 4; i++) v[i] = stk[sp][i] * v[i];
                shift4s(v, sc, thr);
            }
            if ((o.code & FOP_PUSH_AFTER) && ip < tb.op1) {     // (the last instruction's push stays in the program)
#pragma unroll
                for (int i = 0; i < 4; i++) stk[sp][i] = v[i];
                sp++;
            }
        }
        if (tb.pre >= 0) {                                      // the consumer's branch matrix, applied here once
            double y[4];
            matvec4(Pc + (DIRECT ? (size_t)tb.pre : (size_t)inner_nodes[tb.pre / 16] * 16), v, y);
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = y[i];
        }
        const size_t e = (size_t)c * ntab + tb.ent0 + idx;
        fold_exponent(v, sc.e, fold_flag);
#pragma unroll
        for (int i = 0; i < 4; i++) tabV[e * 4 + i] = v[i];
        tabK[e] = sc.e;
    }
}

// prog: offsets of inner matrices index cPinner (compact inner index * 16); offsets of leaf matrices
// index the shared-memory array of the ntips leaf matrices (compact leaf index * 16).
template <int BLOCK, int SPT, int EXACT_LOG>
__global__ void __launch_bounds__(BLOCK, BLOCK == 256 ? (SPT <= 2 ? 3 : 2) : (BLOCK == 128 ? (SPT <= 2 ? 6 : 4) : 1))
fused4c(const FusedOp2* __restrict__ prog, int nops, int stack_depth,
        const double* __restrict__ P, const int* __restrict__ leaf_nodes, int nleaf, int nnodes, int cat0, int slot0, int ninner,
        const uint64_t* __restrict__ packed, int64_t spad, int64_t tile0, int64_t ntiles,
        const double* __restrict__ pi, double thr, int shift_at_root,
        const double* __restrict__ tabV, const int* __restrict__ tabK, int ntab,
        double* __restrict__ site_cat, int* __restrict__ site_exp)
{
    extern __shared__ double sh[];
    double* shP = sh;                                       // [nleaf][16], transposed, of this category
    double* shS = sh + (size_t)nleaf * 16;                  // [stack_depth][4][BLOCK*SPT]
    const int c = cat0 + blockIdx.y, tid = threadIdx.x;     // this launch covers categories cat0 .. cat0 + gridDim.y - 1
    const int cbase = (slot0 + blockIdx.y) * ninner * 16;   // where this category's inner matrices sit in cPinner
    {
        const double* Pg = P + (size_t)c * nnodes * 16;
        for (int idx = tid; idx < nleaf * 16; idx += BLOCK) {
            const int k = idx >> 4, e = idx & 15;
            shP[k * 16 + (e & 3) * 4 + (e >> 2)] = Pg[(size_t)leaf_nodes[k] * 16 + e];
        }
    }
    __syncthreads();
    const unsigned pbase = (unsigned)__cvta_generic_to_shared(shP);
    double pir[4];
#pragma unroll
    for (int i = 0; i < 4; i++) pir[i] = pi[i];

    // (Dealing WARP tiles of 32 * SPT sites round-robin over blocks and warps, so that a short last round spreads
    // over every SM, was measured slower: 1.34 against 1.22 ms at c2, profiles/r01m_ab_fused4c_variants.jsonl.)
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t s0 = (tile0 + tile) * (BLOCK * SPT) + tid;
        double v[SPT][4];
        Scale<EXACT_LOG> sc[SPT];
        uint64_t cur[SPT], nxt[SPT];
        const uint64_t* pk = packed + s0;
#pragma unroll
        for (int q = 0; q < SPT; q++) {
            sc[q].init();
            cur[q] = pk[q * BLOCK];
            nxt[q] = pk[spad + q * BLOCK];
#pragma unroll
            for (int i = 0; i < 4; i++) v[q][i] = 1.0;
        }
        pk += 2 * spad;
        double* st = shS + tid;
        for (int ip = 0; ip < nops; ip++) {
            const int4 opw = cProg[ip];                      // uniform fetch (fetching one instruction ahead was
                                                             // measured slower: 1.46 against 1.23 ms at c2)
            const int offA = opw.y, offB = opw.z;
            const int shA = opw.w & 0xff, shB = opw.w >> 8;
            const int code = opw.x & 0xff;
            if (code == FOP_APPLY_MUL_TIP) {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double y[4], t[4];
                    lds_col4(pbase + 8u * (unsigned)offB, (int)(cur[q] >> shB) & 3, t);
#if FUSED_PREAPPLY
                    if (opw.x & FOP_PRE_A) {
#pragma unroll
                        for (int i = 0; i < 4; i++) y[i] = v[q][i];
                    } else
#endif
                    matvec4c(cbase + offA, v[q], y);
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = y[i] * t[i];
                }
                shift_all<SPT, EXACT_LOG>(v, sc, thr);
            } else if (code == FOP_ACC_POP) {
                st -= 4 * BLOCK * SPT;
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double x[4], y[4], z[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) x[i] = st[(i * SPT + q) * BLOCK];
#if FUSED_PREAPPLY
                    if (opw.x & FOP_PRE_A) {
#pragma unroll
                        for (int i = 0; i < 4; i++) y[i] = v[q][i];
                    } else {
                        matvec4c(cbase + offA, v[q], y);
                    }
                    if (opw.x & FOP_PRE_B) {
#pragma unroll
                        for (int i = 0; i < 4; i++) z[i] = x[i];
                    } else {
                        matvec4c(cbase + offB, x, z);
                    }
#else
                    matvec4c(cbase + offA, v[q], y);
                    matvec4c(cbase + offB, x, z);
#endif
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = y[i] * z[i];
                }
                shift_all<SPT, EXACT_LOG>(v, sc, thr);
            } else if (code == FOP_CHERRY) {
                if (EXACT_LOG == 2) {                        // (only the power-of-two mode tabulates): offA = first entry, offB = index mask
                    const double2* tv = reinterpret_cast<const double2*>(tabV) + ((size_t)c * ntab + offA) * 2;
                    const int* tk = tabK + ((size_t)c * ntab + offA);
#pragma unroll
                    for (int q = 0; q < SPT; q++) {
                        const int idx = (int)(cur[q] >> shA) & offB;
                        const double2 lo = __ldg(tv + idx * 2), hi = __ldg(tv + idx * 2 + 1);
                        v[q][0] = lo.x; v[q][1] = lo.y; v[q][2] = hi.x; v[q][3] = hi.y;
                        sc[q].e += __ldg(tk + idx);
                    }
                }
            } else if (code == FOP_TIP_TIP) {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double a[4], b[4];
                    lds_col4(pbase + 8u * (unsigned)offA, (int)(cur[q] >> shA) & 3, a);
                    lds_col4(pbase + 8u * (unsigned)offB, (int)(cur[q] >> shB) & 3, b);
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = a[i] * b[i];
                }
                shift_all<SPT, EXACT_LOG>(v, sc, thr);
            } else if (code == FOP_LOADW) {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    cur[q] = nxt[q];
                    nxt[q] = pk[q * BLOCK];
                }
                pk += spad;
            } else if (code == FOP_TIP) {
#pragma unroll
                for (int q = 0; q < SPT; q++) lds_col4(pbase + 8u * (unsigned)offA, (int)(cur[q] >> shA) & 3, v[q]);
            } else if (code == FOP_APPLY) {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double y[4];
                    matvec4c(cbase + offA, v[q], y);
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = y[i];
                }
            } else if (code == FOP_MUL_TIP) {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double t[4];
                    lds_col4(pbase + 8u * (unsigned)offA, (int)(cur[q] >> shA) & 3, t);
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = v[q][i] * t[i];
                    shift4s(v[q], sc[q], thr);
                }
            } else if (code == FOP_POP_MUL) {
                st -= 4 * BLOCK * SPT;
#pragma unroll
                for (int q = 0; q < SPT; q++) {
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = st[(i * SPT + q) * BLOCK] * v[q][i];
                    shift4s(v[q], sc[q], thr);
                }
            } else if (code == FOP_ROOT) {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = v[q][i] * pir[i];
                    if (shift_at_root) shift4s(v[q], sc[q], thr);
                    const double sum = ((v[q][0] + v[q][1]) + v[q][2]) + v[q][3];
                    // power-of-two mode: the likelihood leaves as (value, binary exponent); the logarithm is taken once
                    // per site, of the category mixture, by root_combine (no log here, no exp there)
                    if (EXACT_LOG == SCALE_POW2) {
                        site_cat[(int64_t)c * spad + s0 + q * BLOCK] = sum;
                        site_exp[(int64_t)c * spad + s0 + q * BLOCK] = sc[q].e;
                    } else
                        site_cat[(int64_t)c * spad + s0 + q * BLOCK] = log(sum) + sc[q].value();
                }
            }
            if (opw.x & FOP_PUSH_AFTER) {
#pragma unroll
                for (int q = 0; q < SPT; q++)
#pragma unroll
                    for (int i = 0; i < 4; i++) st[(i * SPT + q) * BLOCK] = v[q][i];
                st += 4 * BLOCK * SPT;
            }
        }
    }
}

// --------------------------------------------------------------------------
// fused4t: the program over the REDUCED tree.
//
// fused4c replaces a tabulated clade's instructions by a look-up but keeps the push / pop structure of the full
// tree, and an ncu capture (profiles/r02a_*) shows what that costs: the L1 data pipe is the limiter (71 % of its
// wavefront slots), 40 % of it stack traffic in shared memory (a push and a pop of 128 bytes per site for every
// node with two inner children, 32 at the bench shape although most of those children are table values) and
// 60 % table gathers (two 16-byte and one 4-byte gather per look-up).  Here the host plans the program over the
// tree whose leaves are the tabulated clades and the loose tips (pb_api.cu: build_reduced_program):
//   * a table value is an OPERAND of its parent's instruction (FOP_TAB_TAB, FOP_APPLY_MUL_TAB, FOP_TAB_TIP,
//     FOP_MUL_TAB), gathered where it is used: only nodes with two computed children push (7 instead of 32);
//   * the order of evaluation and the stack depth follow the reduced tree (smaller stack area);
//   * a table holds P . v, P the matrix of the branch above its clade (the consumer's product, done once per
//     table entry instead of once per site), and its vector is ONE 32-byte gather (LDG.E.256);
//   * only the matrices of the reduced tree's inner branches go to the constant bank and only the loose tips'
//     matrices to shared memory: trees of up to ~1 000 taxa fit the bank.
// The arithmetic of every node is what fused4c does in the power-of-two mode, in the same order: per-site results
// are bit-identical (tests/test_parity_gpu.py::test_fused_reduced_program_*).
// --------------------------------------------------------------------------
constexpr int FOP_TAB = 11;            // ACC = T[offA][field A]                       (carry += its exponent; no shift)
constexpr int FOP_TAB_TAB = 12;        // ACC = shiftmul(T[offA][field A], T[offB][field B])
constexpr int FOP_TAB_TIP = 13;        // ACC = shiftmul(T[offA][field A], col(leaf offB, tip B))
constexpr int FOP_APPLY_MUL_TAB = 14;  // ACC = shiftmul(P[offA] . ACC, T[offB][field B])
constexpr int FOP_MUL_TAB = 15;        // ACC = shiftmul(ACC, T[offA][field A])          (arity >= 3)
// sh word of these instructions: shA | shB << 8 | bitsA << 16 | bitsB << 24  (a field is `bits` wide at bit `sh`
// of the current window word; a tip is a 2-bit field)

// (FOLD: every exponent of this evaluation's tables is folded into its vector, see fold_exponent)
template <bool FOLD>
__device__ __forceinline__ void tab_gather(const double* __restrict__ tv, const int* __restrict__ tk, int ent0,
                                           uint64_t cur, int sh, unsigned mask, double (&v)[4], int& e)
{
    const unsigned idx = (unsigned)(cur >> sh) & mask;
    const double* p = tv + (size_t)(ent0 + idx) * 4;
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
    e = FOLD ? 0 : __ldg(tk + ent0 + idx);
}

// ---- tip state SETS in the site-resident kernel (Missing_values / Ambiguous, lib/phylo_ctmc.ml:145-172, :232-302) ----
// MASKS = true: a tip is one of the `radix` distinct state sets the alignment contains (code 0..radix-1; pb_set_tip_masks
// scans them), a tabulated clade of k tips has radix^k entries indexed in mixed radix, and a loose tip's column is the sum
// of the set's columns of P (dgemv against the 0/1 indicator, ascending j).  The empty set is a missing observation:
// "None".  None is carried in memory (table entries, leaf columns, the stack) by a marker, v[0] == -1, and in registers
// as one flag per site (below); the reference's rule (Missing_values.maybe_sv_mul :146-150, Ambiguous.pruning :289-296) is
//     None x None = None,  None x v = v (NO shift),  v x w = SV.mul v w (shift),
// and the branch matrix of a None subtree is never applied; a site with no observation at all has log-likelihood 0.
constexpr double NONE_MARK = -1.0;
// In MEMORY a None vector is (-1, 1, 1, 1): the marker in component 0 and ones elsewhere.  In REGISTERS a None operand is
// the vector of ones plus its flag — x * 1.0 is x exactly, so the rule above is a plain product and a flag AND, without a
// select per component: is_none reads the marker (an integer compare on the high word) and turns it into the 1.0 it
// stands for (one select; the low words of -1.0 and 1.0 are both zero); a matrix-vector product of a None accumulator
// (P . ones is only approximately ones) is put back by keep_none.
template <bool MASKS>
__device__ __forceinline__ bool is_none(double (&x)[4])
{
    if constexpr (MASKS) {
        const int hi = __double2hiint(x[0]);
        const bool none = hi == (int)0xBFF00000;
        x[0] = __hiloint2double(none ? 0x3FF00000 : hi, __double2loint(x[0]));
        return none;
    } else
        return false;
}

template <bool MASKS>
__device__ __forceinline__ void keep_none(double (&y)[4], bool none)
{
    if constexpr (MASKS) {
#pragma unroll
        for (int i = 0; i < 4; i++) y[i] = none ? 1.0 : y[i];
    }
}

// v = a (x) b under the rule above (operands normalised as described); returns whether the product was formed (only
// then SV.mul shifts)
template <bool MASKS>
__device__ __forceinline__ bool combine4(double (&v)[4], bool& nv, const double (&a)[4], bool na, const double (&b)[4], bool nb)
{
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = a[i] * b[i];
    nv = MASKS && na && nb;
    return !MASKS || !(na || nb);
}

// SV.shift with exact power-of-two factors (shift_all's rule), suppressed when `on` is false
__device__ __forceinline__ void shift_pow2(double (&v)[4], Scale<SCALE_POW2>& sc, int thr_hi, bool on)
{
    const int h0 = __double2hiint(v[0]), h1 = __double2hiint(v[1]), h2 = __double2hiint(v[2]), h3 = __double2hiint(v[3]);
    const int hmin = min(min(h0, h1), min(h2, h3)), hmax = max(max(h0, h1), max(h2, h3));
    const int d = (!on || hmin > thr_hi) ? 1023 : (hmax >> 20);
    const double f = __hiloint2double((2046 - d) << 20, 0);
    v[0] *= f; v[1] *= f; v[2] *= f; v[3] *= f;
    sc.e += d - 1023;
}

// Packed fields of the masked form: field f of a site = sum_j code(mask of tip row rows[first + j]) * radix^j, `bits`
// wide at bit `shift` of word `word` (pb_api.cu: build_reduced_program).  One thread per site.
struct PackField { int word, shift, first, k; };

__global__ void __launch_bounds__(256)
pack_fields_masks(const uint64_t* __restrict__ masks, int64_t ld, const PackField* __restrict__ fields, int nfields,
                  const int* __restrict__ rows, const unsigned char* __restrict__ code_of_mask /* [16] */, int radix, int nwords,
                  uint64_t* __restrict__ packed, int64_t spad, int64_t site0, int64_t site1)
{
    __shared__ unsigned char lut[16];
    if (threadIdx.x < 16) lut[threadIdx.x] = code_of_mask[threadIdx.x];
    __syncthreads();
    const int64_t s = site0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= site1) return;
    uint64_t word = 0;
    int cur_word = 0;
    for (int f = 0; f < nfields; f++) {
        const PackField fd = fields[f];
        if (fd.word != cur_word) {                      // fields are sorted by word
            packed[(int64_t)cur_word * spad + s] = word;
            for (int w = cur_word + 1; w < fd.word; w++) packed[(int64_t)w * spad + s] = 0;
            cur_word = fd.word;
            word = 0;
        }
        unsigned idx = 0;
        for (int j = fd.k - 1; j >= 0; j--)             // Horner: tip 0 is the lowest digit
            idx = idx * radix + lut[masks[(int64_t)rows[fd.first + j] * ld + s] & 15u];
        word |= (uint64_t)idx << fd.shift;
    }
    packed[(int64_t)cur_word * spad + s] = word;
    for (int w = cur_word + 1; w < nwords; w++) packed[(int64_t)w * spad + s] = 0;
}

// Which of the 16 possible state sets of a 4-state alphabet occur in the alignment: bit m of *present.
__global__ void __launch_bounds__(256)
mask_alphabet(const uint64_t* __restrict__ masks, int64_t ld, int64_t nsites, int ntaxa, unsigned* __restrict__ present)
{
    unsigned seen = 0;
    const int64_t total = (int64_t)ntaxa * nsites;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        seen |= 1u << (unsigned)(masks[(i / nsites) * ld + (i % nsites)] & 15u);
    seen = __reduce_or_sync(0xffffffffu, seen);
    if ((threadIdx.x & 31) == 0 && seen) atomicOr(present, seen);
}

// The tables of the masked form: entry idx of a clade decodes (mixed radix) to one state set per tip; the clade's own
// instruction range is interpreted with the None rule above and set-column sums.  mask_of_code[c] = the state set of code c.
__global__ void __launch_bounds__(256)
build_clade_tables_masks(const FusedOp2* __restrict__ prog, const CladeTab* __restrict__ tabs, const double* __restrict__ P,
                         int nnodes, int ntab, double thr, int radix, const unsigned char* __restrict__ mask_of_code,
                         double* __restrict__ tabV, int* __restrict__ tabK, int* __restrict__ fold_flag = nullptr)
{
    const CladeTab tb = tabs[blockIdx.x];
    const int c = blockIdx.y;
    const double* Pc = P + (size_t)c * nnodes * 16;
    int nent = 1;
    for (int j = 0; j < tb.k; j++) nent *= radix;
    const int thr_hi = __double2hiint(thr);
    for (int idx = blockIdx.z * blockDim.x + threadIdx.x; idx < nent; idx += gridDim.z * blockDim.x) {
        unsigned sets[8];                                        // state set of every tip of the clade
        {
            int r = idx;
            for (int j = 0; j < tb.k; j++) { sets[j] = mask_of_code[r % radix]; r /= radix; }
        }
        auto column = [&](int off, int tip, double (&t)[4]) {   // P[leaf] . indicator(set): set columns added, ascending
            const double* lf = Pc + off;
            const unsigned m = sets[tip];
            if (m == 0) { t[0] = NONE_MARK; t[1] = t[2] = t[3] = 1.0; return; }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if ((m >> j) & 1u) a += lf[i * 4 + j];
                t[i] = a;
            }
        };
        double v[4] = {1.0, 1.0, 1.0, 1.0}, stk[8][4];
        bool nv = true;                                          // nothing seen yet
        int sp = 0;
        Scale<SCALE_POW2> sc;
        sc.init();
        for (int ip = tb.op0; ip <= tb.op1; ip++) {
            const FusedOp2 o = prog[ip];
            const int code = o.code & 0xff, tA = (o.sh & 0xff) >> 1, tB = ((o.sh >> 8) & 0xff) >> 1;   // tip ordinals
            if (code == FOP_APPLY_MUL_TIP) {
                double y[4], t[4];
                matvec4(Pc + o.offA, v, y);
                keep_none<true>(y, nv);
                column(o.offB, tB, t);
                if (combine4<true>(v, nv, y, nv, t, is_none<true>(t))) shift_pow2(v, sc, thr_hi, true);
            } else if (code == FOP_ACC_POP) {
                double y[4], z[4];
                sp--;
                const bool nb = is_none<true>(stk[sp]);
                matvec4(Pc + o.offA, v, y);
                matvec4(Pc + o.offB, stk[sp], z);
                keep_none<true>(y, nv);
                keep_none<true>(z, nb);
                if (combine4<true>(v, nv, y, nv, z, nb)) shift_pow2(v, sc, thr_hi, true);
            } else if (code == FOP_TIP_TIP) {
                double a[4], b[4];
                column(o.offA, tA, a);
                column(o.offB, tB, b);
                if (combine4<true>(v, nv, a, is_none<true>(a), b, is_none<true>(b))) shift_pow2(v, sc, thr_hi, true);
            } else if (code == FOP_TIP) {
                column(o.offA, tA, v);
                nv = is_none<true>(v);
            } else if (code == FOP_APPLY) {
                double y[4];
                matvec4(Pc + o.offA, v, y);
                if (!nv) { v[0] = y[0]; v[1] = y[1]; v[2] = y[2]; v[3] = y[3]; }
            } else if (code == FOP_MUL_TIP) {
                double t[4], a[4] = {v[0], v[1], v[2], v[3]};
                column(o.offA, tA, t);
                if (combine4<true>(v, nv, a, nv, t, is_none<true>(t))) shift_pow2(v, sc, thr_hi, true);
            } else if (code == FOP_POP_MUL) {
                sp--;
                double a[4] = {v[0], v[1], v[2], v[3]};
                if (combine4<true>(v, nv, stk[sp], is_none<true>(stk[sp]), a, nv)) shift_pow2(v, sc, thr_hi, true);
            }
            if ((o.code & FOP_PUSH_AFTER) && ip < tb.op1) {
                stk[sp][0] = nv ? NONE_MARK : v[0];
                stk[sp][1] = v[1]; stk[sp][2] = v[2]; stk[sp][3] = v[3];
                sp++;
            }
        }
        if (tb.pre >= 0 && !nv) {                                // the consumer's branch matrix, applied here once
            double y[4];
            matvec4(Pc + tb.pre, v, y);
            v[0] = y[0]; v[1] = y[1]; v[2] = y[2]; v[3] = y[3];
        }
        const size_t e = (size_t)c * ntab + tb.ent0 + idx;
        if (!nv) fold_exponent(v, sc.e, fold_flag);
        tabV[e * 4 + 0] = nv ? NONE_MARK : v[0];
        tabV[e * 4 + 1] = v[1]; tabV[e * 4 + 2] = v[2]; tabV[e * 4 + 3] = v[3];
        tabK[e] = nv ? 0 : sc.e;
    }
}

// tip_stride: doubles per loose tip in shared memory (16, or 4 * radix with MASKS); tip_mask: (1 << bits of a tip field) - 1
// FOLD: the body for tables whose exponents are all folded into their vectors (no exponent gather, no exponent
// arithmetic).  Whether they are is known on the device only (*fold_flag, written by the table builder just before):
// the kernel reads the flag and runs the FOLD body inline, or calls the general body (a separate function, so that
// the common path keeps its own register allocation).
#define FUSED4T_PARAMS                                                                                                      \
    int nops, int stack_depth, const double *__restrict__ P, const int *__restrict__ leaf_nodes, int nleaf, int nnodes,    \
        int cat0, int slot0, int nmat, const uint64_t *__restrict__ packed, int64_t spad, int64_t tile0, int64_t ntiles,    \
        const double *__restrict__ pi, double thr, int shift_at_root, const double *__restrict__ tabV,                     \
        const int *__restrict__ tabK, int ntab, double *__restrict__ site_cat, int *__restrict__ site_exp, int radix,      \
        unsigned tip_mask, const unsigned char *__restrict__ mask_of_code
#define FUSED4T_ARGS                                                                                                        \
    nops, stack_depth, P, leaf_nodes, nleaf, nnodes, cat0, slot0, nmat, packed, spad, tile0, ntiles, pi, thr, shift_at_root, \
        tabV, tabK, ntab, site_cat, site_exp, radix, tip_mask, mask_of_code

template <int BLOCK, int SPT, bool MASKS, bool FOLD>
__device__ __forceinline__ void fused4t_body(FUSED4T_PARAMS)
{
    extern __shared__ double sh[];
    double* shP = sh;                                       // loose tips' columns: [nleaf][16] transposed, or [nleaf][radix][4]
    const int tip_stride = MASKS ? 4 * radix : 16;
    double* shS = sh + (size_t)nleaf * tip_stride;          // [stack_depth][4][BLOCK*SPT]
    const int c = cat0 + blockIdx.y, tid = threadIdx.x;
    const int cbase = (slot0 + blockIdx.y) * nmat * 16;     // this category's matrices in cPinner
    {
        const double* Pg = P + (size_t)c * nnodes * 16;
        if (!MASKS) {
            for (int idx = tid; idx < nleaf * 16; idx += BLOCK) {
                const int k = idx >> 4, e = idx & 15;
                shP[k * 16 + (e & 3) * 4 + (e >> 2)] = Pg[(size_t)leaf_nodes[k] * 16 + e];
            }
        } else {
            for (int idx = tid; idx < nleaf * radix * 4; idx += BLOCK) {     // column of state-set code cc, row i
                const int k = idx / (radix * 4), cc = (idx >> 2) % radix, i = idx & 3;
                const unsigned m = mask_of_code[cc];
                const double* lf = Pg + (size_t)leaf_nodes[k] * 16 + i * 4;
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if ((m >> j) & 1u) a += lf[j];
                shP[idx] = m ? a : (i == 0 ? NONE_MARK : 1.0);
            }
        }
    }
    __syncthreads();
    const unsigned pbase = (unsigned)__cvta_generic_to_shared(shP);
    const double* tv = tabV + (size_t)c * ntab * 4;
    const int* tk = tabK + (size_t)c * ntab;
    const int thr_hi = __double2hiint(thr);
    double pir[4];
#pragma unroll
    for (int i = 0; i < 4; i++) pir[i] = pi[i];

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t s0 = (tile0 + tile) * (BLOCK * SPT) + tid;
        double v[SPT][4];
        bool nv[SPT];                                       // the accumulator is None (MASKS only)
        Scale<SCALE_POW2> sc[SPT];
        uint64_t cur[SPT], nxt[SPT];
        const uint64_t* pk = packed + s0;
#pragma unroll
        for (int q = 0; q < SPT; q++) {
            sc[q].init();
            nv[q] = false;
            cur[q] = pk[q * BLOCK];
            nxt[q] = pk[spad + q * BLOCK];
#pragma unroll
            for (int i = 0; i < 4; i++) v[q][i] = 1.0;
        }
        pk += 2 * spad;
        double* st = shS + tid;
        for (int ip = 0; ip < nops; ip++) {
            const int4 opw = cProg[ip];
            const int offA = opw.y, offB = opw.z;
            const int shA = opw.w & 0xff, shB = (opw.w >> 8) & 0xff;
            const unsigned mA = (1u << ((opw.w >> 16) & 0xff)) - 1u, mB = (1u << ((unsigned)opw.w >> 24)) - 1u;
            const int code = opw.x & 0xff;
            switch (code) {            // (a dense switch: one indexed uniform branch instead of a chain of compares)
            case FOP_APPLY_MUL_TAB: {
                double b[SPT][4];
                int eb[SPT];
#pragma unroll
                for (int q = 0; q < SPT; q++) tab_gather<FOLD>(tv, tk, offB, cur[q], shB, mB, b[q], eb[q]);
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double y[4];
                    matvec4c(cbase + offA, v[q], y);
                    keep_none<MASKS>(y, nv[q]);
                    const bool on = combine4<MASKS>(v[q], nv[q], y, nv[q], b[q], is_none<MASKS>(b[q]));
                    sc[q].e += eb[q];
                    shift_pow2(v[q], sc[q], thr_hi, on);
                }
            } break;
            case FOP_APPLY_MUL_TIP: {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double y[4], t[4];
                    lds_col4(pbase + 8u * (unsigned)offB, (int)((unsigned)(cur[q] >> shB) & tip_mask), t);
                    matvec4c(cbase + offA, v[q], y);
                    keep_none<MASKS>(y, nv[q]);
                    const bool on = combine4<MASKS>(v[q], nv[q], y, nv[q], t, is_none<MASKS>(t));
                    shift_pow2(v[q], sc[q], thr_hi, on);
                }
            } break;
            case FOP_TAB_TAB: {
                double b[SPT][4];
                int ea[SPT], eb[SPT];
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    tab_gather<FOLD>(tv, tk, offA, cur[q], shA, mA, v[q], ea[q]);      // (operand A straight into the accumulator)
                    tab_gather<FOLD>(tv, tk, offB, cur[q], shB, mB, b[q], eb[q]);
                }
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    const bool on = combine4<MASKS>(v[q], nv[q], v[q], is_none<MASKS>(v[q]), b[q], is_none<MASKS>(b[q]));
                    sc[q].e += ea[q] + eb[q];
                    shift_pow2(v[q], sc[q], thr_hi, on);
                }
            } break;
            case FOP_ACC_POP: {
                st -= 4 * BLOCK * SPT;
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double x[4], y[4], z[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) x[i] = st[(i * SPT + q) * BLOCK];
                    const bool nb = is_none<MASKS>(x);
                    matvec4c(cbase + offA, v[q], y);
                    matvec4c(cbase + offB, x, z);
                    keep_none<MASKS>(y, nv[q]);
                    keep_none<MASKS>(z, nb);
                    const bool on = combine4<MASKS>(v[q], nv[q], y, nv[q], z, nb);
                    shift_pow2(v[q], sc[q], thr_hi, on);
                }
            } break;
            case FOP_TAB_TIP: {
                int ea[SPT];
#pragma unroll
                for (int q = 0; q < SPT; q++) tab_gather<FOLD>(tv, tk, offA, cur[q], shA, mA, v[q], ea[q]);
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double t[4];
                    lds_col4(pbase + 8u * (unsigned)offB, (int)((unsigned)(cur[q] >> shB) & tip_mask), t);
                    const bool on = combine4<MASKS>(v[q], nv[q], v[q], is_none<MASKS>(v[q]), t, is_none<MASKS>(t));
                    sc[q].e += ea[q];
                    shift_pow2(v[q], sc[q], thr_hi, on);
                }
            } break;
            case FOP_LOADW: {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    cur[q] = nxt[q];
                    nxt[q] = pk[q * BLOCK];
                }
                pk += spad;
            } break;
            case FOP_TAB: {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    int ea;
                    tab_gather<FOLD>(tv, tk, offA, cur[q], shA, mA, v[q], ea);
                    nv[q] = is_none<MASKS>(v[q]);
                    sc[q].e += ea;
                }
            } break;
            case FOP_MUL_TAB: {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double a[4], w[4] = {v[q][0], v[q][1], v[q][2], v[q][3]};
                    int ea;
                    tab_gather<FOLD>(tv, tk, offA, cur[q], shA, mA, a, ea);
                    const bool on = combine4<MASKS>(v[q], nv[q], w, nv[q], a, is_none<MASKS>(a));
                    sc[q].e += ea;
                    shift_pow2(v[q], sc[q], thr_hi, on);
                }
            } break;
            case FOP_TIP_TIP: {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double a[4], b[4];
                    lds_col4(pbase + 8u * (unsigned)offA, (int)((unsigned)(cur[q] >> shA) & tip_mask), a);
                    lds_col4(pbase + 8u * (unsigned)offB, (int)((unsigned)(cur[q] >> shB) & tip_mask), b);
                    const bool on = combine4<MASKS>(v[q], nv[q], a, is_none<MASKS>(a), b, is_none<MASKS>(b));
                    shift_pow2(v[q], sc[q], thr_hi, on);
                }
            } break;
            case FOP_TIP: {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    lds_col4(pbase + 8u * (unsigned)offA, (int)((unsigned)(cur[q] >> shA) & tip_mask), v[q]);
                    nv[q] = is_none<MASKS>(v[q]);
                }
            } break;
            case FOP_APPLY: {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double y[4];
                    matvec4c(cbase + offA, v[q], y);
                    if (!(MASKS && nv[q])) {
#pragma unroll
                        for (int i = 0; i < 4; i++) v[q][i] = y[i];
                    }
                }
            } break;
            case FOP_MUL_TIP: {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double t[4], w[4] = {v[q][0], v[q][1], v[q][2], v[q][3]};
                    lds_col4(pbase + 8u * (unsigned)offA, (int)((unsigned)(cur[q] >> shA) & tip_mask), t);
                    const bool on = combine4<MASKS>(v[q], nv[q], w, nv[q], t, is_none<MASKS>(t));
                    shift_pow2(v[q], sc[q], thr_hi, on);
                }
            } break;
            case FOP_POP_MUL: {
                st -= 4 * BLOCK * SPT;
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    double x[4], w[4] = {v[q][0], v[q][1], v[q][2], v[q][3]};
#pragma unroll
                    for (int i = 0; i < 4; i++) x[i] = st[(i * SPT + q) * BLOCK];
                    const bool on = combine4<MASKS>(v[q], nv[q], x, is_none<MASKS>(x), w, nv[q]);
                    shift_pow2(v[q], sc[q], thr_hi, on);
                }
            } break;
            case FOP_ROOT: {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q][i] = v[q][i] * pir[i];
                    shift_pow2(v[q], sc[q], thr_hi, shift_at_root != 0);
                    const double sum = ((v[q][0] + v[q][1]) + v[q][2]) + v[q][3];
                    // a column without any observation: +inf here, 0 after root_combine (lib/phylo_ctmc.ml:172, :302)
                    // (value, binary exponent): root_combine takes the one logarithm of the site, of the category mixture
                    site_cat[(int64_t)c * spad + s0 + q * BLOCK] = (MASKS && nv[q]) ? INFINITY : sum;
                    site_exp[(int64_t)c * spad + s0 + q * BLOCK] = sc[q].e;
                }
            } break;
            default: break;
            }
            if (opw.x & FOP_PUSH_AFTER) {
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    st[(0 * SPT + q) * BLOCK] = (MASKS && nv[q]) ? NONE_MARK : v[q][0];
#pragma unroll
                    for (int i = 1; i < 4; i++) st[(i * SPT + q) * BLOCK] = v[q][i];
                }
                st += 4 * BLOCK * SPT;
            }
        }
    }
}

template <int BLOCK, int SPT, bool MASKS>
__device__ __noinline__ void fused4t_general(FUSED4T_PARAMS)
{
    fused4t_body<BLOCK, SPT, MASKS, false>(FUSED4T_ARGS);
}

template <int BLOCK, int SPT, int MINB = 0, bool MASKS = false>
__global__ void __launch_bounds__(BLOCK, MINB ? MINB : (BLOCK == 256 ? (SPT <= 2 ? 3 : 2) : (BLOCK == 128 ? (SPT <= 2 ? 6 : 4) : 1)))
fused4t(FUSED4T_PARAMS, const int* __restrict__ fold_flag)
{
    if (fold_flag && *fold_flag == 0)       // (uniform over the grid)
        fused4t_body<BLOCK, SPT, MASKS, true>(FUSED4T_ARGS);
    else
        fused4t_general<BLOCK, SPT, MASKS>(FUSED4T_ARGS);
}

}  // namespace pbk
