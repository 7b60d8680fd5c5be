// Hand-written sm_100a kernels of the phylogenetic-likelihood hot path.
//
// Arithmetic is IEEE FP64 throughout.  Every kernel restates a piece of the
// reference (biocaml/phylogenetics, paths relative to its checkout):
//   K1 pmat_eigen        Site_evolution_model.Make_diag        lib/site_evolution_model.ml:43-47
//   K2 level4/level_dense/fused4   Phylo_ctmc.pruning recursion lib/phylo_ctmc.ml:83-98
//                        SV.mul + SV.shift rescaling           lib/phylo_ctmc.ml:33-40,76-78
//   K3 root_combine/final_sum      Felsenstein.multisite sum   lib/felsenstein.ml:73-75
//
// Data layout in HBM:
//   tips   uint8  [ntaxa][tips_ld]            site contiguous
//   P      f64    [ncat][nnodes][n*n]         row-major (i = parent state, j = child state)
//   CLV    f64    slot -> [ncat][n+1][spad]   plane-per-state, site contiguous; plane n = carry
//   site_cat f64  [ncat][spad]                log-likelihood of every (category, site)
#pragma once
#include <climits>
#include <cstdint>
#include <cuda_runtime.h>

namespace pbk {

// --------------------------------------------------------------------------
// Tree programs built on the host (pb_api.cu) and read, uniformly, by every thread.
// --------------------------------------------------------------------------
struct LevelChild {
    int kind;    // 0 = tip (src = row of the tip matrix), 1 = inner node (src = CLV slot)
    int src;
    int pnode;   // node id at the lower end of the branch = index of its P matrix
    int pad;
};
struct LevelNode {
    int nchild;
    int child_off;   // first entry in the LevelChild array
    int out_slot;    // CLV slot written (ignored for the root unless KEEP_CLV)
    int is_root;
};

enum FusedOpCode : int {
    FOP_TIP_TIP = 0,   // ACC = shiftmul(col(pA, tip rowA), col(pB, tip rowB))
    FOP_TIP = 1,       // ACC = col(pA, tip rowA), carry 0
    FOP_APPLY = 2,     // ACC.v = P[pA] . ACC.v            (branch term of a finished subtree)
    FOP_MUL_TIP = 3,   // ACC = shiftmul(ACC, col(pA, tip rowA))
    FOP_PUSH = 4,      // stack[sp++] = ACC
    FOP_POP_MUL = 5,   // ACC = shiftmul(stack[--sp], ACC)
    FOP_APPLY_MUL_TIP = 6,  // ACC = shiftmul(P[pA].ACC, col(pB, tip rowB))
    FOP_ROOT = 7       // x pi, optional shift, log(sum) + carry -> site_cat
};

// --------------------------------------------------------------------------
// K1: P[c][b] = U diag(exp(t_b r_c lambda)) Uinv for every branch x category, one launch.
// Follows Make_diag (lib/site_evolution_model.ml:43-47): d = exp(t*lambda) elementwise,
// then (U . diagm d) . Uinv with k accumulated in ascending order (dgemm).
// grid = (nnodes, ncat); U and Uinv row-major here (converted at the ABI).
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pmat_eigen(int n, int nnodes, int skip_node, const double* __restrict__ U,
           const double* __restrict__ lambda, const double* __restrict__ Uinv,
           const double* __restrict__ blen, const double* __restrict__ rates,
           double* __restrict__ P)
{
    extern __shared__ double sh[];
    double* e = sh;              // [n]
    double* ud = sh + n;         // [n*n]  U . diag(e), row-major
    double* ui = ud + n * n;     // [n*n]  Uinv, row-major
    const int b = blockIdx.x, c = blockIdx.y;
    if (b == skip_node) return;
    const double rate = rates ? rates[c] : 1.0;
    const double t = blen[b] * rate;
    for (int k = threadIdx.x; k < n; k += blockDim.x) e[k] = exp(t * lambda[k]);
    __syncthreads();
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        ud[idx] = U[idx] * e[idx % n];
        ui[idx] = Uinv[idx];
    }
    __syncthreads();
    double* out = P + ((size_t)c * nnodes + b) * n * n;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int i = idx / n, j = idx - i * n;
        double acc = 0.0;
        for (int k = 0; k < n; k++) acc = fma(ud[i * n + k], ui[k * n + j], acc);
        out[idx] = acc;
    }
}

// --------------------------------------------------------------------------
// SV.shift (lib/phylo_ctmc.ml:33-40): keep when min v > threshold, otherwise
// v := (1/max v) * v and carry += log(max v).
// --------------------------------------------------------------------------
__device__ __forceinline__ void shift4(double (&v)[4], double& carry, double thr)
{
    const double mn = fmin(fmin(v[0], v[1]), fmin(v[2], v[3]));
    if (!(mn > thr)) {
        const double mx = fmax(fmax(v[0], v[1]), fmax(v[2], v[3]));
        const double inv = 1.0 / mx;
        v[0] = inv * v[0]; v[1] = inv * v[1]; v[2] = inv * v[2]; v[3] = inv * v[3];
        carry += log(mx);
    }
}

template <int N>
__device__ __forceinline__ void shiftN(double (&v)[N], double& carry, double thr)
{
    double mn = v[0], mx = v[0];
#pragma unroll
    for (int i = 1; i < N; i++) { mn = fmin(mn, v[i]); mx = fmax(mx, v[i]); }
    if (!(mn > thr)) {
        const double inv = 1.0 / mx;
#pragma unroll
        for (int i = 0; i < N; i++) v[i] = inv * v[i];
        carry += log(mx);
    }
}

// y = P x for a row-major 4x4 P (dgemv: j accumulated in ascending order).
__device__ __forceinline__ void matvec4(const double* __restrict__ Pm, const double (&x)[4],
                                        double (&y)[4])
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double a = Pm[i * 4 + 0] * x[0];
        a = fma(Pm[i * 4 + 1], x[1], a);
        a = fma(Pm[i * 4 + 2], x[2], a);
        a = fma(Pm[i * 4 + 3], x[3], a);
        y[i] = a;
    }
}

// Root of Phylo_ctmc.pruning (lib/phylo_ctmc.ml:97-98) or of Felsenstein
// (lib/felsenstein.ml:47-48, no shift): log(sum(v * pi)) + carry.
__device__ __forceinline__ double root4(double (&v)[4], double carry, const double* __restrict__ pi,
                                        double thr, int shift_at_root)
{
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = v[i] * pi[i];
    if (shift_at_root) shift4(v, carry, thr);
    const double s = ((v[0] + v[1]) + v[2]) + v[3];
    return log(s) + carry;
}

// --------------------------------------------------------------------------
// K2 (level-scheduled, 4 states): one launch per tree level; blockIdx.y = node of the level,
// blockIdx.z = rate category; each thread owns TWO adjacent sites (16-byte vector loads and
// stores on every state plane).  Children are folded left to right exactly as
// List.reduce_exn ~f:SV.mul does (lib/phylo_ctmc.ml:91-95).  Binary nodes (the common case)
// issue the loads of both children before any arithmetic so that ~160 bytes per thread are
// in flight.  Algorithmic traffic per site.node: 8*C*5*(k_inner + 1) + k_tip bytes.
// --------------------------------------------------------------------------
struct Term4 {           // one child's contribution for the thread's two sites
    double a[4], b[4], ca, cb;
};

__device__ __forceinline__ void load_child4(const LevelChild cd, const double* __restrict__ Pm,
                                            const uint8_t* __restrict__ tips, int64_t tips_ld,
                                            const double* __restrict__ pool, int64_t slot_stride, int64_t spad,
                                            int c, int64_t s0, double2 (&raw)[5], uchar2& st)
{
    if (cd.kind == 0) {
        st = *reinterpret_cast<const uchar2*>(tips + (int64_t)cd.src * tips_ld + s0);
    } else {
        const double* src = pool + (int64_t)cd.src * slot_stride + (int64_t)c * 5 * spad + s0;
#pragma unroll
        for (int i = 0; i < 5; i++) raw[i] = __ldcs(reinterpret_cast<const double2*>(src + (int64_t)i * spad));
    }
}

__device__ __forceinline__ void term_child4(const LevelChild cd, const double* __restrict__ Pm,
                                            const double2 (&raw)[5], const uchar2 st, Term4& t)
{
    if (cd.kind == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) { t.a[i] = Pm[i * 4 + st.x]; t.b[i] = Pm[i * 4 + st.y]; }
        t.ca = 0.0; t.cb = 0.0;
    } else {
        double xa[4], xb[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { xa[i] = raw[i].x; xb[i] = raw[i].y; }
        matvec4(Pm, xa, t.a);
        matvec4(Pm, xb, t.b);
        t.ca = raw[4].x; t.cb = raw[4].y;
    }
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB)
level4(const LevelNode* __restrict__ nodes, const LevelChild* __restrict__ children, int node_base,
       const double* __restrict__ P, int nnodes,
       const uint8_t* __restrict__ tips, int64_t tips_ld,
       double* __restrict__ pool, int64_t slot_stride, int64_t spad,
       const double* __restrict__ pi, double thr, int shift_at_root,
       double* __restrict__ site_cat, int store_root)
{
    __shared__ double shP[2 * 16];    // the two matrices of a binary node (other arities read P from L1)
    const LevelNode nd = nodes[node_base + blockIdx.y];
    const LevelChild* ch = children + nd.child_off;
    const int c = blockIdx.z;
    const double* Pc = P + (size_t)c * nnodes * 16;
    if (nd.nchild == 2) {
        if (threadIdx.x < 32) shP[threadIdx.x] = Pc[(size_t)ch[threadIdx.x >> 4].pnode * 16 + (threadIdx.x & 15)];
        __syncthreads();
    }
    const int64_t s0 = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (s0 >= spad) return;

    double va[4], vb[4], ca, cb;
    if (nd.nchild == 2) {
        const LevelChild c0 = ch[0], c1 = ch[1];
        double2 r0[5], r1[5];
        uchar2 st0 = make_uchar2(0, 0), st1 = make_uchar2(0, 0);
        load_child4(c0, shP, tips, tips_ld, pool, slot_stride, spad, c, s0, r0, st0);
        load_child4(c1, shP + 16, tips, tips_ld, pool, slot_stride, spad, c, s0, r1, st1);
        Term4 t0, t1;
        term_child4(c0, shP, r0, st0, t0);
        term_child4(c1, shP + 16, r1, st1, t1);
#pragma unroll
        for (int i = 0; i < 4; i++) { va[i] = t0.a[i] * t1.a[i]; vb[i] = t0.b[i] * t1.b[i]; }
        ca = t0.ca + t1.ca; cb = t0.cb + t1.cb;
        shift4(va, ca, thr);
        shift4(vb, cb, thr);
    } else {
        ca = 0.0; cb = 0.0;
        for (int k = 0; k < nd.nchild; k++) {
            const LevelChild cd = ch[k];
            const double* Pm = Pc + (size_t)cd.pnode * 16;
            double2 r[5];
            uchar2 st = make_uchar2(0, 0);
            load_child4(cd, Pm, tips, tips_ld, pool, slot_stride, spad, c, s0, r, st);
            Term4 t;
            term_child4(cd, Pm, r, st, t);
            if (k == 0) {
#pragma unroll
                for (int i = 0; i < 4; i++) { va[i] = t.a[i]; vb[i] = t.b[i]; }
                ca = t.ca; cb = t.cb;
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) { va[i] = va[i] * t.a[i]; vb[i] = vb[i] * t.b[i]; }
                ca = ca + t.ca; cb = cb + t.cb;
                shift4(va, ca, thr);
                shift4(vb, cb, thr);
            }
        }
    }
    if (!nd.is_root || store_root) {
        double* dst = pool + (int64_t)nd.out_slot * slot_stride + (int64_t)c * 5 * spad + s0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            *reinterpret_cast<double2*>(dst + (int64_t)i * spad) = make_double2(va[i], vb[i]);
        *reinterpret_cast<double2*>(dst + 4 * spad) = make_double2(ca, cb);
    }
    if (nd.is_root) {
        const double la = root4(va, ca, pi, thr, shift_at_root);
        const double lb = root4(vb, cb, pi, thr, shift_at_root);
        *reinterpret_cast<double2*>(site_cat + (int64_t)c * spad + s0) = make_double2(la, lb);
    }
}

// --------------------------------------------------------------------------
// level4_async: level4 for binary nodes, software-pipelined with per-thread cp.async.
//
// A compute-free kernel with level4's traffic (10 planes read, 5 written, 16 B per thread per
// plane) reaches 6.7 TB/s on this B200 (tools/stream_probe.cu) while level4 sits at 5.6 TB/s inside
// its kernels: it holds 64 registers per thread, so too few loads are in flight while threads
// compute and store.  Here every thread copies the 160 bytes of its NEXT tile straight into its own
// shared-memory slots (LDGSTS: asynchronous, no registers, no barrier - a thread reads back only what
// it copied) and meanwhile works on the current tile: two tiles of traffic stay in flight per thread.
// grid = (chunks, nodes of the level, ncat); a CTA walks tiles_per_cta consecutive 512-site tiles.
// --------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}

__global__ void __launch_bounds__(256, 2)
level4_async(const LevelNode* __restrict__ nodes, const LevelChild* __restrict__ children, int node_base,
             const double* __restrict__ P, int nnodes,
             const uint8_t* __restrict__ tips, int64_t tips_ld,
             double* __restrict__ pool, int64_t slot_stride, int64_t spad,
             const double* __restrict__ pi, double thr, int shift_at_root,
             double* __restrict__ site_cat, int store_root, int tiles_per_cta)
{
    extern __shared__ double shbuf[];           // P[2][16] | stage[2][child 2][plane 5][256 threads] double2 | tips
    double* shP = shbuf;
    double2* shX = reinterpret_cast<double2*>(shbuf + 32);
    unsigned* shT = reinterpret_cast<unsigned*>(shX + 2 * 2 * 5 * 256);          // [stage 2][child 2][256]
    const LevelNode nd = nodes[node_base + blockIdx.y];
    const LevelChild* ch = children + nd.child_off;
    const int c = blockIdx.z, tid = threadIdx.x;
    const LevelChild c0 = ch[0], c1 = ch[1];
    if (tid < 32) shP[tid] = P[((size_t)c * nnodes + (tid < 16 ? c0.pnode : c1.pnode)) * 16 + (tid & 15)];
    __syncthreads();
    const int64_t ntiles = spad / 512;
    const int64_t t_begin = (int64_t)blockIdx.x * tiles_per_cta;
    const int64_t t_end = min(t_begin + tiles_per_cta, ntiles);

    auto issue = [&](int64_t tile, int stage) {
        const int64_t s0 = tile * 512 + 2 * tid;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const LevelChild cd = k ? c1 : c0;
            if (cd.kind == 0) {
                // 4 aligned tip bytes holding this thread's two sites
                cp_async_4(shT + (stage * 2 + k) * 256 + tid, tips + (int64_t)cd.src * tips_ld + (s0 & ~(int64_t)3));
            } else {
                const double* src = pool + (int64_t)cd.src * slot_stride + (int64_t)c * 5 * spad + s0;
                double2* dst = shX + ((stage * 2 + k) * 5) * 256 + tid;
#pragma unroll
                for (int i = 0; i < 5; i++) cp_async_16(dst + i * 256, src + (int64_t)i * spad);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (t_begin < t_end) issue(t_begin, 0);
    int stage = 0;
    for (int64_t tile = t_begin; tile < t_end; tile++, stage ^= 1) {
        if (tile + 1 < t_end) {
            issue(tile + 1, stage ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        const int64_t s0 = tile * 512 + 2 * tid;
        Term4 t[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const LevelChild cd = k ? c1 : c0;
            const double* Pm = shP + 16 * k;
            if (cd.kind == 0) {
                const unsigned w = shT[(stage * 2 + k) * 256 + tid];
                const unsigned sh = (unsigned)(s0 & 3) * 8;          // 0 or 16
                const int sa = (w >> sh) & 0xff, sb = (w >> (sh + 8)) & 0xff;
#pragma unroll
                for (int i = 0; i < 4; i++) { t[k].a[i] = Pm[i * 4 + sa]; t[k].b[i] = Pm[i * 4 + sb]; }
                t[k].ca = 0.0; t[k].cb = 0.0;
            } else {
                const double2* x = shX + ((stage * 2 + k) * 5) * 256 + tid;
                double xa[4], xb[4];
#pragma unroll
                for (int i = 0; i < 4; i++) { const double2 q = x[i * 256]; xa[i] = q.x; xb[i] = q.y; }
                const double2 qc = x[4 * 256];
                matvec4(Pm, xa, t[k].a);
                matvec4(Pm, xb, t[k].b);
                t[k].ca = qc.x; t[k].cb = qc.y;
            }
        }
        double va[4], vb[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { va[i] = t[0].a[i] * t[1].a[i]; vb[i] = t[0].b[i] * t[1].b[i]; }
        double ca = t[0].ca + t[1].ca, cb = t[0].cb + t[1].cb;
        shift4(va, ca, thr);
        shift4(vb, cb, thr);
        if (!nd.is_root || store_root) {
            double* dst = pool + (int64_t)nd.out_slot * slot_stride + (int64_t)c * 5 * spad + s0;
#pragma unroll
            for (int i = 0; i < 4; i++)
                *reinterpret_cast<double2*>(dst + (int64_t)i * spad) = make_double2(va[i], vb[i]);
            *reinterpret_cast<double2*>(dst + 4 * spad) = make_double2(ca, cb);
        }
        if (nd.is_root) {
            const double la = root4(va, ca, pi, thr, shift_at_root);
            const double lb = root4(vb, cb, pi, thr, shift_at_root);
            *reinterpret_cast<double2*>(site_cat + (int64_t)c * spad + s0) = make_double2(la, lb);
        }
    }
}

// --------------------------------------------------------------------------
// K2 (level-scheduled, N states, scalar DFMA version): one thread per site, the child's
// CLV tile and its P matrix staged in shared memory, running product in registers.
// Baseline and fallback for the tensor-core kernel (level_dmma, pb_kernels_dmma.cuh).
//
// MASKS = true is the Missing_values / Ambiguous form (lib/phylo_ctmc.ml:145-172, :280-302):
// a tip is a SET of states (bit i of a 64-bit mask; Ambiguous.leaf_indicator :234-242) and its
// term is P . indicator; an empty set is a missing observation (None): that child is skipped,
// no product and no shift, and a node all of whose children are None is None itself.  "None" is
// carried in HBM as a NaN in the carry plane, and as +inf in site_cat (-> log-likelihood 0).
// --------------------------------------------------------------------------
template <int N, bool MASKS>
__global__ void __launch_bounds__(128)
level_dense(const LevelNode* __restrict__ nodes, const LevelChild* __restrict__ children,
            int node_base, const double* __restrict__ P, int nnodes, int ncat,
            const uint8_t* __restrict__ tips, const uint64_t* __restrict__ masks, int64_t tips_ld,
            double* __restrict__ pool, int64_t slot_stride, int64_t spad,
            const double* __restrict__ pi, double thr, int shift_at_root,
            double* __restrict__ site_cat, int store_root)
{
    extern __shared__ double sh[];
    double* shP = sh;                 // [N*N]
    double* shX = sh + N * N;         // [N][128]
    const LevelNode nd = nodes[node_base + blockIdx.y];
    const LevelChild* ch = children + nd.child_off;
    const int tid = threadIdx.x;
    const int64_t s = (int64_t)blockIdx.x * 128 + tid;   // spad is a multiple of 128

    for (int c = 0; c < ncat; c++) {
        double acc[N];
        double carry = 0.0;
        bool have = false;            // MASKS: at least one child had an observation below it
        for (int k = 0; k < nd.nchild; k++) {
            const LevelChild cd = ch[k];
            __syncthreads();
            const double* Pg = P + ((size_t)c * nnodes + cd.pnode) * N * N;
            for (int idx = tid; idx < N * N; idx += 128) shP[idx] = Pg[idx];
            double tc = 0.0;
            int st = 0;
            uint64_t m = 0;
            bool ok = true;
            if (cd.kind == 0) {
                if (MASKS) { m = masks[(int64_t)cd.src * tips_ld + s]; ok = m != 0; }
                else st = tips[(int64_t)cd.src * tips_ld + s];
            } else {
                const double* src = pool + (int64_t)cd.src * slot_stride + (int64_t)c * (N + 1) * spad + s;
#pragma unroll 4
                for (int j = 0; j < N; j++) shX[j * 128 + tid] = src[(int64_t)j * spad];
                tc = src[(int64_t)N * spad];
                if (MASKS) ok = !isnan(tc);
            }
            __syncthreads();
            const bool first = MASKS ? !have : (k == 0);
            if (ok) carry = first ? tc : carry + tc;
#pragma unroll
            for (int i = 0; i < N; i++) {
                double t;
                if (cd.kind == 0) {
                    if (MASKS) {          // dgemv against the 0/1 indicator: sum of the set columns, ascending
                        t = 0.0;
                        for (int j = 0; j < N; j++)
                            if ((m >> j) & 1) t += shP[i * N + j];
                    } else {
                        t = shP[i * N + st];
                    }
                } else {
                    t = shP[i * N] * shX[tid];
#pragma unroll 4
                    for (int j = 1; j < N; j++) t = fma(shP[i * N + j], shX[j * 128 + tid], t);
                }
                if (ok) acc[i] = first ? t : acc[i] * t;
            }
            if (ok && !first) shiftN<N>(acc, carry, thr);
            if (ok) have = true;
        }
        const bool none = MASKS && !have;
        if (!nd.is_root || store_root) {
            double* dst = pool + (int64_t)nd.out_slot * slot_stride + (int64_t)c * (N + 1) * spad + s;
#pragma unroll
            for (int i = 0; i < N; i++) dst[(int64_t)i * spad] = none ? 0.0 : acc[i];
            dst[(int64_t)N * spad] = none ? NAN : carry;
        }
        if (nd.is_root) {
            double out;
            if (none) {
                out = INFINITY;        // all-missing site: root_combine turns it into 0. (:172, :302)
            } else {
#pragma unroll
                for (int i = 0; i < N; i++) acc[i] = acc[i] * pi[i];
                if (shift_at_root) shiftN<N>(acc, carry, thr);
                double sum = acc[0];
#pragma unroll
                for (int i = 1; i < N; i++) sum += acc[i];
                out = log(sum) + carry;
            }
            site_cat[(int64_t)c * spad + s] = out;
        }
    }
}

// The same kernel for ANY state count n <= NP (Phylo_ctmc.pruning ~nstates is generic, lib/phylo_ctmc.mli:55-61):
// NP in {8, 16, 32, 64} is the register capacity, loops are unrolled over NP and predicated on i < n, P is staged
// with row stride n exactly as above.  Used for the alphabets that have no specialised instantiation (e.g. 16 or 21
// states, a 60-codon code); the specialised sizes (4, 20, 61, ...) never come here.
template <int NP, bool MASKS>
__global__ void __launch_bounds__(128)
level_dense_rt(int n, const LevelNode* __restrict__ nodes, const LevelChild* __restrict__ children,
               int node_base, const double* __restrict__ P, int nnodes, int ncat,
               const uint8_t* __restrict__ tips, const uint64_t* __restrict__ masks, int64_t tips_ld,
               double* __restrict__ pool, int64_t slot_stride, int64_t spad,
               const double* __restrict__ pi, double thr, int shift_at_root,
               double* __restrict__ site_cat, int store_root)
{
    extern __shared__ double sh[];
    double* shP = sh;                 // [n*n]
    double* shX = sh + NP * NP;       // [n][128]
    const LevelNode nd = nodes[node_base + blockIdx.y];
    const LevelChild* ch = children + nd.child_off;
    const int tid = threadIdx.x;
    const int64_t s = (int64_t)blockIdx.x * 128 + tid;

    auto shift = [&](double (&v)[NP], double& carry) {       // SV.shift over the n live components
        double mn = v[0], mx = v[0];
#pragma unroll
        for (int i = 1; i < NP; i++) if (i < n) { mn = fmin(mn, v[i]); mx = fmax(mx, v[i]); }
        if (!(mn > thr)) {
            const double inv = 1.0 / mx;
#pragma unroll
            for (int i = 0; i < NP; i++) if (i < n) v[i] = inv * v[i];
            carry += log(mx);
        }
    };

    for (int c = 0; c < ncat; c++) {
        double acc[NP];
        double carry = 0.0;
        bool have = false;
        for (int k = 0; k < nd.nchild; k++) {
            const LevelChild cd = ch[k];
            __syncthreads();
            const double* Pg = P + ((size_t)c * nnodes + cd.pnode) * n * n;
            for (int idx = tid; idx < n * n; idx += 128) shP[idx] = Pg[idx];
            double tc = 0.0;
            int st = 0;
            uint64_t m = 0;
            bool ok = true;
            if (cd.kind == 0) {
                if (MASKS) { m = masks[(int64_t)cd.src * tips_ld + s]; ok = m != 0; }
                else st = tips[(int64_t)cd.src * tips_ld + s];
            } else {
                const double* src = pool + (int64_t)cd.src * slot_stride + (int64_t)c * (n + 1) * spad + s;
                for (int j = 0; j < n; j++) shX[j * 128 + tid] = src[(int64_t)j * spad];
                tc = src[(int64_t)n * spad];
                if (MASKS) ok = !isnan(tc);
            }
            __syncthreads();
            const bool first = MASKS ? !have : (k == 0);
            if (ok) carry = first ? tc : carry + tc;
#pragma unroll
            for (int i = 0; i < NP; i++) {
                if (i >= n) continue;
                double t;
                if (cd.kind == 0) {
                    if (MASKS) {
                        t = 0.0;
                        for (int j = 0; j < n; j++)
                            if ((m >> j) & 1) t += shP[i * n + j];
                    } else {
                        t = shP[i * n + st];
                    }
                } else {
                    t = shP[i * n] * shX[tid];
                    for (int j = 1; j < n; j++) t = fma(shP[i * n + j], shX[j * 128 + tid], t);
                }
                if (ok) acc[i] = first ? t : acc[i] * t;
            }
            if (ok && !first) shift(acc, carry);
            if (ok) have = true;
        }
        const bool none = MASKS && !have;
        if (!nd.is_root || store_root) {
            double* dst = pool + (int64_t)nd.out_slot * slot_stride + (int64_t)c * (n + 1) * spad + s;
#pragma unroll
            for (int i = 0; i < NP; i++) if (i < n) dst[(int64_t)i * spad] = none ? 0.0 : acc[i];
            dst[(int64_t)n * spad] = none ? NAN : carry;
        }
        if (nd.is_root) {
            double out;
            if (none) {
                out = INFINITY;
            } else {
#pragma unroll
                for (int i = 0; i < NP; i++) if (i < n) acc[i] = acc[i] * pi[i];
                if (shift_at_root) shift(acc, carry);
                double sum = acc[0];
#pragma unroll
                for (int i = 1; i < NP; i++) if (i < n) sum += acc[i];
                out = log(sum) + carry;
            }
            site_cat[(int64_t)c * spad + s] = out;
        }
    }
}

// --------------------------------------------------------------------------
// K3: per-site log-likelihood from the per-category values (log-sum-exp with the category
// weights; identity for one category) and a deterministic two-stage sum of all sites.
// Sites >= nsites (padding) contribute nothing.
// --------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256(double x, double* red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = x;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (l < (int)(blockDim.x >> 5)) ? red[l] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    return r;   // valid in thread 0
}

// (at most 40 registers: its blocks run beside the three resident blocks of the pruning kernel on a side stream)
__global__ void __launch_bounds__(256, 6)
root_combine(const double* __restrict__ site_cat,
             const int* __restrict__ site_exp,          // nullable: site_cat holds likelihoods as value x 2^site_exp (the
                                                        // site-resident 4-state kernels in power-of-two mode), else logarithms
             int ncat, const double* __restrict__ weights,
             const double* __restrict__ site_weights,   // nullable: multiplicity of every column (pattern compression)
             int64_t nsites, int64_t spad, int64_t block0, double* __restrict__ per_site,
             double* __restrict__ partials)
{
    __shared__ double red[8];
    const int64_t blk = block0 + blockIdx.x;        // a launch may cover a window of the site blocks
    const int64_t s = blk * blockDim.x + threadIdx.x;
    double out = 0.0;
    if (s < nsites) {
        if (site_cat[s] == INFINITY) {          // no observation at all in this column: 0. like the reference
            out = 0.0;
        } else if (site_exp) {
            // log(sum_c w_c v_c 2^e_c) with ONE logarithm: the terms are aligned to the largest exponent by exact
            // multiplications by powers of two (a term more than 2^1022 below the largest contributes nothing)
            constexpr double LN2 = 0.693147180559945309417232121458;
            // a likelihood is normalised first to mantissa in [1, 2) x 2^E — however the kernel split it into value
            // and exponent (tables with folded exponents, rescaling at other nodes) the result is the same to the bit
            auto norm = [&](int c, double& mant, int& E) {
                const double x = site_cat[(int64_t)c * spad + s];
                const int ex = ((__double2hiint(x) >> 20) & 0x7ff) - 1023;
                mant = x * __hiloint2double((1023 - max(ex, -1022)) << 20, 0);
                E = site_exp[(int64_t)c * spad + s] + max(ex, -1022);
                return x > 0.0;
            };
            if (ncat == 1) {
                double mant;
                int E;
                out = norm(0, mant, E) ? (log(mant) + (double)E * LN2) + log(weights[0]) : -INFINITY;
            } else {
                int emax = INT_MIN;
                for (int c = 0; c < ncat; c++) {
                    double mant;
                    int E;
                    if (norm(c, mant, E)) emax = max(emax, E);
                }
                if (emax == INT_MIN) {
                    out = -INFINITY;
                } else {
                    double acc = 0.0;
                    for (int c = 0; c < ncat; c++) {
                        double mant;
                        int E;
                        if (!norm(c, mant, E)) continue;
                        const int d = E - emax;                                            // <= 0
                        acc += weights[c] * (mant * (d >= -1022 ? __hiloint2double((1023 + d) << 20, 0) : 0.0));
                    }
                    out = log(acc) + (double)emax * LN2;
                }
            }
        } else if (ncat == 1) {
            // one category: log(w_0 L) (a single weight other than 1 is the same model as a 2-category mixture
            // whose other weight is 0; log(1.0) is exactly 0, so the reference's ncat = 1 case is untouched)
            out = site_cat[s] + log(weights[0]);
        } else {
            double m = site_cat[s];
            for (int c = 1; c < ncat; c++) m = fmax(m, site_cat[(int64_t)c * spad + s]);
            if (m == -INFINITY) {
                out = m;
            } else {
                double acc = 0.0;
                for (int c = 0; c < ncat; c++)
                    acc += weights[c] * exp(site_cat[(int64_t)c * spad + s] - m);
                out = m + log(acc);
            }
        }
        per_site[s] = out;
        if (site_weights) out = site_weights[s] * out;
    }
    const double bs = block_sum_256(out, red);
    if (threadIdx.x == 0) partials[blk] = bs;
}

__global__ void __launch_bounds__(256)
final_sum(const double* __restrict__ partials, int64_t count, double* __restrict__ total)
{
    __shared__ double red[8];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < count; i += 256) acc += partials[i];
    const double bs = block_sum_256(acc, red);
    if (threadIdx.x == 0) *total = bs;
}

// --------------------------------------------------------------------------
// f3: ancestral states drawn from their joint conditional distribution given the leaves, from the
// conditional likelihoods the level kernels left in HBM.  Restates Phylo_ctmc.conditional_simulation
// (lib/phylo_ctmc.ml:122-143) and its Missing_values / Ambiguous variants (:199-229, :331-360): top
// down, a node's state is drawn with weights prior_i * cl_i, where prior = root frequencies at the root
// and the parent's row of the branch matrix below it; a node without information under it (NaN
// carry) and a missing leaf are drawn from the prior alone, an ambiguous leaf from the prior
// restricted to its state set, an observed leaf keeps its state.
// One thread per site walks the nodes in pre-order (parents first); every plane read is coalesced
// across sites.  The uniform variates come from a counter-based generator keyed by (seed, site, node)
// — the distribution is the reference's, the random stream is not GSL's.
// --------------------------------------------------------------------------
struct SimNode {
    int node;      // tree node
    int parent;    // its parent (-1 for the root)
    int slot;      // CLV slot of an inner node, -1 for a leaf
    int tip_row;   // row of the tip matrix for a leaf
};

__device__ __forceinline__ double uniform01(uint64_t seed, uint64_t site, uint64_t node, uint64_t nnodes)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (site * nnodes + node + 1);      // splitmix64 of a unique counter
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;                                  // second round: seeds differing in few bits
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * 0x1p-53;
}

__global__ void __launch_bounds__(256)
conditional_simulation(const SimNode* __restrict__ nodes, int count, int nnodes, int n,
                       const double* __restrict__ P, const double* __restrict__ pi,
                       const double* __restrict__ pool, int64_t slot_stride, int64_t spad,
                       const uint8_t* __restrict__ tips, const uint64_t* __restrict__ masks, int64_t nsites,
                       uint64_t seed, uint8_t* __restrict__ states)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsites) return;
    const uint64_t all = n >= 64 ? ~0ull : ((1ull << n) - 1);
    for (int k = 0; k < count; k++) {
        const SimNode nd = nodes[k];
        const double* prior = nd.parent < 0 ? pi : P + ((size_t)nd.node * n + states[(int64_t)nd.parent * spad + s]) * n;
        const double* cl = nullptr;            // weights = prior_i * cl_i (cl == nullptr: prior alone)
        uint64_t allowed = all;
        int st = -1;
        if (nd.slot < 0) {
            if (!masks) {
                st = tips[(int64_t)nd.tip_row * spad + s];
            } else {
                const uint64_t m = masks[(int64_t)nd.tip_row * spad + s] & all;
                if (m && !(m & (m - 1))) st = __ffsll((long long)m) - 1;      // a single observed state
                else if (m) allowed = m;                                       // ambiguous: the prior restricted to the set
            }
        } else {
            const double* base = pool + (size_t)nd.slot * slot_stride + s;
            if (!isnan(base[(int64_t)n * spad])) cl = base;                    // NaN carry: nothing observed below this node
        }
        if (st < 0) {
            double total = 0.0;
            for (int i = 0; i < n; i++)
                if ((allowed >> i) & 1) total += prior[i] * (cl ? cl[(int64_t)i * spad] : 1.0);
            const double u = uniform01(seed, (uint64_t)s, (uint64_t)nd.node, (uint64_t)nnodes) * total;
            double cum = 0.0;
            for (int i = 0; i < n; i++) {
                if (!((allowed >> i) & 1)) continue;
                const double w = prior[i] * (cl ? cl[(int64_t)i * spad] : 1.0);
                if (w > 0.0) st = i;                                           // the last state with weight: guard against u*total == cum
                cum += w;
                if (cum > u) break;
            }
            if (st < 0) st = 0;
        }
        states[(int64_t)nd.node * spad + s] = (uint8_t)st;
    }
}

}  // namespace pbk
