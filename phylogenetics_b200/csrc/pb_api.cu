// C-ABI implementation of include/phylo_b200.h: context, tree compiler, launches.
// No torch types, no CPU fallback: every compute entry point needs a CUDA device.
#include "../../include/phylo_b200.h"

#include <algorithm>
#include <cstring>
#include <memory>
#include <cmath>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <map>
#include <vector>

#include <cstdio>
#include <cstdlib>
#include "pb_kernels.cuh"
#include "pb_kernels_fused.cuh"
#include "pb_kernels_dmma.cuh"
#include "pb_kernels_fdmma.cuh"
#include "pb_kernels_expm.cuh"
#include "pb_host_linalg.h"

using pbk::FusedOp2;
using pbk::LevelChild;
using pbk::LevelNode;

namespace {

thread_local std::string g_create_error;
thread_local int g_create_status = 0;

// The fused 4-state kernel stages matrices in the device's constant bank, which every context on
// that device shares: users are serialised by a host mutex around the enqueue and, on the device,
// by an event recorded after the last kernel that read the bank.
std::recursive_mutex g_const_mutex;      // (recursive: the pipelined host evaluation holds it around its launches)
cudaEvent_t g_const_event[64] = {};

enum PSource { PSRC_NONE = 0, PSRC_EIGEN = 1, PSRC_DIRECT = 2, PSRC_EXPM = 3 };
enum Path { PATH_FUSED4 = 0, PATH_LEVEL4 = 1, PATH_LEVEL_DENSE = 2, PATH_LEVEL_DMMA = 3, PATH_FUSED_DMMA = 4 };

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t count = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= count && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        count = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) count = n;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        count = 0;
    }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
};

}  // namespace

struct pb_ctx {
    int device = 0, n = 0, ncat = 1, ntaxa = 0;
    int64_t nsites = 0, spad = 0;
    unsigned flags = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    double threshold = 1e-6;
    int shift_at_root = 1;
    mutable std::string err;
    mutable int err_status = 0;
    int64_t launches = 0;
    int sm_count = 148;
    size_t smem_optin = 0;

    // tree
    int nnodes = 0, root = -1;
    std::vector<int> child_ptr, child_idx, tip_row, parent;
    std::vector<double> branch_len;
    bool have_tree = false, have_blen = false;

    // compiled schedules
    std::vector<LevelNode> lv_nodes;
    std::vector<LevelChild> lv_children;
    std::vector<int> lv_first;          // first LevelNode of every level (+ end)
    std::vector<int> slot_of_node;      // CLV slot of every internal node (level path)
    int nslots = 0, max_arity = 0;
    std::vector<int> lv_node_id;        // tree node of every LevelNode
    std::vector<char> dirty;            // nodes whose CLV must be recomputed (partial re-evaluation)
    bool all_dirty = true, partial_ok = false;
    int last_nodes_recomputed = 0;
    DevBuf<LevelNode> d_sel_nodes;
    DevBuf<pbk::SimNode> d_sim_nodes;   // pb_conditional_simulation: nodes in pre-order
    DevBuf<uint8_t> d_states;           //   and the sampled states [nnodes][spad]
    // sequential (depth-first, deeper subtree first) schedule for the persistent dense kernel
    std::vector<LevelNode> sq_nodes;
    std::vector<LevelChild> sq_children;
    int sq_nslots = 0;
    DevBuf<LevelNode> d_sq_nodes;
    DevBuf<LevelChild> d_sq_children;
    int dmma_per_level = 1;             // option "dmma_per_level": one launch per tree level instead
    int dmma_version = 2;               // option "dmma_version": 2 = cp.async-pipelined kernel (arity <= 2), 1 = direct loads
    int dmma_tip_levels_v1 = 0;         // option: levels made of tip-only nodes use the unstaged kernel (2 CTAs/SM)
    int dmma_seq_tiles = 16;            // option "dmma_group_tiles": 8-site tiles per group (x8 sites)
    std::vector<FusedOp2> fprog;
    std::vector<int> tip_order;         // tip rows in the order the fused program consumes them
    int fstack = 0, fwords = 0;
    int fused_block = 256, fused_spt = 4, fused_exact_log = 0;
    bool fused_spt_set = false;            // "fused_spt" given explicitly; else fused4t runs 2 sites per thread (72 registers,
                                           // 3 blocks per SM: measured faster than 4 sites per thread, profiles/r02c_*)
    std::vector<FusedOp2> fprog_tab;       // fprog_c with the small clades turned into table look-ups (FOP_CHERRY)
    std::vector<pbk::CladeTab> clades;     // the tabulated clades of fprog_tab
    int fused_preapply = 0;                // option "fused_preapply" (only in a build with -DFUSED_PREAPPLY=1)
    bool tab_preapply = false;             // what fprog_tab was built with
    bool plan_only = false;                // a context without a device, used by pb_debug_fused_program only
    bool tables_stale = true;              // the tables must be rebuilt before the next launch (new program variant)
    int tab_variant = 0;                   // whose tables d_cherry_* hold: 1 = fused4c (look-ups in the full program), 2 = fused4t
    int tab_leaves = -1;                   // the clade size limit fprog_tab was built for (-1: not built)
    int ntab = 0;                          // table entries per category
    int fprog_bank_nops = 0;               // length of the program variant launch_fused4c put in the constant bank
    int fused_triple = 1;                  // option "fused_triple": 0 = tabulate cherries only
    int fused_clade_force = 0;             // option "fused_clade_force" (test hook): tabulate up to the limit however few the sites
    int fused_clade_leaves = 7;            // option "fused_clade_leaves": largest tabulated clade (2..8 leaves; 7: profiles/r02c_*)
    // fused4t: the program over the reduced tree (build_reduced_program)
    int fused_reduced = 1;                 // option "fused_reduced": 1 (default) = fused4t, 0 = fused4c with look-ups in the full program
    std::vector<FusedOp2> rprog, cprog;    // reduced program; the clades' own (untabulated, node-addressed) instruction ranges
    std::vector<int> rorder;               // packed-tip slot -> tip row (-1: padding), 32 slots per word, in reduced-program order
    std::vector<int> rxorder;              // ... and the row whose state is XORed into the slot (-1: none): difference coding
    int fused_xor = 1;                     // option "fused_xor": difference-code the tips of a tabulated clade
    int fused_masks = 1;                   // option "fused_masks": tip state sets run the reduced program (0: the level kernel)
    int fused_occ = 0;                     // option "fused_occ": 4 = fused4t<256, 2> compiled for 4 blocks per SM (64 registers)
    std::vector<int> rmat_nodes, rleaf_nodes;   // branches whose matrices the reduced program applies / looks columns up in
    std::vector<pbk::CladeTab> rclades;
    int rwords = 0, rstack = 0, rntab = 0, r_leaves = -1;
    // tip state sets in the reduced program (4 states): the alignment's distinct sets are its alphabet (pb_set_tip_masks)
    int mask_radix = 0;                    // how many distinct state sets occur (0: not scanned)
    unsigned char mask_code[16] = {}, mask_set[16] = {};    // set -> code, code -> set
    DevBuf<unsigned char> d_mask_code, d_mask_set;
    DevBuf<unsigned> d_mask_present;
    std::vector<pbk::PackField> rfields;   // the packed fields of the masked plan and the tip rows they encode
    std::vector<int> rfield_rows;
    DevBuf<pbk::PackField> d_rfields;
    DevBuf<int> d_rfield_rows;
    int r_radix = 0;                       // alphabet the plan was made for (4 = plain states)
    bool r_masked = false;
    int r_tipbits = 2;
    bool packed_for_reduced = false;       // which slot order d_packed currently holds
    bool reduced_fits = false;             // the last plan fits the constant bank and shared memory (want_reduced)
    DevBuf<FusedOp2> d_rprog, d_cprog;
    DevBuf<int> d_rorder, d_rxorder, d_rmat_nodes, d_rleaf_nodes;
    DevBuf<pbk::CladeTab> d_rclades;
    DevBuf<FusedOp2> d_fprog_tab;
    DevBuf<pbk::CladeTab> d_clades;
    DevBuf<int> d_cherry_k;
    DevBuf<double> d_cherry_v;
    int fused_cherry = 1;                  // option "fused_cherry": tabulate leaf-leaf nodes (power-of-two mode of fused4c)
    uint64_t fprog_version = 0;         // identifies the compiled program in the constant bank (launch_fused4c)
    int fused_pow2 = 1;                 // option "fused_pow2": power-of-two scale factors in the site-resident 4-state kernels
                                        // (exact and branch-free; 0 = the reference's factor 1/max)
    DevBuf<LevelNode> d_lv_nodes;
    DevBuf<LevelChild> d_lv_children;
    DevBuf<FusedOp2> d_fprog;
    DevBuf<int> d_tip_order;
    DevBuf<uint8_t> d_is_leaf;
    // fused4c: inner-branch matrices through constant memory
    std::vector<FusedOp2> fprog_c;
    std::vector<int> inner_nodes, leaf_nodes;
    DevBuf<FusedOp2> d_fprog_c;
    DevBuf<int> d_inner_nodes, d_leaf_nodes;
    DevBuf<double> d_Pc;
    int fused_const_cats = 1;           // option "fused_const_cats": categories per launch (fused4c: 1 unless set; fused4t: as many
    bool fused_const_cats_set = false;  // as the constant bank holds unless set — all four at c2: profiles/r02i_ab_e2e_options.jsonl)
    int fused_const = 1;                // option "fused_const": use fused4c when the tree has <= 128 inner branches
    // fused_dmma: the same program for 20 / 61 states (compact matrix indices, tip rows, matrix order)
    std::vector<FusedOp2> fprog_d;
    std::vector<int> mlist;
    int dstack = 0;
    int fused_dmma = 1;                 // option "fused_dmma": site-resident tensor-core kernel for 20 / 61 states
    int fused_dmma_version = 3;         // option "fused_dmma_version": 3 = decoupled warps + bulk-copy ring on m8n8k4 tiles,
                                        // 2 = the same on m16n8k8 tiles, 1 = CTA-synchronous
    int fused_dmma_stagger = 0;         // option "fused_dmma_stagger": start delay (ns) of every other warp quad
    int fused_dmma_slots = 0;           // option "fused_dmma_slots": cap on the ring slots (0 = as many as fit)
    int fused_dmma_spill = 0;           // option "fused_dmma_spill": test hook, all stack levels in the L2 scratch
    DevBuf<FusedOp2> d_fprog_d;
    DevBuf<int> d_mlist;
    // fused_dmma3 over the reduced tree: nodes with two leaf children are tables (build_dense_reduced_program)
    int fused_dmma_cherry = 1;          // option "fused_dmma_cherry"
    std::vector<FusedOp2> dprog_r;
    std::vector<int> mlist_r;
    std::vector<int4> dtabs;               // per table: {first entry, tip row A, tip row B, tip row C or -1}
    std::vector<pbk::DenseTabNodes> dtab_nodes;
    int64_t dtab_entries = 0;              // entries of all tables of one category
    bool dtab_three = false;               // some table is a three-leaf clade
    int dense_builder_waves = 0;           // option "fused_dmma_builder_waves": waves of blocks of the cherry-table builder (0: default)
    int fused_dmma_triple = 0;             // option "fused_dmma_triple": clades of three leaves (cherry + leaf) as tables, 20 states
                                           // (off by default: with two CTAs per SM cherries alone are faster, 0.560 against 0.568 ms
                                           // at c3, profiles/r03c_sweep_dense_2ctas.jsonl)
    int dstack_r = 0;
    bool dense_reduced_built = false;
    DevBuf<FusedOp2> d_dprog_r;
    DevBuf<int> d_mlist_r;
    DevBuf<int4> d_dtabs;
    DevBuf<pbk::DenseTabNodes> d_dtab_nodes;
    DevBuf<double> d_dtabV;
    DevBuf<int> d_dtabK;
    DevBuf<double> d_Ppad, d_PTleaf, d_gstack;
    DevBuf<uint64_t> d_packed;
    DevBuf<int> d_bad;
    bool pack_dirty = true, tips_checked = false;
    bool level4_async_ready = false;
    int level4_async = 1;               // option "level4_async": cp.async-pipelined level kernel for binary trees
    int level4_waves = 16;              // option "level4_waves": CTAs per SM slot per launch (tile granularity)
    int min_arity = 0;
    int level4_minb = 4;                // option "level4_min_blocks": occupancy target of the level4 kernel
    bool force_dense = false;           // option "dense_scalar": use the DFMA level kernel instead of DMMA

    // inputs
    DevBuf<uint8_t> d_tips;
    DevBuf<uint8_t> d_tips2;            // 2-bit packed upload buffer [ntaxa][spad/4] (pb_*_packed2)
    DevBuf<uint64_t> d_masks;
    bool have_tips = false, have_masks = false;
    DevBuf<double> d_pi, d_rates, d_weights, d_blen;
    bool have_pi = false;
    std::vector<double> h_rates, h_weights;
    DevBuf<double> d_U, d_lambda, d_Uinv, d_Q;
    DevBuf<double> d_P;
    int psrc = PSRC_NONE;
    bool p_dirty = true;

    // work + outputs
    DevBuf<double> d_pool, d_site_cat, d_per_site, d_partials, d_total, d_site_weights;
    DevBuf<int> d_site_exp;             // binary exponents beside d_site_cat (site_cat_raw)
    DevBuf<unsigned char> d_text, d_lut;
    bool have_site_weights = false;
    int64_t slot_stride = 0;
    double* h_total = nullptr;   // pinned: [0] = total, [1] = bad-tip flag (as int)
    bool evaluated = false;      // an evaluation has run since pb_set_tree (partial re-evaluation needs resident CLVs)
    bool fresh = false;          // ... and no input has changed since: what pb_get_clv / pb_conditional_simulation /
                                 // pb_result_device_ptrs hand out belongs to the current inputs

    // pipelined host evaluation (pb_loglik_host)
    cudaStream_t copy_stream = nullptr;
    cudaStream_t lane_streams[4] = {};  // the lanes of the pipelined host evaluation (option "host_lanes" of them)
    cudaEvent_t lane_events[4] = {};
    int host_lanes = 2;
    cudaEvent_t ev_root_done = nullptr;
    bool bank_held = false;             // inside the lanes of a pipelined host evaluation (launch_fused4t)
    int host_side_streams = 1;          // option "host_side_streams": 0 = pack, prune and reduce in one stream
    bool tip_bytes_stale = false;       // the alignment is resident 2-bit packed (d_tips2); d_tips is expanded on demand
    std::vector<cudaEvent_t> chunk_events;
    cudaEvent_t ev_start = nullptr;
    uint64_t pc_epoch = 1;              // compactions of the inner matrices (d_Pc) so far: what the constant bank was loaded from
    std::map<const void*, size_t> smem_allowed;                // dynamic shared memory the kernels launched so far may use
    std::map<std::pair<const void*, size_t>, int> occ_cache;   // resident blocks per SM of the kernels launched so far
    int fused_fold = 1;                 // option "fused_fold": table exponents folded into the vectors (one gather per look-up)
    int fused_const_keep = 1;           // option "fused_const_keep": do not reload a constant bank that already holds these matrices
    int fused_split = 1;                // option "fused_split" (measurement hook): the resident evaluation in this many launches
    int fused_const_group = 1;          // option "fused_const_group": load the constant bank for several categories at once
    int host_lead_cats = 64;            // option "host_lead_cats": rate categories pruned chunk by chunk behind a 2-bit packed copy
                                        // (all of them, in one launch per chunk: profiles/r02i_ab_e2e_options.jsonl)
    int host_chunk_align = 1;           // option "host_chunk_align": chunk bounds where a round of the persistent blocks ends
    int fused4t_occ = 3;                // resident blocks per SM of the fused4t shape last launched
    int64_t host_tail_sites = 0;        // option "host_tail_sites": the schedule ends on a chunk of about this many sites (0: one round
                                        // of the persistent blocks, -1: no short last chunk)
    int host_trace = 0;                 // option "host_trace": timing events per chunk (pb_debug_host_trace)
    std::vector<cudaEvent_t> trace_events;      // [0] = start, then (copy done, compute done) per chunk, then the end
    std::vector<int64_t> trace_sites;           // first site and width of every chunk
    int host_chunks = 6;                // option "host_chunks" (6-8 is best at 1M sites x 128 taxa: profiles/r02u_e2e_trace.jsonl)

    // option "graph": one evaluation of the reduced 4-state program captured as a CUDA graph and replayed while nothing
    // that shapes the launches changes (parameter VALUES live in device buffers whose addresses are stable)
    int use_graph = 0;
    cudaGraphExec_t graph_exec = nullptr;
    uint64_t config_epoch = 1, graph_epoch = 0, warm_epoch = 0;   // bumped by every setter that changes launches or kernel arguments
    bool capturing = false;
    int64_t graph_launches = 0;         // kernels inside the captured evaluation
    int64_t graph_replays = 0;

    // optional per-phase device timing (pb_set_timing)
    bool timing = false;
    static constexpr int TIMING_RING = 64;        // evaluations whose phase events are kept (pb_get_timing_at)
    cudaEvent_t ev_ring[TIMING_RING][4] = {};
    cudaEvent_t* ev = ev_ring[0];                 // event set of the evaluation being enqueued
    int64_t timed_evals = 0;
    // ... and of the dominant kernel alone (fused4t / fused_dmma3): one event pair per launch of the last timed evaluation
    std::vector<cudaEvent_t> kev;
    size_t kev_n = 0;
    bool kernel_timing = false;                   // pb_set_timing(ctx, 2)
};

namespace {

int fail(pb_ctx* c, int code, const std::string& msg)
{
    if (c) { c->err = msg; c->err_status = code; } else { g_create_error = msg; g_create_status = code; }
    return code;
}
int cuda_fail(pb_ctx* c, cudaError_t e, const char* what)
{
    return fail(c, PB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define PB_CUDA(c, call)                                         \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return cuda_fail(c, e__, #call); \
    } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// Column-major (ABI) -> row-major (device) for a batch of n x n matrices.
void to_row_major(int n, size_t count, const double* in, double* out)
{
    for (size_t m = 0; m < count; m++)
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) out[m * n * n + i * n + j] = in[m * n * n + j * n + i];
}

size_t fused_smem_bytes_for(const pb_ctx* c, int block, int spt)
{
    return (size_t)c->nnodes * 16 * 8 + (size_t)std::max(c->fstack, 1) * 4 * block * spt * 8 +
           c->fprog.size() * sizeof(FusedOp2);
}

// The all-shared-memory fused kernel keeps every branch matrix of the category in smem; large trees
// shrink the tile (and with it the stack area) before giving up on the fused path.
bool fused_shape(const pb_ctx* c, int* block, int* spt)
{
    const int cand[4][2] = {{c->fused_block, c->fused_spt}, {256, 2}, {256, 1}, {128, 1}};
    for (const auto& k : cand)
        if (fused_smem_bytes_for(c, k[0], k[1]) <= c->smem_optin) { *block = k[0]; *spt = k[1]; return true; }
    return false;
}

size_t fused_smem_bytes(const pb_ctx* c)
{
    int b = c->fused_block, q = c->fused_spt;
    fused_shape(c, &b, &q);
    return fused_smem_bytes_for(c, b, q);
}

int path_of(const pb_ctx* c)
{
    const bool level = (c->flags & (PB_FLAG_KEEP_CLV | PB_FLAG_LEVEL_KERNELS)) != 0;
    // tip state sets: the reduced 4-state program when it fits, else the scalar level kernel
    if (c->have_masks) return (c->n == 4 && !level && c->have_tree && c->reduced_fits) ? PATH_FUSED4 : PATH_LEVEL_DENSE;
    if (c->n == 4) {
        if (!level && c->have_tree) {
            // the reduced program (fused4t) keeps only the loose tips' matrices in shared memory; the all-shared-memory
            // kernel (fused4) needs every branch's
            if (c->reduced_fits || fused_smem_bytes(c) <= c->smem_optin) return PATH_FUSED4;
        }
        return PATH_LEVEL4;
    }
    if (!c->have_masks && pbk::dmma_supported(c->n) && !c->force_dense) {
        if (!level && c->have_tree && c->fused_dmma) {
            int ssm = 0, nslot = 0;
            size_t smem = 0;
            if ((c->fused_dmma_version == 3 &&
                 pbk::fused_dmma3_plan(c->n, c->smem_optin, c->dstack, c->ntaxa, (int)c->fprog_d.size(), (int)c->mlist.size(), &nslot, &ssm, &smem)) ||
                (c->fused_dmma_version == 2 &&
                 pbk::fused_dmma2_plan(c->n, c->smem_optin, c->dstack, c->ntaxa, (int)c->fprog_d.size(), (int)c->mlist.size(), &nslot, &ssm, &smem)) ||
                pbk::fused_dmma_plan(c->n, c->smem_optin, c->dstack, c->ntaxa, (int)c->fprog_d.size(), (int)c->mlist.size(), &ssm, &smem))
                return PATH_FUSED_DMMA;
        }
        const int mp = (c->n + 15) / 16 * 16, ld = (c->n + 7) / 8 * 8 + 4;
        if (((size_t)std::max(c->max_arity, 1) * mp * ld + mp) * sizeof(double) <= c->smem_optin) return PATH_LEVEL_DMMA;
    }
    return PATH_LEVEL_DENSE;
}

// ---- tree compiler ------------------------------------------------------
int compile_tree(pb_ctx* c)
{
    const int nn = c->nnodes;
    // post-order by explicit stack (trees can be caterpillars: no recursion)
    std::vector<int> order;
    order.reserve(nn);
    {
        std::vector<std::pair<int, int>> st;
        st.push_back({c->root, c->child_ptr[c->root]});
        while (!st.empty()) {
            auto& top = st.back();
            if (top.second < c->child_ptr[top.first + 1]) {
                const int ch = c->child_idx[top.second++];
                st.push_back({ch, c->child_ptr[ch]});
            } else {
                order.push_back(top.first);
                st.pop_back();
            }
        }
    }
    if ((int)order.size() != nn) return fail(c, PB_ERR_INVALID_ARG, "tree: not every node is reachable from the root");

    std::vector<int> level(nn, 0), su(nn, 0);
    c->max_arity = 0;
    c->min_arity = 1 << 30;
    for (int v : order) {
        const int c0 = c->child_ptr[v], c1 = c->child_ptr[v + 1];
        if (c0 == c1) continue;
        c->max_arity = std::max(c->max_arity, c1 - c0);
        c->min_arity = std::min(c->min_arity, c1 - c0);
        int lv = 0;
        for (int k = c0; k < c1; k++) lv = std::max(lv, level[c->child_idx[k]]);
        level[v] = lv + 1;
        // stack need of the accumulator machine (Sethi-Ullman style)
        if (c1 - c0 == 2) {
            const int a = c->child_idx[c0], b = c->child_idx[c0 + 1];
            const bool ia = c->child_ptr[a] != c->child_ptr[a + 1], ib = c->child_ptr[b] != c->child_ptr[b + 1];
            if (ia && ib) su[v] = std::max(std::max(su[a], su[b]), std::min(su[a], su[b]) + 1);
            else su[v] = ia ? su[a] : (ib ? su[b] : 0);
        } else {
            int need = 0;
            for (int k = c0; k < c1; k++) {
                const int ch = c->child_idx[k];
                const bool inner = c->child_ptr[ch] != c->child_ptr[ch + 1];
                if (inner) need = std::max(need, su[ch] + (k > c0 ? 1 : 0));
            }
            su[v] = need;
        }
    }

    // ---- level schedule + CLV slots
    std::vector<int> internals;
    for (int v : order) if (c->child_ptr[v] != c->child_ptr[v + 1]) internals.push_back(v);
    std::stable_sort(internals.begin(), internals.end(), [&](int a, int b) { return level[a] < level[b]; });
    c->lv_nodes.clear();
    c->lv_node_id.clear();
    c->dirty.assign(nn, 0);
    c->all_dirty = true;
    c->lv_children.clear();
    c->lv_first.clear();
    c->slot_of_node.assign(nn, -1);
    const bool keep = (c->flags & PB_FLAG_KEEP_CLV) != 0;
    std::vector<int> free_slots;
    int nslots = 0;
    size_t i = 0;
    while (i < internals.size()) {
        const int lv = level[internals[i]];
        size_t j = i;
        while (j < internals.size() && level[internals[j]] == lv) j++;
        c->lv_first.push_back((int)c->lv_nodes.size());
        for (size_t q = i; q < j; q++) {
            const int v = internals[q];
            int slot;
            if (!keep && !free_slots.empty()) { slot = free_slots.back(); free_slots.pop_back(); }
            else slot = nslots++;
            c->slot_of_node[v] = slot;
        }
        for (size_t q = i; q < j; q++) {
            const int v = internals[q];
            LevelNode nd;
            nd.nchild = c->child_ptr[v + 1] - c->child_ptr[v];
            nd.child_off = (int)c->lv_children.size();
            nd.out_slot = c->slot_of_node[v];
            nd.is_root = (v == c->root) ? 1 : 0;
            for (int k = c->child_ptr[v]; k < c->child_ptr[v + 1]; k++) {
                const int ch = c->child_idx[k];
                LevelChild cd;
                cd.pad = 0;
                cd.pnode = ch;
                if (c->child_ptr[ch] == c->child_ptr[ch + 1]) { cd.kind = 0; cd.src = c->tip_row[ch]; }
                else { cd.kind = 1; cd.src = c->slot_of_node[ch]; }
                c->lv_children.push_back(cd);
            }
            c->lv_nodes.push_back(nd);
            c->lv_node_id.push_back(v);
        }
        if (!keep)
            for (size_t q = i; q < j; q++) {
                const int v = internals[q];
                for (int k = c->child_ptr[v]; k < c->child_ptr[v + 1]; k++) {
                    const int ch = c->child_idx[k];
                    if (c->slot_of_node[ch] >= 0) free_slots.push_back(c->slot_of_node[ch]);
                }
            }
        i = j;
    }
    c->lv_first.push_back((int)c->lv_nodes.size());
    c->nslots = nslots;

    // ---- sequential schedule: depth-first post-order, the subtree needing more live CLVs first
    {
        c->sq_nodes.clear();
        c->sq_children.clear();
        std::vector<int> sq_slot(nn, -1), free_sq;
        int nsq = 0;
        std::vector<std::pair<int, int>> st;          // (node, next child rank)
        std::vector<std::vector<int>> kids(nn);
        for (int v = 0; v < nn; v++) {
            for (int k = c->child_ptr[v]; k < c->child_ptr[v + 1]; k++) kids[v].push_back(c->child_idx[k]);
            std::stable_sort(kids[v].begin(), kids[v].end(), [&](int a, int b) { return su[a] > su[b]; });
        }
        st.push_back({c->root, 0});
        while (!st.empty()) {
            auto& top = st.back();
            const int v = top.first;
            if (top.second < (int)kids[v].size()) {
                const int ch = kids[v][top.second++];
                if (c->child_ptr[ch] != c->child_ptr[ch + 1]) st.push_back({ch, 0});
                continue;
            }
            st.pop_back();
            int slot;
            if (keep) slot = c->slot_of_node[v];
            else if (!free_sq.empty()) { slot = free_sq.back(); free_sq.pop_back(); }
            else slot = nsq++;
            // the output slot must differ from the children's: allocate before freeing theirs
            sq_slot[v] = slot;
            LevelNode nd;
            nd.nchild = c->child_ptr[v + 1] - c->child_ptr[v];
            nd.child_off = (int)c->sq_children.size();
            nd.out_slot = slot;
            nd.is_root = (v == c->root) ? 1 : 0;
            for (int k = c->child_ptr[v]; k < c->child_ptr[v + 1]; k++) {     // reference branch order
                const int ch = c->child_idx[k];
                LevelChild cd;
                cd.pad = 0;
                cd.pnode = ch;
                if (c->child_ptr[ch] == c->child_ptr[ch + 1]) { cd.kind = 0; cd.src = c->tip_row[ch]; }
                else { cd.kind = 1; cd.src = sq_slot[ch]; if (!keep) free_sq.push_back(sq_slot[ch]); }
                c->sq_children.push_back(cd);
            }
            c->sq_nodes.push_back(nd);
        }
        c->sq_nslots = keep ? c->nslots : nsq;
    }

    // ---- fused accumulator/stack program (4-state path)
    c->fprog.clear();
    c->tip_order.clear();
    c->fstack = su[c->root];
    {
        auto is_inner = [&](int v) { return c->child_ptr[v] != c->child_ptr[v + 1]; };
        // abstract ops first: code, pA (+tip A), pB (+tip B); tips get their packed position below
        struct AOp { int code, pA, tipA, pB, tipB; };
        std::vector<AOp> aops;
        struct Item { int node; AOp emit; bool is_emit; };
        std::vector<Item> st;
        st.push_back({c->root, {}, false});
        while (!st.empty()) {
            Item it = st.back();
            st.pop_back();
            if (it.is_emit) { aops.push_back(it.emit); continue; }
            const int v = it.node;
            const int c0 = c->child_ptr[v], c1 = c->child_ptr[v + 1];
            std::vector<Item> seq;   // in execution order
            if (c1 - c0 == 2) {
                int a = c->child_idx[c0], b = c->child_idx[c0 + 1];
                const bool ia = is_inner(a), ib = is_inner(b);
                if (!ia && !ib) {
                    seq.push_back({0, {pbk::FOP_TIP_TIP, a, a, b, b}, true});
                } else if (ia && ib) {
                    if (su[b] > su[a]) std::swap(a, b);   // deeper first: exact, x*y and c1+c2 commute
                    seq.push_back({a, {}, false});
                    seq.push_back({0, {pbk::FOP_PUSH, 0, -1, 0, -1}, true});     // raw CLV of a; P[a] is applied at the pop
                    seq.push_back({b, {}, false});
                    seq.push_back({0, {pbk::FOP_ACC_POP, b, -1, a, -1}, true});
                } else {
                    const int in = ia ? a : b, tp = ia ? b : a;
                    seq.push_back({in, {}, false});
                    seq.push_back({0, {pbk::FOP_APPLY_MUL_TIP, in, -1, tp, tp}, true});
                }
            } else {
                for (int k = c0; k < c1; k++) {      // reference order (left fold)
                    const int ch = c->child_idx[k];
                    if (!is_inner(ch)) {
                        seq.push_back({0, {k == c0 ? pbk::FOP_TIP : pbk::FOP_MUL_TIP, ch, ch, 0, -1}, true});
                    } else {
                        if (k > c0) seq.push_back({0, {pbk::FOP_PUSH, 0, -1, 0, -1}, true});
                        seq.push_back({ch, {}, false});
                        seq.push_back({0, {pbk::FOP_APPLY, ch, -1, 0, -1}, true});
                        if (k > c0) seq.push_back({0, {pbk::FOP_POP_MUL, 0, -1, 0, -1}, true});
                    }
                }
            }
            for (auto r = seq.rbegin(); r != seq.rend(); ++r) st.push_back(*r);
        }
        aops.push_back({pbk::FOP_ROOT, 0, -1, 0, -1});
        // lower to FusedOp2: tips are consumed in program order, 32 per packed word
        auto mk = [](int code, int pA, int shA, int pB, int shB) {
            FusedOp2 o;
            o.code = code; o.offA = pA * 16; o.offB = pB * 16; o.sh = shA | (shB << 8);
            return o;
        };
        int consumed = 0;
        auto next_tip = [&](int leaf) {
            if (consumed > 0 && consumed % 32 == 0) c->fprog.push_back(mk(pbk::FOP_LOADW, 0, 0, 0, 0));
            c->tip_order.push_back(c->tip_row[leaf]);
            return 2 * (consumed++ % 32);
        };
        for (const AOp& a : aops) {
            if (a.code == pbk::FOP_TIP_TIP) {
                if ((consumed + 1) % 32 == 0) {       // the pair straddles two words: TIP ; LOADW ; MUL_TIP (same arithmetic)
                    const int sa = next_tip(a.tipA);
                    c->fprog.push_back(mk(pbk::FOP_TIP, a.pA, sa, 0, 0));
                    const int sb = next_tip(a.tipB);
                    c->fprog.push_back(mk(pbk::FOP_MUL_TIP, a.pB, sb, 0, 0));
                } else {
                    const int sa = next_tip(a.tipA), sb = next_tip(a.tipB);
                    c->fprog.push_back(mk(pbk::FOP_TIP_TIP, a.pA, sa, a.pB, sb));
                }
            } else if (a.code == pbk::FOP_TIP || a.code == pbk::FOP_MUL_TIP) {
                const int sa = next_tip(a.tipA);
                c->fprog.push_back(mk(a.code, a.pA, sa, 0, 0));
            } else if (a.code == pbk::FOP_APPLY_MUL_TIP) {
                const int sb = next_tip(a.tipB);
                c->fprog.push_back(mk(a.code, a.pA, 0, a.pB, sb));
            } else {
                c->fprog.push_back(mk(a.code, a.pA, 0, a.pB, 0));
            }
        }
        c->fwords = (consumed + 31) / 32;
        // fold every PUSH into the instruction before it (skipping window loads, which commute with it)
        std::vector<FusedOp2> folded;
        for (const FusedOp2& o : c->fprog) {
            if (o.code == pbk::FOP_PUSH) {
                int j = (int)folded.size() - 1;
                while (j >= 0 && folded[j].code == pbk::FOP_LOADW) j--;
                if (j >= 0 && !(folded[j].code & pbk::FOP_PUSH_AFTER)) { folded[j].code |= pbk::FOP_PUSH_AFTER; continue; }
                FusedOp2 p2 = o;
                p2.code = pbk::FOP_PUSH | pbk::FOP_PUSH_AFTER;      // stand-alone push
                folded.push_back(p2);
                continue;
            }
            folded.push_back(o);
        }
        c->fprog.swap(folded);
        // fused4c addressing: inner matrices by compact index into constant memory, leaf matrices by
        // compact index into shared memory
        c->inner_nodes.clear();
        c->leaf_nodes.clear();
        std::vector<int> compact(nn, -1);
        for (int v = 0; v < nn; v++) {
            if (v == c->root) continue;
            if (is_inner(v)) { compact[v] = (int)c->inner_nodes.size(); c->inner_nodes.push_back(v); }
            else { compact[v] = (int)c->leaf_nodes.size(); c->leaf_nodes.push_back(v); }
        }
        c->fprog_c = c->fprog;
        for (FusedOp2& o : c->fprog_c) {
            const int code = o.code & 0xff;
            const bool usesA = code == pbk::FOP_TIP_TIP || code == pbk::FOP_TIP || code == pbk::FOP_MUL_TIP ||
                               code == pbk::FOP_APPLY || code == pbk::FOP_APPLY_MUL_TIP || code == pbk::FOP_ACC_POP;
            const bool usesB = code == pbk::FOP_TIP_TIP || code == pbk::FOP_APPLY_MUL_TIP || code == pbk::FOP_ACC_POP;
            if (usesA) o.offA = compact[o.offA / 16] * 16;
            if (usesB) o.offB = compact[o.offB / 16] * 16;
        }
        c->tab_leaves = -1;                     // the tabulated variant is built on demand (build_table_program)
        c->r_leaves = -1;                       // ... and so is the program over the reduced tree (build_reduced_program)
        c->dense_reduced_built = false;         // ... and its dense-alphabet form (build_dense_reduced_program)
        // fused_dmma: compact matrix indices, tip rows, and the order the inner matrices are consumed in
        c->fprog_d.clear();
        c->mlist.clear();
        c->dstack = 0;
        int sp = 0;
        for (const AOp& a : aops) {
            if (a.code == pbk::FOP_PUSH) {
                c->dstack = std::max(c->dstack, ++sp);
                if (!c->fprog_d.empty() && !(c->fprog_d.back().code & pbk::FOP_PUSH_AFTER)) { c->fprog_d.back().code |= pbk::FOP_PUSH_AFTER; continue; }
                c->fprog_d.push_back({pbk::FOP_PUSH | pbk::FOP_PUSH_AFTER, 0, 0, 0});
                continue;
            }
            FusedOp2 o{a.code, 0, 0, 0};
            const int rowA = a.tipA >= 0 ? c->tip_row[a.tipA] : 0, rowB = a.tipB >= 0 ? c->tip_row[a.tipB] : 0;
            o.sh = rowA | (rowB << 16);
            switch (a.code) {
            case pbk::FOP_TIP_TIP: o.offA = compact[a.pA]; o.offB = compact[a.pB]; break;
            case pbk::FOP_TIP: case pbk::FOP_MUL_TIP: o.offA = compact[a.pA]; break;
            case pbk::FOP_APPLY: o.offA = compact[a.pA]; c->mlist.push_back(o.offA); break;
            case pbk::FOP_APPLY_MUL_TIP: o.offA = compact[a.pA]; o.offB = compact[a.pB]; c->mlist.push_back(o.offA); break;
            case pbk::FOP_ACC_POP: o.offA = compact[a.pA]; o.offB = compact[a.pB]; c->mlist.push_back(o.offA); c->mlist.push_back(o.offB); sp--; break;
            case pbk::FOP_POP_MUL: sp--; break;
            default: break;
            }
            c->fprog_d.push_back(o);
        }
    }
    c->pack_dirty = true;
    {
        static std::atomic<uint64_t> next_version{1};
        c->fprog_version = next_version++;
    }

    if (c->plan_only) return PB_OK;          // pb_debug_fused_program: host-side compilation only, no device

    // upload
    PB_CUDA(c, c->d_fprog_d.ensure(std::max<size_t>(c->fprog_d.size(), 1)));
    PB_CUDA(c, c->d_mlist.ensure(std::max<size_t>(c->mlist.size(), 1)));
    PB_CUDA(c, cudaMemcpyAsync(c->d_fprog_d.p, c->fprog_d.data(), c->fprog_d.size() * sizeof(FusedOp2), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_mlist.p, c->mlist.data(), c->mlist.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, c->d_lv_nodes.ensure(c->lv_nodes.size()));
    PB_CUDA(c, c->d_lv_children.ensure(c->lv_children.size()));
    PB_CUDA(c, c->d_fprog.ensure(c->fprog.size()));
    PB_CUDA(c, c->d_sq_nodes.ensure(c->sq_nodes.size()));
    PB_CUDA(c, c->d_sq_children.ensure(c->sq_children.size()));
    PB_CUDA(c, cudaMemcpyAsync(c->d_sq_nodes.p, c->sq_nodes.data(), c->sq_nodes.size() * sizeof(LevelNode), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_sq_children.p, c->sq_children.data(), c->sq_children.size() * sizeof(LevelChild), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_lv_nodes.p, c->lv_nodes.data(), c->lv_nodes.size() * sizeof(LevelNode), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_lv_children.p, c->lv_children.data(), c->lv_children.size() * sizeof(LevelChild), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_fprog.p, c->fprog.data(), c->fprog.size() * sizeof(FusedOp2), cudaMemcpyHostToDevice, c->stream));
    {
        std::vector<uint8_t> leaf(nn);
        for (int v = 0; v < nn; v++) leaf[v] = c->child_ptr[v] == c->child_ptr[v + 1];
        PB_CUDA(c, c->d_is_leaf.ensure(nn));
        PB_CUDA(c, cudaMemcpy(c->d_is_leaf.p, leaf.data(), nn, cudaMemcpyHostToDevice));
    }
    PB_CUDA(c, c->d_fprog_c.ensure(c->fprog_c.size()));
    PB_CUDA(c, c->d_inner_nodes.ensure(std::max<size_t>(c->inner_nodes.size(), 1)));
    PB_CUDA(c, c->d_leaf_nodes.ensure(std::max<size_t>(c->leaf_nodes.size(), 1)));
    PB_CUDA(c, cudaMemcpyAsync(c->d_fprog_c.p, c->fprog_c.data(), c->fprog_c.size() * sizeof(FusedOp2), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_inner_nodes.p, c->inner_nodes.data(), c->inner_nodes.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_leaf_nodes.p, c->leaf_nodes.data(), c->leaf_nodes.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, c->d_tip_order.ensure(c->tip_order.size()));
    PB_CUDA(c, cudaMemcpyAsync(c->d_tip_order.p, c->tip_order.data(), c->tip_order.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PB_OK;
}

// ---- P(t) ----------------------------------------------------------------
int build_pmatrices(pb_ctx* c)
{
    if (c->psrc == PSRC_NONE) return fail(c, PB_ERR_STATE, "no transition model: call pb_set_eigensystem / pb_set_rate_matrix / pb_set_pmatrices");
    if (c->psrc == PSRC_DIRECT || !c->p_dirty) return PB_OK;
    if (!c->have_blen) return fail(c, PB_ERR_STATE, "branch lengths not set");
    const int n = c->n;
    PB_CUDA(c, c->d_P.ensure((size_t)c->ncat * c->nnodes * n * n));
    if (c->psrc == PSRC_EIGEN) {
        const size_t smem = (size_t)(n + 2 * n * n) * sizeof(double);
        PB_CUDA(c, cudaFuncSetAttribute(pbk::pmat_eigen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(c->nnodes, c->ncat);
        pbk::pmat_eigen<<<grid, 256, smem, c->stream>>>(n, c->nnodes, c->root, c->d_U.p, c->d_lambda.p, c->d_Uinv.p,
                                                        c->d_blen.p, c->d_rates.p, c->d_P.p);
        c->launches++;
    } else {
        int rc = pbk::launch_expm_batched(n, c->nnodes, c->ncat, c->root, c->d_Q.p, c->d_blen.p, c->d_rates.p,
                                          c->d_P.p, c->smem_optin, c->stream);
        if (rc != 0) return fail(c, PB_ERR_CUDA, "expm kernel: not enough shared memory for this state count");
        c->launches++;
    }
    PB_CUDA(c, cudaGetLastError());
    c->p_dirty = false;
    return PB_OK;
}

template <int N, bool MASKS>
int launch_level_dense_m(pb_ctx* c, const LevelNode* nodes, int first, int count, int store_root)
{
    const size_t smem = (size_t)(N * N + N * 128) * sizeof(double);
    auto kern = pbk::level_dense<N, MASKS>;
    PB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(c->spad / 128), count);
    kern<<<grid, 128, smem, c->stream>>>(nodes, c->d_lv_children.p, first, c->d_P.p, c->nnodes, c->ncat,
                                         c->d_tips.p, c->d_masks.p, c->spad, c->d_pool.p, c->slot_stride, c->spad,
                                         c->d_pi.p, c->threshold, c->shift_at_root, c->d_site_cat.p, store_root);
    c->launches++;
    return PB_OK;
}

// Any other state count (2..64): the runtime-n form with register capacity NP.
template <int NP, bool MASKS>
int launch_level_dense_rt_m(pb_ctx* c, const LevelNode* nodes, int first, int count, int store_root)
{
    const size_t smem = (size_t)(NP * NP + NP * 128) * sizeof(double);
    auto kern = pbk::level_dense_rt<NP, MASKS>;
    PB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(c->spad / 128), count);
    kern<<<grid, 128, smem, c->stream>>>(c->n, nodes, c->d_lv_children.p, first, c->d_P.p, c->nnodes, c->ncat,
                                         c->d_tips.p, c->d_masks.p, c->spad, c->d_pool.p, c->slot_stride, c->spad,
                                         c->d_pi.p, c->threshold, c->shift_at_root, c->d_site_cat.p, store_root);
    c->launches++;
    return PB_OK;
}

int launch_level_dense_rt(pb_ctx* c, const LevelNode* nodes, int first, int count, int store_root)
{
#define PB_RT(NP) (c->have_masks ? launch_level_dense_rt_m<NP, true>(c, nodes, first, count, store_root) \
                                 : launch_level_dense_rt_m<NP, false>(c, nodes, first, count, store_root))
    if (c->n <= 8) return PB_RT(8);
    if (c->n <= 16) return PB_RT(16);
    if (c->n <= 32) return PB_RT(32);
    return PB_RT(64);
#undef PB_RT
}

template <int N>
int launch_level_dense(pb_ctx* c, const LevelNode* nodes, int first, int count, int store_root)
{
    return c->have_masks ? launch_level_dense_m<N, true>(c, nodes, first, count, store_root)
                         : launch_level_dense_m<N, false>(c, nodes, first, count, store_root);
}

// Fused pruning of the site columns [site0, site1) (multiples of the tile size; the whole shard by default).
template <int BLOCK, int SPT, int EXACT>
int launch_fused4_variant(pb_ctx* c, int64_t site0, int64_t site1)
{
    auto kern = pbk::fused4<BLOCK, SPT, EXACT>;
    const size_t smem = fused_smem_bytes(c);
    PB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, smem));
    if (occ < 1) return fail(c, PB_ERR_CUDA, "fused4: kernel does not fit on an SM");
    const int64_t tile0 = site0 / (BLOCK * SPT), ntiles = (site1 - site0) / (BLOCK * SPT);
    int64_t gx = ((int64_t)c->sm_count * occ + c->ncat - 1) / c->ncat;
    gx = std::max<int64_t>(1, std::min<int64_t>(gx, ntiles));
    dim3 grid((unsigned)gx, c->ncat);
    kern<<<grid, BLOCK, smem, c->stream>>>(c->d_fprog.p, (int)c->fprog.size(), std::max(c->fstack, 1), c->d_P.p,
                                           c->d_is_leaf.p, c->nnodes, c->d_packed.p, c->spad, tile0, ntiles, c->d_pi.p, c->threshold,
                                           c->shift_at_root, c->d_site_cat.p, c->d_site_exp.p);
    c->launches++;
    return PB_OK;
}

// true when every node of the launch has only leaf children (the cherries of level 1)
bool segment_all_tips(const pb_ctx* c, int first, int count)
{
    for (int q = first; q < first + count; q++) {
        const LevelNode& nd = c->lv_nodes[q];
        for (int k = 0; k < nd.nchild; k++)
            if (c->lv_children[nd.child_off + k].kind != 0) return false;
    }
    return true;
}

size_t fused4c_smem_for(const pb_ctx* c, int block, int spt)
{
    return (size_t)c->leaf_nodes.size() * 16 * 8 + (size_t)std::max(c->fstack, 1) * 4 * block * spt * 8;
}

// The requested shape if its stack fits shared memory, else the next smaller compiled one.
bool fused4c_shape(const pb_ctx* c, int* block, int* spt)
{
    const int cand[5][2] = {{c->fused_block, c->fused_spt}, {256, 2}, {128, 2}, {256, 1}, {0, 0}};
    for (const auto& k : cand) {
        const int key = k[0] * 10 + k[1];
        const bool compiled = key == 2561 || key == 2562 || key == 2564 || key == 1282 || key == 1284 || key == 5121 || key == 5122;
        if (compiled && fused4c_smem_for(c, k[0], k[1]) <= c->smem_optin) { *block = k[0]; *spt = k[1]; return true; }
    }
    return false;
}

template <int BLOCK, int SPT, int SCALING>
int launch_fused4c_variant(pb_ctx* c, int cat0, int ncats, int slot0, int64_t site0, int64_t site1)
{
    auto kern = pbk::fused4c<BLOCK, SPT, SCALING>;
    const size_t smem = fused4c_smem_for(c, BLOCK, SPT);
    PB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, smem));
    if (occ < 1) return fail(c, PB_ERR_CUDA, "fused4c: kernel does not fit on an SM");
    const int64_t tile0 = site0 / (BLOCK * SPT), ntiles = (site1 - site0) / (BLOCK * SPT);
    int64_t gx = ((int64_t)c->sm_count * occ + ncats - 1) / ncats;
    gx = std::max<int64_t>(1, std::min<int64_t>(gx, ntiles));
    kern<<<dim3((unsigned)gx, ncats), BLOCK, smem, c->stream>>>(
        c->d_fprog_c.p, c->fprog_bank_nops, std::max(c->fstack, 1), c->d_P.p, c->d_leaf_nodes.p,
        (int)c->leaf_nodes.size(), c->nnodes, cat0, slot0, (int)c->inner_nodes.size(), c->d_packed.p, c->spad, tile0, ntiles,
        c->d_pi.p, c->threshold, c->shift_at_root, c->d_cherry_v.p, c->d_cherry_k.p, c->ntab,
        c->d_site_cat.p, c->d_site_exp.p);
    c->launches++;
    return PB_OK;
}

template <int SCALING>
int launch_fused4c_shape(pb_ctx* c, int block, int spt, int cat0, int k, int slot0, int64_t site0, int64_t site1)
{
    switch (block * 10 + spt) {
    case 2562: return launch_fused4c_variant<256, 2, SCALING>(c, cat0, k, slot0, site0, site1);
    case 2564: return launch_fused4c_variant<256, 4, SCALING>(c, cat0, k, slot0, site0, site1);
    case 1282: return launch_fused4c_variant<128, 2, SCALING>(c, cat0, k, slot0, site0, site1);
    case 1284: return launch_fused4c_variant<128, 4, SCALING>(c, cat0, k, slot0, site0, site1);
    case 5121: return launch_fused4c_variant<512, 1, SCALING>(c, cat0, k, slot0, site0, site1);
    case 5122: return launch_fused4c_variant<512, 2, SCALING>(c, cat0, k, slot0, site0, site1);
    default: return launch_fused4c_variant<256, 1, SCALING>(c, cat0, k, slot0, site0, site1);
    }
}

bool fused_const_ok(const pb_ctx* c)
{
    int b = 0, q = 0;
    return c->fused_const && !c->fused_exact_log && !c->inner_nodes.empty() &&
           (int)c->inner_nodes.size() <= pbk::FUSED_CONST_MATS && (int)c->inner_nodes.size() <= 128 &&
           (int)c->fprog_c.size() <= pbk::FUSED_CONST_OPS && fused4c_shape(c, &b, &q);
}

// The largest clade that is tabulated (power-of-two mode only): a table of 4^k entries is only worth building when
// the alignment has many more sites than that.
int clade_limit(const pb_ctx* c)
{
    int want = (c->fused_cherry && c->fused_pow2) ? (c->fused_triple ? c->fused_clade_leaves : 2) : 0;
    while (!c->fused_clade_force && want > 2 && (int64_t)8 << (2 * want) > c->nsites) want--;
    return want;
}

// Finds the clades of at most `max_leaves` leaves whose instructions can be replaced by one table look-up and
// writes the program variant that does so (pbk::build_clade_tables explains the tables).  A clade qualifies when
// its leaves are consumed back to back from ONE packed-tip word: the walk below keeps, for the accumulator and
// every stack entry, the instruction range that produced it, how many tips it consumed and where the first one
// sits in the window; a window load (FOP_LOADW) closes everything live for further merging.
int build_table_program(pb_ctx* c, int max_leaves, bool preapply = false)
{
    struct Sym { int start, k, sh0; bool open; };
    struct Cand { int start, end, k, sh0; };
    const std::vector<FusedOp2>& prog = c->fprog_c;
    std::vector<Cand> cands;
    if (max_leaves >= 2) {
        Sym acc{0, 0, 0, false};
        std::vector<Sym> stack;
        for (int i = 0; i < (int)prog.size(); i++) {
            const FusedOp2& o = prog[i];
            const int code = o.code & 0xff, shA = o.sh & 0xff, shB = o.sh >> 8;
            bool produces = true;
            if (code == pbk::FOP_LOADW) {
                acc.open = false;
                for (Sym& e : stack) e.open = false;
                produces = false;
            } else if (code == pbk::FOP_TIP_TIP) {
                acc = {i, 2, shA, shB == shA + 2};
            } else if (code == pbk::FOP_TIP) {
                acc = {i, 1, shA, true};
            } else if (code == pbk::FOP_MUL_TIP) {
                acc.open = acc.open && shA == acc.sh0 + 2 * acc.k;
                acc.k++;
            } else if (code == pbk::FOP_APPLY_MUL_TIP) {
                acc.open = acc.open && shB == acc.sh0 + 2 * acc.k;
                acc.k++;
            } else if (code == pbk::FOP_ACC_POP || code == pbk::FOP_POP_MUL) {
                if (stack.empty()) return fail(c, PB_ERR_STATE, "fused program: stack underflow");
                const Sym x = stack.back();
                stack.pop_back();
                acc = {x.start, x.k + acc.k, x.sh0, x.open && acc.open && acc.sh0 == x.sh0 + 2 * x.k};
            } else if (code == pbk::FOP_APPLY) {
                // the branch matrix of a finished subtree: same tips
            } else {                                     // FOP_PUSH (stand-alone), FOP_ROOT
                if (code == pbk::FOP_ROOT) acc.open = false;
                produces = false;
            }
            if (produces && acc.open && acc.k >= 2 && acc.k <= max_leaves) cands.push_back({acc.start, i, acc.k, acc.sh0});
            if (o.code & pbk::FOP_PUSH_AFTER) stack.push_back(acc);
        }
    }
    // maximal candidates (ranges are nested or disjoint: keep those not inside a kept one), largest first
    std::vector<Cand> kept;
    for (int j = (int)cands.size() - 1; j >= 0; j--) {
        bool inside = false;
        for (const Cand& kq : kept) inside = inside || (cands[j].start >= kq.start && cands[j].end <= kq.end);
        if (!inside) kept.push_back(cands[j]);
    }
    std::sort(kept.begin(), kept.end(), [](const Cand& a, const Cand& b) { return a.start < b.start; });
    c->clades.clear();
    c->fprog_tab.clear();
    c->ntab = 0;
    size_t next = 0;
    for (int i = 0; i < (int)prog.size(); i++) {
        if (next < kept.size() && kept[next].start == i) {
            const Cand& kq = kept[next++];
            pbk::CladeTab t{};
            t.op0 = kq.start; t.op1 = kq.end; t.sh0 = kq.sh0; t.k = kq.k; t.ent0 = c->ntab; t.pre = -1;
            c->clades.push_back(t);
            FusedOp2 o;
            o.code = pbk::FOP_CHERRY | (prog[kq.end].code & pbk::FOP_PUSH_AFTER);
            o.offA = c->ntab;                            // first table entry; index = (window >> sh) & offB
            o.offB = (1 << (2 * kq.k)) - 1;
            o.sh = kq.sh0;
            c->fprog_tab.push_back(o);
            c->ntab += 1 << (2 * kq.k);
            i = kq.end;
            continue;
        }
        c->fprog_tab.push_back(prog[i]);
    }
    if (preapply) {
        // Which instruction consumes each table value, and does it apply a branch matrix to it first?  Walk the
        // program with, for the accumulator and every stack entry, the index of the clade whose raw table value it
        // still is (-1 otherwise).  FOP_ACC_POP / FOP_APPLY_MUL_TIP apply a matrix to such an operand: the table gets
        // that matrix (CladeTab::pre) and the instruction a flag to skip the product.
        int acc = -1, next_clade = 0;
        std::vector<int> stack;
        for (FusedOp2& o : c->fprog_tab) {
            const int code = o.code & 0xff;
            if (code == pbk::FOP_CHERRY) {
                acc = next_clade++;
            } else if (code == pbk::FOP_ACC_POP) {
                const int b = stack.back();
                stack.pop_back();
                if (acc >= 0) { c->clades[acc].pre = o.offA; o.code |= pbk::FOP_PRE_A; }
                if (b >= 0) { c->clades[b].pre = o.offB; o.code |= pbk::FOP_PRE_B; }
                acc = -1;
            } else if (code == pbk::FOP_APPLY_MUL_TIP) {
                if (acc >= 0) { c->clades[acc].pre = o.offA; o.code |= pbk::FOP_PRE_A; }
                acc = -1;
            } else if (code == pbk::FOP_POP_MUL) {
                stack.pop_back();
                acc = -1;
            } else if (code != pbk::FOP_LOADW && code != pbk::FOP_PUSH) {
                acc = -1;                              // FOP_TIP_TIP, FOP_TIP, FOP_MUL_TIP, FOP_APPLY, FOP_ROOT: no table value survives
            }
            if (o.code & pbk::FOP_PUSH_AFTER) stack.push_back(acc);
        }
    }
    c->tab_leaves = max_leaves;
    c->tables_stale = true;
    if (!c->clades.empty() && !c->plan_only) {
        PB_CUDA(c, c->d_fprog_tab.ensure(c->fprog_tab.size()));
        PB_CUDA(c, c->d_clades.ensure(c->clades.size()));
        // (pageable host vectors: the copies complete before the calls return)
        PB_CUDA(c, cudaMemcpyAsync(c->d_fprog_tab.p, c->fprog_tab.data(), c->fprog_tab.size() * sizeof(FusedOp2), cudaMemcpyHostToDevice, c->stream));
        PB_CUDA(c, cudaMemcpyAsync(c->d_clades.p, c->clades.data(), c->clades.size() * sizeof(pbk::CladeTab), cudaMemcpyHostToDevice, c->stream));
    }
    return PB_OK;
}

// ---- the program over the reduced tree (kernel fused4t) --------------------------------------------------
// Leaves of the reduced tree: every maximal clade of at most `max_leaves` leaves (one table each) and the tips
// outside such clades.  The emitter below follows the same rules as compile_tree (binary nodes: the child that
// needs more stack first; other arities: the reference's left fold), with a table value or a tip as a direct
// operand of its parent's instruction.  The same emitter, called on a clade's own subtree without tables, gives
// the instruction range build_clade_tables interprets to fill that clade's table: the order of every
// floating-point operation is the untabulated program's.
struct RAbs { int code, a, b; };           // abstract instruction: operand nodes (branch = node at its lower end)

void emit_reduced(const pb_ctx* c, int top, const std::vector<char>& kind, const std::vector<int>& su,
                  std::vector<RAbs>& out)
{
    enum { K_INNER = 0, K_TIP = 1, K_TAB = 2 };
    struct Item { int node; RAbs emit; bool is_emit; };
    std::vector<Item> st;
    st.push_back({top, {}, false});
    while (!st.empty()) {
        Item it = st.back();
        st.pop_back();
        if (it.is_emit) { out.push_back(it.emit); continue; }
        const int v = it.node;
        const int c0 = c->child_ptr[v], c1 = c->child_ptr[v + 1];
        std::vector<Item> seq;
        if (c1 - c0 == 2) {
            int a = c->child_idx[c0], b = c->child_idx[c0 + 1];
            const int ka = kind[a], kb = kind[b];
            if (ka == K_INNER && kb == K_INNER) {
                if (su[b] > su[a]) std::swap(a, b);
                seq.push_back({a, {}, false});
                seq.push_back({0, {pbk::FOP_PUSH, 0, 0}, true});
                seq.push_back({b, {}, false});
                seq.push_back({0, {pbk::FOP_ACC_POP, b, a}, true});
            } else if (ka == K_INNER || kb == K_INNER) {
                const int in = ka == K_INNER ? a : b, lf = ka == K_INNER ? b : a;
                seq.push_back({in, {}, false});
                seq.push_back({0, {kind[lf] == K_TIP ? (int)pbk::FOP_APPLY_MUL_TIP : pbk::FOP_APPLY_MUL_TAB, in, lf}, true});
            } else if (ka == K_TIP && kb == K_TIP) {
                seq.push_back({0, {pbk::FOP_TIP_TIP, a, b}, true});
            } else if (ka == K_TAB && kb == K_TAB) {
                seq.push_back({0, {pbk::FOP_TAB_TAB, a, b}, true});
            } else {
                seq.push_back({0, {pbk::FOP_TAB_TIP, ka == K_TAB ? a : b, ka == K_TAB ? b : a}, true});
            }
        } else {
            for (int k = c0; k < c1; k++) {                  // reference order (left fold)
                const int ch = c->child_idx[k];
                if (kind[ch] == K_TIP) {
                    seq.push_back({0, {k == c0 ? (int)pbk::FOP_TIP : (int)pbk::FOP_MUL_TIP, ch, 0}, true});
                } else if (kind[ch] == K_TAB) {
                    seq.push_back({0, {k == c0 ? pbk::FOP_TAB : pbk::FOP_MUL_TAB, ch, 0}, true});
                } else {
                    if (k > c0) seq.push_back({0, {pbk::FOP_PUSH, 0, 0}, true});
                    seq.push_back({ch, {}, false});
                    seq.push_back({0, {pbk::FOP_APPLY, ch, 0}, true});
                    if (k > c0) seq.push_back({0, {pbk::FOP_POP_MUL, 0, 0}, true});
                }
            }
        }
        for (auto r = seq.rbegin(); r != seq.rend(); ++r) st.push_back(*r);
    }
}

// stack need of every node of a tree whose leaves are the nodes with kind != 0 (Sethi-Ullman, as in compile_tree)
void stack_need(const pb_ctx* c, const std::vector<int>& postorder, const std::vector<char>& kind, std::vector<int>& su)
{
    su.assign(c->nnodes, 0);
    for (int v : postorder) {
        if (kind[v] != 0) continue;
        const int c0 = c->child_ptr[v], c1 = c->child_ptr[v + 1];
        if (c1 - c0 == 2) {
            const int a = c->child_idx[c0], b = c->child_idx[c0 + 1];
            const bool ia = kind[a] == 0, ib = kind[b] == 0;
            if (ia && ib) su[v] = std::max(std::max(su[a], su[b]), std::min(su[a], su[b]) + 1);
            else su[v] = ia ? su[a] : (ib ? su[b] : 0);
        } else {
            int need = 0;
            for (int k = c0; k < c1; k++) {
                const int ch = c->child_idx[k];
                if (kind[ch] == 0) need = std::max(need, su[ch] + (k > c0 ? 1 : 0));
            }
            su[v] = need;
        }
    }
}

// folds every stand-alone PUSH into the instruction before it (window loads commute with it)
void fold_pushes(std::vector<FusedOp2>& prog)
{
    std::vector<FusedOp2> folded;
    for (const FusedOp2& o : prog) {
        if (o.code == pbk::FOP_PUSH) {
            int j = (int)folded.size() - 1;
            while (j >= 0 && folded[j].code == pbk::FOP_LOADW) j--;
            if (j >= 0 && !(folded[j].code & pbk::FOP_PUSH_AFTER)) { folded[j].code |= pbk::FOP_PUSH_AFTER; continue; }
            FusedOp2 p2 = o;
            p2.code = pbk::FOP_PUSH | pbk::FOP_PUSH_AFTER;
            folded.push_back(p2);
            continue;
        }
        folded.push_back(o);
    }
    prog.swap(folded);
}

int build_reduced_program(pb_ctx* c, int max_leaves)
{
    const int nn = c->nnodes;
    // plain states: radix 4, a tip is a 2-bit field; state sets (pb_set_tip_masks): the alignment's distinct sets are
    // the alphabet, a clade of k tips has radix^k entries indexed in mixed radix, capped at the 4^max_leaves of the plain form
    const bool masked = c->have_masks;
    const int radix = masked ? std::max(c->mask_radix, 1) : 4;
    auto bits_for = [](int64_t entries) { int b = 0; while (((int64_t)1 << b) < entries) b++; return b; };
    std::vector<int64_t> pw(10, 1);
    for (int k = 1; k < 10; k++) pw[k] = pw[k - 1] * radix;
    int kmax = 0;
    if (max_leaves >= 2)
        for (int k = 1; k <= 8; k++)
            if (pw[k] <= ((int64_t)1 << (2 * max_leaves))) kmax = k;
    const int tipbits = masked ? bits_for(radix) : 2;
    auto is_leaf = [&](int v) { return c->child_ptr[v] == c->child_ptr[v + 1]; };
    // post-order and leaves below every node
    std::vector<int> order;
    order.reserve(nn);
    {
        std::vector<std::pair<int, int>> st;
        st.push_back({c->root, c->child_ptr[c->root]});
        while (!st.empty()) {
            auto& top = st.back();
            if (top.second < c->child_ptr[top.first + 1]) {
                const int ch = c->child_idx[top.second++];
                st.push_back({ch, c->child_ptr[ch]});
            } else {
                order.push_back(top.first);
                st.pop_back();
            }
        }
    }
    std::vector<int> nleaves(nn, 0);
    for (int v : order) {
        if (is_leaf(v)) { nleaves[v] = 1; continue; }
        for (int k = c->child_ptr[v]; k < c->child_ptr[v + 1]; k++) nleaves[v] += nleaves[c->child_idx[k]];
    }
    // kinds in the reduced tree: 0 inner, 1 loose tip, 2 tabulated clade (its root), 3 inside a clade
    std::vector<char> kind(nn, 0);
    for (int i = nn - 1; i >= 0; i--) {                       // reverse post-order: parents before children
        const int v = order[i];
        const int par = c->parent[v];
        if (par >= 0 && kind[par] >= 2) { kind[v] = 3; continue; }
        if (is_leaf(v)) kind[v] = 1;
        else kind[v] = (kmax >= 2 && nleaves[v] <= kmax) ? 2 : 0;
    }
    std::vector<int> su;
    stack_need(c, order, kind, su);
    c->rstack = kind[c->root] == 0 ? su[c->root] : 0;

    std::vector<RAbs> aops;
    if (kind[c->root] == 2) aops.push_back({pbk::FOP_TAB, c->root, 0});
    else emit_reduced(c, c->root, kind, su, aops);
    aops.push_back({pbk::FOP_ROOT, 0, 0});

    // ---- the clades: tables and their own instruction ranges
    c->rclades.clear();
    c->cprog.clear();
    c->rntab = 0;
    std::vector<int> clade_of(nn, -1);
    std::vector<std::vector<int>> clade_tips;                 // tip NODES of every clade in the order its range consumes them
    {
        std::vector<char> plain(nn);
        for (int v = 0; v < nn; v++) plain[v] = is_leaf(v) ? 1 : 0;
        std::vector<int> su_plain;
        stack_need(c, order, plain, su_plain);
        for (int v : order) {
            if (kind[v] != 2) continue;
            std::vector<RAbs> sub;
            emit_reduced(c, v, plain, su_plain, sub);
            pbk::CladeTab t{};
            t.op0 = (int)c->cprog.size();
            t.sh0 = 0;
            t.k = nleaves[v];
            t.ent0 = c->rntab;
            t.pre = v == c->root ? -1 : v * 16;               // the table holds P[v] . clv(v): the consumer's product
            t.xor_coded = (c->fused_xor && !masked) ? 1 : 0;
            t.radix = radix;
            std::vector<int> tips;
            auto field = [&](int leaf) { tips.push_back(leaf); return 2 * ((int)tips.size() - 1); };
            std::vector<FusedOp2> ops;
            for (const RAbs& a : sub) {
                FusedOp2 o{a.code, a.a * 16, a.b * 16, 0};
                switch (a.code) {
                case pbk::FOP_TIP_TIP: { const int sa = field(a.a), sb = field(a.b); o.sh = sa | (sb << 8); break; }
                case pbk::FOP_TIP: case pbk::FOP_MUL_TIP: o.sh = field(a.a); break;
                case pbk::FOP_APPLY_MUL_TIP: o.sh = field(a.b) << 8; break;
                default: break;
                }
                ops.push_back(o);
            }
            fold_pushes(ops);
            c->cprog.insert(c->cprog.end(), ops.begin(), ops.end());
            t.op1 = (int)c->cprog.size() - 1;
            clade_of[v] = (int)c->rclades.size();
            c->rclades.push_back(t);
            clade_tips.push_back(tips);
            c->rntab += (int)pw[t.k];
        }
    }

    // ---- lower: compact matrix indices, packed-tip fields (no field straddles a 64-bit word, an instruction's
    // fields sit in one word), window loads
    c->rmat_nodes.clear();
    c->rleaf_nodes.clear();
    std::vector<int> mat_idx(nn, -1), leaf_idx(nn, -1);
    auto mat = [&](int v) { if (mat_idx[v] < 0) { mat_idx[v] = (int)c->rmat_nodes.size(); c->rmat_nodes.push_back(v); } return mat_idx[v] * 16; };
    // (a loose tip's columns in shared memory: 16 doubles, or 4 per state set of the alphabet)
    auto leafm = [&](int v) { if (leaf_idx[v] < 0) { leaf_idx[v] = (int)c->rleaf_nodes.size(); c->rleaf_nodes.push_back(v); }
                              return leaf_idx[v] * (masked ? 4 * radix : 16); };
    c->rprog.clear();
    c->rorder.clear();
    c->rxorder.clear();
    c->rfields.clear();
    c->rfield_rows.clear();
    int bitpos = 0, wordno = 0;
    auto tab_bits = [&](int v) { return masked ? bits_for(pw[nleaves[v]]) : 2 * nleaves[v]; };
    auto reserve_bits = [&](int bits) {
        if (bitpos + bits > 64) {
            c->rorder.resize((c->rorder.size() + 31) / 32 * 32, -1);
            c->rxorder.resize(c->rorder.size(), -1);
            c->rprog.push_back({pbk::FOP_LOADW, 0, 0, 0});
            bitpos = 0;
            wordno++;
        }
    };
    auto tip_field = [&](int leaf) {
        const int sh = bitpos;
        if (masked) {
            c->rfields.push_back({wordno, sh, (int)c->rfield_rows.size(), 1});
            c->rfield_rows.push_back(c->tip_row[leaf]);
        } else {
            c->rorder.push_back(c->tip_row[leaf]);
            c->rxorder.push_back(-1);
        }
        bitpos += tipbits;
        return sh;
    };
    auto tab_field = [&](int v) {       // the clade's tips in the order its own range consumes them; plain states: all
        const int sh = bitpos;          // but the first as differences from the first (a table's hot entries become neighbours)
        const std::vector<int>& tips = clade_tips[clade_of[v]];
        if (masked) {
            c->rfields.push_back({wordno, sh, (int)c->rfield_rows.size(), (int)tips.size()});
            for (int leaf : tips) c->rfield_rows.push_back(c->tip_row[leaf]);
        } else {
            for (size_t j = 0; j < tips.size(); j++) {
                c->rorder.push_back(c->tip_row[tips[j]]);
                c->rxorder.push_back(j > 0 && c->fused_xor ? c->tip_row[tips[0]] : -1);
            }
        }
        bitpos += tab_bits(v);
        return sh;
    };
    auto ent = [&](int v) { return c->rclades[clade_of[v]].ent0; };
    for (const RAbs& a : aops) {
        FusedOp2 o{a.code, 0, 0, 0};
        switch (a.code) {
        case pbk::FOP_TIP_TIP: { reserve_bits(2 * tipbits); const int sa = tip_field(a.a), sb = tip_field(a.b);
                                 o.offA = leafm(a.a); o.offB = leafm(a.b); o.sh = sa | (sb << 8); break; }
        case pbk::FOP_TIP: case pbk::FOP_MUL_TIP: { reserve_bits(tipbits); o.offA = leafm(a.a); o.sh = tip_field(a.a); break; }
        case pbk::FOP_APPLY: o.offA = mat(a.a); break;
        case pbk::FOP_APPLY_MUL_TIP: { reserve_bits(tipbits); o.offA = mat(a.a); o.offB = leafm(a.b); o.sh = tip_field(a.b) << 8; break; }
        case pbk::FOP_ACC_POP: o.offA = mat(a.a); o.offB = mat(a.b); break;
        case pbk::FOP_TAB: case pbk::FOP_MUL_TAB: { reserve_bits(tab_bits(a.a)); o.offA = ent(a.a);
                                                    o.sh = tab_field(a.a) | (tab_bits(a.a) << 16); break; }
        case pbk::FOP_TAB_TAB: { reserve_bits(tab_bits(a.a) + tab_bits(a.b)); o.offA = ent(a.a); o.offB = ent(a.b);
                                 const int sa = tab_field(a.a), sb = tab_field(a.b);
                                 o.sh = sa | (sb << 8) | (tab_bits(a.a) << 16) | (tab_bits(a.b) << 24); break; }
        case pbk::FOP_TAB_TIP: { reserve_bits(tab_bits(a.a) + tipbits); o.offA = ent(a.a); o.offB = leafm(a.b);
                                 const int sa = tab_field(a.a), sb = tip_field(a.b);
                                 o.sh = sa | (sb << 8) | (tab_bits(a.a) << 16); break; }
        case pbk::FOP_APPLY_MUL_TAB: { reserve_bits(tab_bits(a.b)); o.offA = mat(a.a); o.offB = ent(a.b);
                                       o.sh = (tab_field(a.b) << 8) | (tab_bits(a.b) << 24); break; }
        default: break;                                       // FOP_PUSH, FOP_POP_MUL, FOP_ROOT
        }
        c->rprog.push_back(o);
    }
    fold_pushes(c->rprog);
    c->rorder.resize(std::max<size_t>((c->rorder.size() + 31) / 32 * 32, 32), -1);
    c->rxorder.resize(c->rorder.size(), -1);
    c->rwords = masked ? wordno + 1 : (int)c->rorder.size() / 32;
    c->r_leaves = max_leaves;
    c->r_radix = radix;
    c->r_masked = masked;
    c->r_tipbits = tipbits;
    c->tables_stale = true;
    c->pack_dirty = true;                 // the slot order follows the plan: tips packed for another limit are stale
    if (c->plan_only) return PB_OK;
    PB_CUDA(c, c->d_rprog.ensure(c->rprog.size()));
    PB_CUDA(c, c->d_cprog.ensure(std::max<size_t>(c->cprog.size(), 1)));
    PB_CUDA(c, c->d_rorder.ensure(c->rorder.size()));
    PB_CUDA(c, c->d_rxorder.ensure(c->rxorder.size()));
    PB_CUDA(c, c->d_rmat_nodes.ensure(std::max<size_t>(c->rmat_nodes.size(), 1)));
    PB_CUDA(c, c->d_rleaf_nodes.ensure(std::max<size_t>(c->rleaf_nodes.size(), 1)));
    PB_CUDA(c, c->d_rclades.ensure(std::max<size_t>(c->rclades.size(), 1)));
    // (pageable host vectors: the copies have read them when the calls return)
    PB_CUDA(c, cudaMemcpyAsync(c->d_rprog.p, c->rprog.data(), c->rprog.size() * sizeof(FusedOp2), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_cprog.p, c->cprog.data(), c->cprog.size() * sizeof(FusedOp2), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_rorder.p, c->rorder.data(), c->rorder.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_rxorder.p, c->rxorder.data(), c->rxorder.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_rmat_nodes.p, c->rmat_nodes.data(), c->rmat_nodes.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_rleaf_nodes.p, c->rleaf_nodes.data(), c->rleaf_nodes.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_rclades.p, c->rclades.data(), c->rclades.size() * sizeof(pbk::CladeTab), cudaMemcpyHostToDevice, c->stream));
    if (masked) {
        PB_CUDA(c, c->d_rfields.ensure(std::max<size_t>(c->rfields.size(), 1)));
        PB_CUDA(c, c->d_rfield_rows.ensure(std::max<size_t>(c->rfield_rows.size(), 1)));
        PB_CUDA(c, cudaMemcpyAsync(c->d_rfields.p, c->rfields.data(), c->rfields.size() * sizeof(pbk::PackField), cudaMemcpyHostToDevice, c->stream));
        PB_CUDA(c, cudaMemcpyAsync(c->d_rfield_rows.p, c->rfield_rows.data(), c->rfield_rows.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    return PB_OK;
}

size_t fused4t_smem_for(const pb_ctx* c, int block, int spt)
{
    return (size_t)c->rleaf_nodes.size() * (c->r_masked ? 4 * c->r_radix : 16) * 8 +
           (size_t)std::max(c->rstack, 1) * 4 * block * spt * 8;
}

bool fused4t_shape(const pb_ctx* c, int* block, int* spt)
{
    const int cand[5][2] = {{c->fused_block, c->fused_spt_set ? c->fused_spt : 2}, {256, 2}, {128, 2}, {256, 1}, {0, 0}};
    for (const auto& k : cand) {
        const int key = k[0] * 10 + k[1];
        const bool compiled = c->r_masked ? (key == 2561 || key == 2562)       // (the state-set form: two shapes)
                                          : (key == 2561 || key == 2562 || key == 2564 || key == 1282 || key == 1284);
        if (compiled && fused4t_smem_for(c, k[0], k[1]) <= c->smem_optin) { *block = k[0]; *spt = k[1]; return true; }
    }
    return false;
}

// Does the next evaluation run the reduced program (fused4t)?  Plans it for the current clade limit when needed.
int want_reduced(pb_ctx* c, bool* yes)
{
    *yes = false;
    c->reduced_fits = false;
    if (!c->fused_reduced || !c->fused_const || !c->fused_pow2 || c->fused_exact_log || c->n != 4 || !c->have_tree ||
        (c->flags & (PB_FLAG_KEEP_CLV | PB_FLAG_LEVEL_KERNELS)) || (c->have_masks && (!c->fused_masks || c->mask_radix < 1)))
        return PB_OK;
    const int want_leaves = clade_limit(c);
    if (c->r_leaves != want_leaves || c->r_masked != c->have_masks || (c->have_masks && c->r_radix != c->mask_radix)) {
        const int rc = build_reduced_program(c, want_leaves);
        if (rc != PB_OK) return rc;
    }
    int b = 0, q = 0;
    *yes = (int)c->rmat_nodes.size() <= pbk::FUSED_CONST_MATS && (int)c->rprog.size() <= pbk::FUSED_CONST_OPS &&
           fused4t_shape(c, &b, &q);
    c->reduced_fits = *yes;
    return PB_OK;
}

// Timing on: an event on the context's stream before / after a launch of the dominant kernel (pb_get_kernel_timing).
cudaError_t kernel_mark(pb_ctx* c)
{
    if (!c->timing || !c->kernel_timing || c->capturing) return cudaSuccess;
    if (c->kev_n == c->kev.size()) {
        cudaEvent_t e;
        cudaError_t err = cudaEventCreate(&e);
        if (err != cudaSuccess) return err;
        c->kev.push_back(e);
    }
    return cudaEventRecord(c->kev[c->kev_n++], c->stream);
}

template <int BLOCK, int SPT, int MINB = 0, bool MASKS = false>
int launch_fused4t_variant(pb_ctx* c, int cat0, int ncats, int slot0, int64_t site0, int64_t site1)
{
    auto kern = pbk::fused4t<BLOCK, SPT, MINB, MASKS>;
    const size_t smem = fused4t_smem_for(c, BLOCK, SPT);
    size_t& allowed = c->smem_allowed[(const void*)kern];                      // (the attribute only ever grows)
    if (smem > allowed) {
        PB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        allowed = smem;
    }
    int& occ = c->occ_cache[std::make_pair((const void*)kern, smem)];          // (occupancy once per kernel and size)
    if (occ == 0) {
        PB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, smem));
        if (occ < 1) { occ = 0; return fail(c, PB_ERR_CUDA, "fused4t: kernel does not fit on an SM"); }
    }
    c->fused4t_occ = occ;
    const int64_t tile0 = site0 / (BLOCK * SPT), ntiles = (site1 - site0) / (BLOCK * SPT);
    int64_t gx = ((int64_t)c->sm_count * occ + ncats - 1) / ncats;
    gx = std::max<int64_t>(1, std::min<int64_t>(gx, ntiles));
    // (the device-side flag of the table builder decides between the kernel's two bodies: exponents folded into the
    // table vectors, or gathered beside them; no tables or option "fused_fold" off: the general body)
    PB_CUDA(c, kernel_mark(c));
    kern<<<dim3((unsigned)gx, ncats), BLOCK, smem, c->stream>>>(
        (int)c->rprog.size(), std::max(c->rstack, 1), c->d_P.p, c->d_rleaf_nodes.p, (int)c->rleaf_nodes.size(), c->nnodes,
        cat0, slot0, (int)c->rmat_nodes.size(), c->d_packed.p, c->spad, tile0, ntiles, c->d_pi.p, c->threshold,
        c->shift_at_root, c->d_cherry_v.p, c->d_cherry_k.p, c->rntab, c->d_site_cat.p, c->d_site_exp.p,
        c->r_radix, (1u << c->r_tipbits) - 1u, c->d_mask_set.p,
        c->rclades.empty() ? nullptr : c->d_cherry_k.p + (size_t)c->ncat * c->rntab);
    PB_CUDA(c, kernel_mark(c));
    c->launches++;
    return PB_OK;
}

int launch_fused4t_shape(pb_ctx* c, int block, int spt, int cat0, int k, int slot0, int64_t site0, int64_t site1)
{
    if (c->r_masked)
        return spt == 2 ? launch_fused4t_variant<256, 2, 0, true>(c, cat0, k, slot0, site0, site1)
                        : launch_fused4t_variant<256, 1, 0, true>(c, cat0, k, slot0, site0, site1);
    switch (block * 10 + spt) {
    case 2562: return c->fused_occ == 4 ? launch_fused4t_variant<256, 2, 4>(c, cat0, k, slot0, site0, site1)   // 64 registers
                                        : launch_fused4t_variant<256, 2>(c, cat0, k, slot0, site0, site1);     // 72 registers
    case 2564: return launch_fused4t_variant<256, 4>(c, cat0, k, slot0, site0, site1);
    case 1282: return launch_fused4t_variant<128, 2>(c, cat0, k, slot0, site0, site1);
    case 1284: return launch_fused4t_variant<128, 4>(c, cat0, k, slot0, site0, site1);
    default: return launch_fused4t_variant<256, 1>(c, cat0, k, slot0, site0, site1);
    }
}

// Categories one launch of the reduced program covers when `ncats` are asked for (what the constant bank holds).
int fused4t_cats_per_launch(const pb_ctx* c, int ncats)
{
    const int nmat = (int)c->rmat_nodes.size();
    return std::max(1, std::min(ncats, std::min(c->fused_const_cats_set ? c->fused_const_cats : 64,
                                                pbk::FUSED_CONST_MATS / std::max(nmat, 1))));
}

// Program and matrices of the constant bank belong to whoever loaded them last (every context of the device
// shares the bank): owner and program version per device.
const pb_ctx* g_prog_owner[64] = {};
uint64_t g_prog_version[64] = {};
const pb_ctx* g_pinner_owner[64] = {};      // ... and whose matrices (which compaction of them, which categories) cPinner holds
uint64_t g_pinner_tag[64] = {};

// The reduced program over the site columns [site0, site1), rate categories [cat_lo, cat_hi).  `regather`: the
// matrices changed since the last call (new P(t)): compact them for the constant bank and refill the tables.
int launch_fused4t(pb_ctx* c, int64_t site0, int64_t site1, bool regather, int cat_lo = 0, int cat_hi = -1)
{
    if (cat_hi < 0) cat_hi = c->ncat;
    std::lock_guard<std::recursive_mutex> lock(g_const_mutex);
    cudaEvent_t& bank_free = g_const_event[c->device & 63];
    // (while an evaluation is being captured the bank's event is handled around the graph launch, pb_loglik_enqueue;
    // the lanes of a pipelined host evaluation only read the bank: loglik_host_impl holds it from its preparation on)
    bool fresh_event = false;
    if (!bank_free) { PB_CUDA(c, cudaEventCreateWithFlags(&bank_free, cudaEventDisableTiming)); fresh_event = true; }
    if (!fresh_event && !c->capturing && !c->bank_held) PB_CUDA(c, cudaStreamWaitEvent(c->stream, bank_free, 0));
    const int nmat = (int)c->rmat_nodes.size();
    int fb = 0, fq = 0;
    if (!fused4t_shape(c, &fb, &fq)) return fail(c, PB_ERR_CUDA, "fused4t: the tree does not fit shared memory");
    if (regather && nmat > 0) {
        PB_CUDA(c, c->d_Pc.ensure((size_t)c->ncat * nmat * 16));
        pbk::gather_inner_p<<<dim3(4, c->ncat), 256, 0, c->stream>>>(c->d_P.p, c->d_rmat_nodes.p, nmat, c->nnodes, c->d_Pc.p);
        c->launches++;
        c->pc_epoch++;
    }
    const uint64_t want_version = (c->fprog_version * 64 + (uint64_t)c->r_leaves * 4 + 2) * 32 + (c->r_masked ? c->r_radix : 0);
    if (c->capturing || g_prog_owner[c->device & 63] != c || g_prog_version[c->device & 63] != want_version) {
        PB_CUDA(c, cudaMemcpyToSymbolAsync(pbk::cProg, c->d_rprog.p, c->rprog.size() * sizeof(FusedOp2), 0,
                                           cudaMemcpyDeviceToDevice, c->stream));
        g_prog_owner[c->device & 63] = c;
        g_prog_version[c->device & 63] = want_version;
    }
    if (!c->rclades.empty() && (regather || c->tables_stale || c->tab_variant != 2)) {
        c->tab_variant = 2;
        PB_CUDA(c, c->d_cherry_v.ensure((size_t)c->ncat * c->rntab * 4));
        PB_CUDA(c, c->d_cherry_k.ensure((size_t)c->ncat * c->rntab + 1));      // (+ the flag of exponents that were not folded)
        int* fold_flag = c->d_cherry_k.p + (size_t)c->ncat * c->rntab;
        PB_CUDA(c, cudaMemsetAsync(fold_flag, c->fused_fold ? 0 : 0xff, sizeof(int), c->stream));
        int64_t emax = 1;
        for (const pbk::CladeTab& t : c->rclades) {
            int64_t e = 1;
            for (int j = 0; j < t.k; j++) e *= t.radix;
            emax = std::max(emax, e);
        }
        const unsigned gz = (unsigned)std::max<int64_t>(1, emax / 256);
        if (c->r_masked)
            pbk::build_clade_tables_masks<<<dim3((unsigned)c->rclades.size(), c->ncat, gz), 256, 0, c->stream>>>(
                c->d_cprog.p, c->d_rclades.p, c->d_P.p, c->nnodes, c->rntab, c->threshold, c->r_radix, c->d_mask_set.p,
                c->d_cherry_v.p, c->d_cherry_k.p, c->fused_fold ? fold_flag : nullptr);
        else
            pbk::build_clade_tables<true><<<dim3((unsigned)c->rclades.size(), c->ncat, gz), 256, 0, c->stream>>>(
                c->d_cprog.p, c->d_rclades.p, c->d_P.p, nullptr, nullptr, c->nnodes, c->rntab, c->threshold,
                c->d_cherry_v.p, c->d_cherry_k.p, c->fused_fold ? fold_flag : nullptr);
        c->launches++;
        c->tables_stale = false;
    }
    const int per_launch = std::max(1, std::min(c->fused_const_cats_set ? c->fused_const_cats : 64,
                                                pbk::FUSED_CONST_MATS / std::max(nmat, 1)));
    const int per_load = c->fused_const_group ? std::max(per_launch, pbk::FUSED_CONST_MATS / std::max(nmat, 1)) : per_launch;
    for (int g0 = cat_lo; g0 < cat_hi; g0 += per_load) {
        const int gk = std::min(per_load, cat_hi - g0);
        // (a chunked evaluation launches the same categories once per column chunk: the bank is loaded for the first)
        const uint64_t tag = (c->pc_epoch * 64 + (uint64_t)g0) * 64 + (uint64_t)gk;
        if (nmat > 0 && (c->capturing || !c->fused_const_keep || g_pinner_owner[c->device & 63] != c || g_pinner_tag[c->device & 63] != tag)) {
            PB_CUDA(c, cudaMemcpyToSymbolAsync(pbk::cPinner, c->d_Pc.p + (size_t)g0 * nmat * 16,
                                               (size_t)gk * nmat * 16 * sizeof(double), 0, cudaMemcpyDeviceToDevice, c->stream));
            g_pinner_owner[c->device & 63] = c;
            g_pinner_tag[c->device & 63] = tag;
        }
        for (int cat0 = g0; cat0 < g0 + gk; cat0 += per_launch) {
            const int k = std::min(per_launch, g0 + gk - cat0);
            if (site1 <= site0) continue;                     // (prepare only: matrices, tables and the constant bank)
            const int rc = launch_fused4t_shape(c, fb, fq, cat0, k, cat0 - g0, site0, site1);
            if (rc != PB_OK) return rc;
        }
    }
    if (!c->capturing && !c->bank_held) PB_CUDA(c, cudaEventRecord(bank_free, c->stream));
    return PB_OK;
}

// The inner matrices of as many categories as the constant bank holds per launch.
int launch_fused4c(pb_ctx* c, int64_t site0, int64_t site1, bool regather, int cat_lo = 0, int cat_hi = -1)
{
    if (cat_hi < 0) cat_hi = c->ncat;
    std::lock_guard<std::recursive_mutex> lock(g_const_mutex);
    cudaEvent_t& bank_free = g_const_event[c->device & 63];
    if (!bank_free) PB_CUDA(c, cudaEventCreateWithFlags(&bank_free, cudaEventDisableTiming));
    else PB_CUDA(c, cudaStreamWaitEvent(c->stream, bank_free, 0));     // another stream may still be reading the bank
    const int ninner = (int)c->inner_nodes.size();
    int fb = 0, fq = 0;
    if (!fused4c_shape(c, &fb, &fq)) return fail(c, PB_ERR_CUDA, "fused4c: the tree does not fit shared memory");
    if (regather) {
        PB_CUDA(c, c->d_Pc.ensure((size_t)c->ncat * ninner * 16));
        pbk::gather_inner_p<<<dim3(4, c->ncat), 256, 0, c->stream>>>(c->d_P.p, c->d_inner_nodes.p, ninner, c->nnodes,
                                                                     c->d_Pc.p);
        c->launches++;
    }
    // One category per launch: with several categories resident on an SM their 16 KB matrix sets evict
    // each other from the constant cache (measured: 4 categories per launch is slower, see
    // profiles/r01f_sweep_c2_fused4c.jsonl); per-category launches keep every block of an SM on one set.
    const int per_launch = std::max(1, std::min(c->fused_const_cats, pbk::FUSED_CONST_MATS / std::max(ninner, 1)));
    // the program shares the bank: upload it unless this context's current program is already there
    // small clades as table look-ups (power-of-two mode): the program variant for the current size limit
    // (a table of 4^k entries is only worth building when the alignment has many more sites than that)
    const int want_leaves = clade_limit(c);
    const bool want_pre = FUSED_PREAPPLY && c->fused_preapply;      // (a kernel build option, off by default)
    if (c->tab_leaves != want_leaves || c->tab_preapply != want_pre) {
        const int rc = build_table_program(c, want_leaves, want_pre);
        if (rc != PB_OK) return rc;
        c->tab_preapply = want_pre;
    }
    const bool tabulated = !c->clades.empty();
    const uint64_t want_version = c->fprog_version * 64 + (uint64_t)want_leaves * 4 + (want_pre ? 1 : 0);
    const FusedOp2* prog = tabulated ? c->d_fprog_tab.p : c->d_fprog_c.p;
    c->fprog_bank_nops = (int)(tabulated ? c->fprog_tab.size() : c->fprog_c.size());
    if (g_prog_owner[c->device & 63] != c || g_prog_version[c->device & 63] != want_version) {
        PB_CUDA(c, cudaMemcpyToSymbolAsync(pbk::cProg, prog, c->fprog_bank_nops * sizeof(FusedOp2), 0,
                                           cudaMemcpyDeviceToDevice, c->stream));
        g_prog_owner[c->device & 63] = c;
        g_prog_version[c->device & 63] = want_version;
    }
    if (tabulated && (regather || c->tables_stale || c->tab_variant != 1)) {
        c->tab_variant = 1;
        PB_CUDA(c, c->d_cherry_v.ensure((size_t)c->ncat * c->ntab * 4));
        PB_CUDA(c, c->d_cherry_k.ensure((size_t)c->ncat * c->ntab));
        int kmax = 2;
        for (const pbk::CladeTab& t : c->clades) kmax = std::max(kmax, t.k);
        const unsigned gz = (unsigned)std::max(1, (1 << (2 * kmax)) / 256);
        pbk::build_clade_tables<false><<<dim3((unsigned)c->clades.size(), c->ncat, gz), 256, 0, c->stream>>>(
            c->d_fprog_c.p, c->d_clades.p, c->d_P.p, c->d_inner_nodes.p, c->d_leaf_nodes.p, c->nnodes, c->ntab, c->threshold,
            c->d_cherry_v.p, c->d_cherry_k.p);
        c->launches++;
        c->tables_stale = false;
    }
    // The bank is loaded for as many categories as it holds (3 x 126 matrices at 128 taxa) and the launches of
    // those categories follow each other without a copy in between.
    const int per_load = c->fused_const_group ? std::max(per_launch, pbk::FUSED_CONST_MATS / std::max(ninner, 1)) : per_launch;
    for (int g0 = cat_lo; g0 < cat_hi; g0 += per_load) {
        const int gk = std::min(per_load, cat_hi - g0);
        g_pinner_owner[c->device & 63] = nullptr;             // (fused4t's record of what the bank holds)
        PB_CUDA(c, cudaMemcpyToSymbolAsync(pbk::cPinner, c->d_Pc.p + (size_t)g0 * ninner * 16,
                                           (size_t)gk * ninner * 16 * sizeof(double), 0, cudaMemcpyDeviceToDevice,
                                           c->stream));
        for (int cat0 = g0; cat0 < g0 + gk; cat0 += per_launch) {
            const int k = std::min(per_launch, g0 + gk - cat0);       // (the last launch may hold fewer)
            const int rc = c->fused_pow2 ? launch_fused4c_shape<pbk::SCALE_POW2>(c, fb, fq, cat0, k, cat0 - g0, site0, site1)
                                         : launch_fused4c_shape<pbk::SCALE_PRODUCT>(c, fb, fq, cat0, k, cat0 - g0, site0, site1);
            if (rc != PB_OK) return rc;
        }
    }
    PB_CUDA(c, cudaEventRecord(bank_free, c->stream));
    return PB_OK;
}

template <int SCALING>
int launch_fused4_shape(pb_ctx* c, int block, int spt, int64_t site0, int64_t site1)
{
    switch (block * 10 + spt) {
    case 1281: return launch_fused4_variant<128, 1, SCALING>(c, site0, site1);
    case 2561: return launch_fused4_variant<256, 1, SCALING>(c, site0, site1);
    case 2562: return launch_fused4_variant<256, 2, SCALING>(c, site0, site1);
    case 5121: return launch_fused4_variant<512, 1, SCALING>(c, site0, site1);
    case 1284: return launch_fused4_variant<128, 4, SCALING>(c, site0, site1);
    case 2564: return launch_fused4_variant<256, 4, SCALING>(c, site0, site1);
    default: return launch_fused4_variant<128, 2, SCALING>(c, site0, site1);
    }
}

int launch_fused4(pb_ctx* c, int64_t site0, int64_t site1, bool reduced)
{
    if (reduced) return launch_fused4t(c, site0, site1, site0 == 0);
    if (fused_const_ok(c)) return launch_fused4c(c, site0, site1, site0 == 0);
    int fb = c->fused_block, fq = c->fused_spt;
    if (!fused_shape(c, &fb, &fq)) return fail(c, PB_ERR_CUDA, "fused4: the tree does not fit shared memory");
    if (c->fused_exact_log) return launch_fused4_shape<pbk::SCALE_EXACT_LOG>(c, fb, fq, site0, site1);
    if (c->fused_pow2) return launch_fused4_shape<pbk::SCALE_POW2>(c, fb, fq, site0, site1);
    return launch_fused4_shape<pbk::SCALE_PRODUCT>(c, fb, fq, site0, site1);
}

// ---- dense alphabets: the program over the tree whose cherries are tables (kernel fused_dmma3) -------------
// Same emitter and stack rule as the 4-state reduced program; operands address matrices by the compact indices
// of compile_tree (inner_nodes / leaf_nodes), tips by their tip-matrix rows, tables by descriptor index.
int build_dense_reduced_program(pb_ctx* c)
{
    const int nn = c->nnodes;
    auto is_leaf = [&](int v) { return c->child_ptr[v] == c->child_ptr[v + 1]; };
    std::vector<int> order;
    order.reserve(nn);
    {
        std::vector<std::pair<int, int>> st;
        st.push_back({c->root, c->child_ptr[c->root]});
        while (!st.empty()) {
            auto& top = st.back();
            if (top.second < c->child_ptr[top.first + 1]) {
                const int ch = c->child_idx[top.second++];
                st.push_back({ch, c->child_ptr[ch]});
            } else {
                order.push_back(top.first);
                st.pop_back();
            }
        }
    }
    std::vector<char> kind(nn, 0);                            // 0 inner, 1 loose tip, 2 table (clade root), 3 inside a table's clade
    std::vector<int> third(nn, -1), cherry_of(nn, -1);        // three-leaf clades: the extra leaf and the cherry under the table's node
    for (int v = 0; v < nn; v++) kind[v] = is_leaf(v) ? 1 : 0;
    for (int v = 0; v < nn; v++) {
        if (is_leaf(v) || c->child_ptr[v + 1] - c->child_ptr[v] != 2) continue;
        const int a = c->child_idx[c->child_ptr[v]], b = c->child_idx[c->child_ptr[v] + 1];
        if (is_leaf(a) && is_leaf(b)) { kind[v] = 2; kind[a] = kind[b] = 3; }
    }
    // 20 states: a cherry joined by a leaf is a table of N^3 = 8 000 entries (one more matrix-vector product saved) -
    // while all the tables of all categories stay well inside L2 (measured at c3, 13 such clades, 20 MB: 0.572 ms against
    // 0.577 ms with cherries only - the products saved are paid back in gather latency; beyond L2 it would lose)
    if (c->fused_dmma_triple && (int64_t)c->n * c->n * c->n <= 8000) {
        std::vector<int> cand;
        for (int v : order) {                                  // (children before parents: the cherries are marked)
            if (is_leaf(v) || c->child_ptr[v + 1] - c->child_ptr[v] != 2) continue;
            const int a = c->child_idx[c->child_ptr[v]], b = c->child_idx[c->child_ptr[v] + 1];
            if ((kind[a] == 2 && kind[b] == 1) || (kind[b] == 2 && kind[a] == 1)) cand.push_back(v);
        }
        const int64_t bytes = (int64_t)c->ncat * (int64_t)cand.size() * c->n * c->n * c->n * ((c->n + 7) / 8 * 8) * 8;
        if (bytes <= ((int64_t)32 << 20))
            for (int v : cand) {
                const int a = c->child_idx[c->child_ptr[v]], b = c->child_idx[c->child_ptr[v] + 1];
                const int ch = kind[a] == 2 ? a : b;
                kind[v] = 2; cherry_of[v] = ch; third[v] = ch == a ? b : a;
                kind[ch] = 3; kind[third[v]] = 3;
            }
    }
    std::vector<int> su;
    stack_need(c, order, kind, su);
    std::vector<RAbs> aops;
    if (kind[c->root] == 2) aops.push_back({pbk::FOP_TAB, c->root, 0});
    else emit_reduced(c, c->root, kind, su, aops);
    aops.push_back({pbk::FOP_ROOT, 0, 0});

    std::vector<int> compact(nn, -1);
    for (size_t k = 0; k < c->inner_nodes.size(); k++) compact[c->inner_nodes[k]] = (int)k;
    for (size_t k = 0; k < c->leaf_nodes.size(); k++) compact[c->leaf_nodes[k]] = (int)k;
    c->dtabs.clear();
    c->dtab_nodes.clear();
    std::vector<int> tab_of(nn, -1);
    const int64_t n2 = (int64_t)c->n * c->n;
    c->dtab_entries = 0;
    c->dtab_three = false;
    for (int v : order) {
        if (kind[v] != 2) continue;
        const int ch = cherry_of[v] >= 0 ? cherry_of[v] : v;                 // the node whose two children are leaves
        const int a = c->child_idx[c->child_ptr[ch]], b = c->child_idx[c->child_ptr[ch] + 1];
        const int lc = third[v];
        tab_of[v] = (int)c->dtabs.size();
        c->dtabs.push_back(make_int4((int)c->dtab_entries, c->tip_row[a], c->tip_row[b], lc >= 0 ? c->tip_row[lc] : -1));
        c->dtab_nodes.push_back(pbk::DenseTabNodes{a, b, ch, lc, v, v == c->root ? 1 : 0, 0, 0});
        c->dtab_entries += lc >= 0 ? n2 * c->n : n2;
        c->dtab_three = c->dtab_three || lc >= 0;
    }
    c->dprog_r.clear();
    c->mlist_r.clear();
    c->dstack_r = 0;
    int sp = 0;
    for (const RAbs& a : aops) {
        if (a.code == pbk::FOP_PUSH) {
            c->dstack_r = std::max(c->dstack_r, ++sp);
            if (!c->dprog_r.empty() && !(c->dprog_r.back().code & pbk::FOP_PUSH_AFTER)) { c->dprog_r.back().code |= pbk::FOP_PUSH_AFTER; continue; }
            c->dprog_r.push_back({pbk::FOP_PUSH | pbk::FOP_PUSH_AFTER, 0, 0, 0});
            continue;
        }
        FusedOp2 o{a.code, 0, 0, 0};
        switch (a.code) {
        case pbk::FOP_TIP_TIP: o.offA = compact[a.a]; o.offB = compact[a.b]; o.sh = c->tip_row[a.a] | (c->tip_row[a.b] << 16); break;
        case pbk::FOP_TIP: case pbk::FOP_MUL_TIP: o.offA = compact[a.a]; o.sh = c->tip_row[a.a]; break;
        case pbk::FOP_APPLY: o.offA = compact[a.a]; c->mlist_r.push_back(o.offA); break;
        case pbk::FOP_APPLY_MUL_TIP: o.offA = compact[a.a]; o.offB = compact[a.b]; o.sh = c->tip_row[a.b] << 16; c->mlist_r.push_back(o.offA); break;
        case pbk::FOP_ACC_POP: o.offA = compact[a.a]; o.offB = compact[a.b]; c->mlist_r.push_back(o.offA); c->mlist_r.push_back(o.offB); sp--; break;
        case pbk::FOP_POP_MUL: sp--; break;
        case pbk::FOP_TAB: case pbk::FOP_MUL_TAB: o.offA = tab_of[a.a]; break;
        case pbk::FOP_TAB_TAB: o.offA = tab_of[a.a]; o.offB = tab_of[a.b]; break;
        case pbk::FOP_TAB_TIP: o.offA = tab_of[a.a]; o.offB = compact[a.b]; o.sh = c->tip_row[a.b] << 16; break;
        case pbk::FOP_APPLY_MUL_TAB: o.offA = compact[a.a]; o.offB = tab_of[a.b]; c->mlist_r.push_back(o.offA); break;
        default: break;
        }
        c->dprog_r.push_back(o);
    }
    c->dense_reduced_built = true;
    if (c->plan_only) return PB_OK;
    PB_CUDA(c, c->d_dprog_r.ensure(c->dprog_r.size()));
    PB_CUDA(c, c->d_mlist_r.ensure(std::max<size_t>(c->mlist_r.size(), 1)));
    PB_CUDA(c, c->d_dtabs.ensure(std::max<size_t>(c->dtabs.size(), 1)));
    PB_CUDA(c, c->d_dtab_nodes.ensure(std::max<size_t>(c->dtab_nodes.size(), 1)));
    PB_CUDA(c, cudaMemcpyAsync(c->d_dprog_r.p, c->dprog_r.data(), c->dprog_r.size() * sizeof(FusedOp2), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_mlist_r.p, c->mlist_r.data(), c->mlist_r.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_dtabs.p, c->dtabs.data(), c->dtabs.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_dtab_nodes.p, c->dtab_nodes.data(), c->dtab_nodes.size() * sizeof(pbk::DenseTabNodes), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));             // (the host vectors may be rebuilt by the next pb_set_tree)
    return PB_OK;
}

// Site-resident tensor-core path (20 / 61 states): pad the branch matrices, then one persistent launch.
// site0 < site1: version 3 over that column range only (a chunk of a pipelined host evaluation); `prepare`: the
// per-evaluation matrix panels and tables are (re)built first; prepare with an empty range builds them only.
int launch_fused_dmma_path(pb_ctx* c, int64_t site0 = 0, int64_t site1 = -1, bool prepare = true)
{
    const bool whole = site1 < 0;
    if (whole) { site0 = 0; site1 = c->spad; }
    const int n = c->n;
    const int ninner = (int)c->inner_nodes.size(), nleaf = (int)c->leaf_nodes.size();
    int ssm = 0, nslot = 0;
    size_t smem = 0;
    // version 3, power-of-two rescaling: the program over the tree whose cherries are tables
    const bool cherries = c->fused_dmma_version == 3 && c->fused_dmma_cherry && c->fused_pow2;
    if (cherries && !c->dense_reduced_built) {
        const int rc = build_dense_reduced_program(c);
        if (rc != PB_OK) return rc;
    }
    const bool red = cherries && !c->dtabs.empty();
    const int dstack = red ? c->dstack_r : c->dstack;
    const int nops_d = red ? (int)c->dprog_r.size() : (int)c->fprog_d.size();
    const int nmat_d = red ? (int)c->mlist_r.size() : (int)c->mlist.size();
    const int ntab_d = red ? (int)c->dtabs.size() : 0;
    if (c->fused_dmma_version == 3 &&
        pbk::fused_dmma3_plan(n, c->smem_optin, dstack, c->ntaxa, nops_d + ntab_d, nmat_d, &nslot, &ssm, &smem)) {
        if (c->fused_dmma_spill) {                            // test hook: keep every stack level in the L2 scratch
            smem -= (size_t)ssm * pbk::fused_dmma3_tilesz(n) * sizeof(double);
            ssm = 0;
        }
        if (c->fused_dmma_slots > 0 && c->fused_dmma_slots < nslot) {     // tuning hook: a shorter ring
            smem -= (size_t)(nslot - c->fused_dmma_slots) * pbk::fused_dmma3_panelsz(n) * sizeof(double);
            nslot = c->fused_dmma_slots;
        }
        if (prepare) {
            PB_CUDA(c, c->d_Ppad.ensure(std::max<size_t>((size_t)c->ncat * ninner * pbk::fused_dmma3_matsz(n), 1)));
            PB_CUDA(c, c->d_PTleaf.ensure(std::max<size_t>((size_t)c->ncat * nleaf * pbk::fused_dmma3_leafsz(n), 1)));
            pbk::launch_fdmma3_prepare(n, c->d_P.p, c->d_inner_nodes.p, ninner, c->d_leaf_nodes.p, nleaf, c->nnodes, c->ncat,
                                       c->d_Ppad.p, c->d_PTleaf.p, c->stream);
            c->launches++;
        }
        const int grid = pbk::fused_dmma3_grid(n, c->sm_count, c->spad);
        const int gdepth = dstack + 1 - ssm;
        PB_CUDA(c, c->d_gstack.ensure(std::max<size_t>((size_t)grid * gdepth * pbk::fused_dmma3_tilesz(n), 1)));
        pbk::FusedDmmaArgs a{red ? c->d_dprog_r.p : c->d_fprog_d.p, nops_d, red ? c->d_mlist_r.p : c->d_mlist.p, nmat_d, ssm,
                             c->d_gstack.p, gdepth,
                             c->d_Ppad.p, c->d_PTleaf.p, ninner, nleaf, c->ncat, c->d_tips.p + site0, c->spad, c->ntaxa,
                             site1 - site0, c->d_pi.p, c->threshold, c->shift_at_root, c->d_site_cat.p + site0};
        a.cat_stride = c->spad;
        if (red) {
            const size_t entries = (size_t)c->ncat * c->dtab_entries;
            if (prepare) {
                PB_CUDA(c, c->d_dtabV.ensure(entries * ((n + 7) / 8 * 8)));
                PB_CUDA(c, c->d_dtabK.ensure(entries + 1));                 // (+ the flag "an exponent was not folded")
                PB_CUDA(c, cudaMemsetAsync(c->d_dtabK.p + entries, (c->fused_fold && c->fused_pow2) ? 0 : 0xff, sizeof(int), c->stream));
                pbk::launch_fdmma3_build_cherries(n, c->d_P.p, c->nnodes, c->d_dtabs.p, c->d_dtab_nodes.p, ntab_d, c->dtab_entries,
                                                  c->dtab_three, c->ncat, c->threshold, c->d_dtabV.p, c->d_dtabK.p, c->stream,
                                                  c->sm_count, c->dense_builder_waves);
                c->launches++;
            }
            a.tabV = c->d_dtabV.p; a.tabK = c->d_dtabK.p; a.tabs = c->d_dtabs.p; a.ntab = ntab_d;
            a.tab_entries = c->dtab_entries;
        }
        if (site1 > site0) {
            PB_CUDA(c, kernel_mark(c));
            if (pbk::launch_fused_dmma3(n, c->fused_pow2 != 0, a, nslot, c->sm_count, smem, c->stream) < 1)
                return fail(c, PB_ERR_CUDA, "fused_dmma3 launch configuration failed");
            PB_CUDA(c, kernel_mark(c));
            c->launches++;
        }
        return PB_OK;
    }
    if (!whole) return fail(c, PB_ERR_STATE, "column chunks need version 3 of the fused tensor-core kernel");
    if (c->fused_dmma_version == 2 &&
        pbk::fused_dmma2_plan(n, c->smem_optin, c->dstack, c->ntaxa, (int)c->fprog_d.size(), (int)c->mlist.size(), &nslot, &ssm, &smem)) {
        if (c->fused_dmma_spill) {                            // test hook: keep every stack level in the L2 scratch
            smem -= (size_t)ssm * pbk::fused_dmma2_tilesz(n) * sizeof(double);
            ssm = 0;
        }
        if (c->fused_dmma_slots > 0 && c->fused_dmma_slots < nslot) {     // tuning hook: a shorter ring
            smem -= (size_t)(nslot - c->fused_dmma_slots) * pbk::fused_dmma2_panelsz(n) * sizeof(double);
            nslot = c->fused_dmma_slots;
        }
        PB_CUDA(c, c->d_Ppad.ensure(std::max<size_t>((size_t)c->ncat * ninner * pbk::fused_dmma2_matsz(n), 1)));
        PB_CUDA(c, c->d_PTleaf.ensure(std::max<size_t>((size_t)c->ncat * nleaf * pbk::fused_dmma_leafsz(n), 1)));
        pbk::launch_fdmma2_prepare(n, c->d_P.p, c->d_inner_nodes.p, ninner, c->d_leaf_nodes.p, nleaf, c->nnodes, c->ncat,
                                   c->d_Ppad.p, c->d_PTleaf.p, c->stream);
        c->launches++;
        const int grid = pbk::fused_dmma2_grid(n, c->sm_count, c->spad);
        const int gdepth = c->dstack + 1 - ssm;
        PB_CUDA(c, c->d_gstack.ensure(std::max<size_t>((size_t)grid * gdepth * pbk::fused_dmma2_tilesz(n), 1)));
        pbk::FusedDmmaArgs a{c->d_fprog_d.p, (int)c->fprog_d.size(), c->d_mlist.p, (int)c->mlist.size(), ssm, c->d_gstack.p, gdepth,
                             c->d_Ppad.p, c->d_PTleaf.p, ninner, nleaf, c->ncat, c->d_tips.p, c->spad, c->ntaxa, c->spad,
                             c->d_pi.p, c->threshold, c->shift_at_root, c->d_site_cat.p};
        if (pbk::launch_fused_dmma2(n, a, nslot, c->fused_dmma_stagger, c->sm_count, smem, c->stream) < 1)
            return fail(c, PB_ERR_CUDA, "fused_dmma2 launch configuration failed");
        c->launches++;
        return PB_OK;
    }
    if (!pbk::fused_dmma_plan(n, c->smem_optin, c->dstack, c->ntaxa, (int)c->fprog_d.size(), (int)c->mlist.size(), &ssm, &smem))
        return fail(c, PB_ERR_CUDA, "fused_dmma: the tree does not fit shared memory");
    PB_CUDA(c, c->d_Ppad.ensure(std::max<size_t>((size_t)c->ncat * ninner * pbk::fused_dmma_matsz(n), 1)));
    PB_CUDA(c, c->d_PTleaf.ensure(std::max<size_t>((size_t)c->ncat * nleaf * pbk::fused_dmma_leafsz(n), 1)));
    pbk::launch_fdmma_prepare(n, c->d_P.p, c->d_inner_nodes.p, ninner, c->d_leaf_nodes.p, nleaf, c->nnodes, c->ncat,
                              c->d_Ppad.p, c->d_PTleaf.p, c->stream);
    c->launches++;
    const int64_t work = c->spad / 64 * c->ncat;
    const int grid = pbk::fused_dmma_grid(n, c->sm_count, smem, work);
    if (grid < 1) return fail(c, PB_ERR_CUDA, "fused_dmma launch configuration failed");
    const int gdepth = c->dstack - ssm;
    PB_CUDA(c, c->d_gstack.ensure(std::max<size_t>((size_t)grid * gdepth * 256 * pbk::fused_dmma_frag(n), 1)));
    pbk::FusedDmmaArgs a{c->d_fprog_d.p, (int)c->fprog_d.size(), c->d_mlist.p, (int)c->mlist.size(), ssm, c->d_gstack.p, gdepth,
                         c->d_Ppad.p, c->d_PTleaf.p, ninner, nleaf, c->ncat, c->d_tips.p, c->spad, c->ntaxa, c->spad,
                         c->d_pi.p, c->threshold, c->shift_at_root, c->d_site_cat.p};
    pbk::launch_fused_dmma(n, a, grid, smem, c->stream);
    c->launches++;
    return PB_OK;
}

// After pb_loglik_host_packed2 the alignment is resident 2-bit packed only; whoever reads the byte-per-site
// matrix expands it first (once).
int ensure_tip_bytes(pb_ctx* c)
{
    if (!c->tip_bytes_stale) return PB_OK;
    const int64_t nbytes = (c->nsites + 3) / 4;
    pbk::expand_tips2<<<dim3(c->ntaxa, (unsigned)std::min<int64_t>((nbytes + 255) / 256, 2048)), 256, 0, c->stream>>>(
        c->d_tips2.p, c->spad / 4, c->d_tips.p, c->spad, 0, nbytes);
    c->launches++;
    PB_CUDA(c, cudaGetLastError());
    c->tip_bytes_stale = false;
    return PB_OK;
}

int run_pruning(pb_ctx* c)
{
    const int n = c->n;
    bool reduced = false;
    {
        const int rc = want_reduced(c, &reduced);
        if (rc != PB_OK) return rc;
    }
    const int path = path_of(c);
    PB_CUDA(c, c->d_site_cat.ensure((size_t)c->ncat * c->spad));
    PB_CUDA(c, c->d_site_exp.ensure((size_t)c->ncat * c->spad));
    if ((path != PATH_FUSED4 || c->pack_dirty) && !c->have_masks) {
        const int rc = ensure_tip_bytes(c);
        if (rc != PB_OK) return rc;
    }
    if (path == PATH_FUSED4) {
        int rc = PB_OK;
        if (c->have_masks && (c->pack_dirty || !c->packed_for_reduced)) {
            PB_CUDA(c, c->d_packed.ensure((size_t)(c->rwords + 2) * c->spad));
            pbk::pack_fields_masks<<<(unsigned)((c->spad + 255) / 256), 256, 0, c->stream>>>(
                c->d_masks.p, c->spad, c->d_rfields.p, (int)c->rfields.size(), c->d_rfield_rows.p, c->d_mask_code.p, c->r_radix,
                c->rwords, c->d_packed.p, c->spad, 0, c->spad);
            c->launches++;
            c->pack_dirty = false;
            c->packed_for_reduced = true;
        } else if (c->pack_dirty || c->packed_for_reduced != reduced) {
            if (c->tip_bytes_stale && (rc = ensure_tip_bytes(c)) != PB_OK) return rc;
            const int nwords = reduced ? c->rwords : c->fwords;
            PB_CUDA(c, c->d_packed.ensure((size_t)(nwords + 2) * c->spad));
            pbk::pack_tips4<<<(unsigned)((c->spad + 255) / 256), 256, 0, c->stream>>>(
                c->d_tips.p, c->spad, reduced ? c->d_rorder.p : c->d_tip_order.p, reduced ? c->d_rxorder.p : nullptr,
                reduced ? (int)c->rorder.size() : (int)c->tip_order.size(), nwords, c->d_packed.p, c->spad,
                0, c->spad, c->d_bad.p);
            c->launches++;
            c->pack_dirty = false;
            c->packed_for_reduced = reduced;
            c->tips_checked = true;
        }
        if (c->fused_split > 1 && reduced) {
            const int64_t step = round_up((c->spad + c->fused_split - 1) / c->fused_split, 1024);
            for (int64_t s0 = 0; s0 < c->spad; s0 += step)
                if ((rc = launch_fused4(c, s0, std::min(s0 + step, c->spad), reduced)) != PB_OK) return rc;
        } else
            rc = launch_fused4(c, 0, c->spad, reduced);
        if (rc != PB_OK) return rc;
        PB_CUDA(c, cudaGetLastError());
        return PB_OK;
    }
    if (path == PATH_FUSED_DMMA) {
        if (!c->tips_checked) {
            pbk::validate_tips<<<c->sm_count * 4, 256, 0, c->stream>>>(c->d_tips.p, (int64_t)c->ntaxa * c->spad, n, c->d_bad.p);
            c->launches++;
            c->tips_checked = true;
        }
        int rc = launch_fused_dmma_path(c);
        if (rc != PB_OK) return rc;
        PB_CUDA(c, cudaGetLastError());
        return PB_OK;
    }
    if (!c->tips_checked && !c->have_masks) {
        pbk::validate_tips<<<c->sm_count * 4, 256, 0, c->stream>>>(c->d_tips.p, (int64_t)c->ntaxa * c->spad, n, c->d_bad.p);
        c->launches++;
        c->tips_checked = true;
    }
    // level-scheduled paths
    c->slot_stride = (int64_t)c->ncat * (n + 1) * c->spad;
    PB_CUDA(c, c->d_pool.ensure((size_t)std::max(c->nslots, c->sq_nslots) * c->slot_stride));
    const int store_root = (c->flags & PB_FLAG_KEEP_CLV) ? 1 : 0;
    if (path == PATH_LEVEL_DMMA && !c->dmma_per_level) {
        // one persistent launch: every block walks the whole tree for its groups of sites
        int rc = pbk::launch_level_dmma(n, c->d_sq_nodes.p, c->d_sq_children.p, 0, (int)c->sq_nodes.size(),
                                        c->max_arity, c->d_P.p, c->nnodes, c->ncat, c->d_tips.p, c->spad,
                                        c->d_pool.p, c->slot_stride, c->spad, c->d_pi.p, c->threshold,
                                        c->shift_at_root, c->d_site_cat.p, store_root, c->sm_count, c->smem_optin,
                                        true, c->dmma_seq_tiles, c->stream);
        if (rc != 0) return fail(c, PB_ERR_CUDA, "level_dmma launch configuration failed");
        c->launches++;
        PB_CUDA(c, cudaGetLastError());
        return PB_OK;
    }
    // What to (re)compute: every level, or — when only some branch lengths changed since the last
    // evaluation and every CLV is resident (PB_FLAG_KEEP_CLV) — just the ancestors of those branches.
    const LevelNode* lvn = c->d_lv_nodes.p;
    std::vector<std::pair<int, int>> segments;      // (first, count) per launch, in level order
    const bool partial = c->partial_ok && !c->all_dirty && (c->flags & PB_FLAG_KEEP_CLV) && c->evaluated;
    if (!partial) {
        for (size_t l = 0; l + 1 < c->lv_first.size(); l++)
            segments.push_back({c->lv_first[l], c->lv_first[l + 1] - c->lv_first[l]});
    } else {
        std::vector<LevelNode> sel;
        for (size_t l = 0; l + 1 < c->lv_first.size(); l++) {
            const int begin = (int)sel.size();
            for (int q = c->lv_first[l]; q < c->lv_first[l + 1]; q++)
                if (c->dirty[c->lv_node_id[q]]) sel.push_back(c->lv_nodes[q]);
            if ((int)sel.size() > begin) segments.push_back({begin, (int)sel.size() - begin});
        }
        PB_CUDA(c, c->d_sel_nodes.ensure(std::max<size_t>(sel.size(), 1)));
        PB_CUDA(c, cudaMemcpyAsync(c->d_sel_nodes.p, sel.data(), sel.size() * sizeof(LevelNode), cudaMemcpyHostToDevice, c->stream));
        lvn = c->d_sel_nodes.p;
        c->last_nodes_recomputed = (int)sel.size();
    }
    if (!partial) c->last_nodes_recomputed = (int)c->lv_nodes.size();
    std::fill(c->dirty.begin(), c->dirty.end(), 0);
    c->all_dirty = false;
    c->partial_ok = true;
    {   // a level of a very large tree may hold more nodes than gridDim.y allows: several launches
        std::vector<std::pair<int, int>> split;
        for (const auto& seg : segments)
            for (int off = 0; off < seg.second; off += 65535)
                split.push_back({seg.first + off, std::min(65535, seg.second - off)});
        segments.swap(split);
    }
    for (const auto& seg : segments) {
        const int first = seg.first, count = seg.second;
        if (path == PATH_LEVEL4 && c->level4_async && c->max_arity == 2 && c->min_arity == 2) {
            // software-pipelined form (binary trees): CTAs walk consecutive tiles with cp.async prefetch
            const size_t smem = 32 * sizeof(double) + (size_t)2 * 2 * 5 * 256 * 16 + (size_t)2 * 2 * 256 * 4;
            if (!c->level4_async_ready) {       // per context: the attribute is per device
                PB_CUDA(c, cudaFuncSetAttribute(pbk::level4_async, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                c->level4_async_ready = true;
            }
            const int64_t ntiles = c->spad / 512;
            const int64_t units = (int64_t)count * c->ncat;
            int64_t chunks = std::max<int64_t>(1, ((int64_t)c->sm_count * 2 * c->level4_waves + units - 1) / units);
            chunks = std::min(chunks, ntiles);
            const int64_t tpc = (ntiles + chunks - 1) / chunks;
            chunks = (ntiles + tpc - 1) / tpc;
            dim3 grid((unsigned)chunks, count, c->ncat);
            pbk::level4_async<<<grid, 256, smem, c->stream>>>(lvn, c->d_lv_children.p, first, c->d_P.p, c->nnodes,
                                                              c->d_tips.p, c->spad, c->d_pool.p, c->slot_stride, c->spad,
                                                              c->d_pi.p, c->threshold, c->shift_at_root, c->d_site_cat.p,
                                                              store_root, (int)tpc);
            c->launches++;
        } else if (path == PATH_LEVEL4) {
            dim3 grid((unsigned)((c->spad / 2 + 255) / 256), count, c->ncat);
#define PB_LEVEL4_ARGS lvn, c->d_lv_children.p, first, c->d_P.p, c->nnodes, c->d_tips.p, c->spad, \
                       c->d_pool.p, c->slot_stride, c->spad, c->d_pi.p, c->threshold, c->shift_at_root,        \
                       c->d_site_cat.p, store_root
            switch (c->level4_minb) {
            case 3: pbk::level4<3><<<grid, 256, 0, c->stream>>>(PB_LEVEL4_ARGS); break;
            case 5: pbk::level4<5><<<grid, 256, 0, c->stream>>>(PB_LEVEL4_ARGS); break;
            case 6: pbk::level4<6><<<grid, 256, 0, c->stream>>>(PB_LEVEL4_ARGS); break;
            default: pbk::level4<4><<<grid, 256, 0, c->stream>>>(PB_LEVEL4_ARGS); break;
            }
#undef PB_LEVEL4_ARGS
            c->launches++;
        } else if (path == PATH_LEVEL_DMMA && c->dmma_version == 2 && c->max_arity <= 2 &&
                   !(c->dmma_tip_levels_v1 && !partial && segment_all_tips(c, first, count))) {
            int rc = pbk::launch_level_dmma2(n, lvn, c->d_lv_children.p, first, count, c->d_P.p, c->nnodes,
                                             c->ncat, c->d_tips.p, c->spad, c->d_pool.p, c->slot_stride, c->spad,
                                             c->d_pi.p, c->threshold, c->shift_at_root, c->d_site_cat.p, store_root,
                                             c->sm_count, c->smem_optin, c->stream);
            if (rc != 0) return fail(c, PB_ERR_CUDA, "level_dmma2 launch configuration failed");
            c->launches++;
        } else if (path == PATH_LEVEL_DMMA) {
            int rc = pbk::launch_level_dmma(n, lvn, c->d_lv_children.p, first, count, c->max_arity,
                                            c->d_P.p, c->nnodes, c->ncat, c->d_tips.p, c->spad, c->d_pool.p,
                                            c->slot_stride, c->spad, c->d_pi.p, c->threshold, c->shift_at_root,
                                            c->d_site_cat.p, store_root, c->sm_count, c->smem_optin, false, 0,
                                            c->stream);
            if (rc != 0) return fail(c, PB_ERR_CUDA, "level_dmma launch configuration failed");
            c->launches++;
        } else {
            int rc;
            switch (n) {
            case 2: rc = launch_level_dense<2>(c, lvn, first, count, store_root); break;
            case 3: rc = launch_level_dense<3>(c, lvn, first, count, store_root); break;
            case 4: rc = launch_level_dense<4>(c, lvn, first, count, store_root); break;
            case 5: rc = launch_level_dense<5>(c, lvn, first, count, store_root); break;
            case 20: rc = launch_level_dense<20>(c, lvn, first, count, store_root); break;
            case 61: rc = launch_level_dense<61>(c, lvn, first, count, store_root); break;
            case 64: rc = launch_level_dense<64>(c, lvn, first, count, store_root); break;
            default: rc = launch_level_dense_rt(c, lvn, first, count, store_root); break;
            }
            if (rc != PB_OK) return rc;
        }
    }
    PB_CUDA(c, cudaGetLastError());
    return PB_OK;
}

// The 4-state site-resident kernels in power-of-two mode leave likelihoods as (value, binary exponent) pairs.
const int* site_exp_of(const pb_ctx* c)
{
    return path_of(c) == PATH_FUSED4 && c->fused_pow2 && !c->fused_exact_log ? c->d_site_exp.p : nullptr;
}

int run_root_reduction(pb_ctx* c)
{
    const int64_t blocks = (c->nsites + 255) / 256;
    PB_CUDA(c, c->d_per_site.ensure((size_t)c->spad));
    PB_CUDA(c, c->d_partials.ensure((size_t)blocks));
    PB_CUDA(c, c->d_total.ensure(1));
    pbk::root_combine<<<(unsigned)blocks, 256, 0, c->stream>>>(c->d_site_cat.p, site_exp_of(c), c->ncat, c->d_weights.p,
                                                               c->have_site_weights ? c->d_site_weights.p : nullptr, c->nsites,
                                                               c->spad, 0, c->d_per_site.p, c->d_partials.p);
    c->launches++;
    pbk::final_sum<<<1, 256, 0, c->stream>>>(c->d_partials.p, blocks, c->d_total.p);
    c->launches++;
    PB_CUDA(c, cudaGetLastError());
    return PB_OK;
}

}  // namespace

// ===========================================================================

// ---- FP64 peak probes (pb_measure_fp64_peak): the denominators of the roofline fractions, measured on the device the
// benchmark runs on.  Register-resident dependent chains, 8 per thread: DFMA, or the native FP64 tensor-core shape.
namespace pbk {
__global__ void __launch_bounds__(256) fp64_peak_dfma(double* out, int iters)
{
    double c[8];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * (threadIdx.x + 1);
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) fp64_peak_dmma(double* out, int iters)
{
    double c[8][2];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * (threadIdx.x + 1);
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace pbk

extern "C" {

const char* pb_version(void) { return "phylo_b200 0.1 (sm_100a)"; }

const char* pb_last_error(const pb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int pb_last_status(const pb_ctx* ctx) { return ctx ? ctx->err_status : g_create_status; }

int pb_get_dims(const pb_ctx* c, int64_t* dims)
{
    if (!c || !dims) return PB_ERR_INVALID_ARG;
    dims[0] = c->n; dims[1] = c->ncat; dims[2] = c->nsites; dims[3] = c->ntaxa;
    dims[4] = c->have_tree ? c->nnodes : 0; dims[5] = c->flags;
    return PB_OK;
}

pb_ctx* pb_create(int device, int nstates, int ncat, int64_t nsites, int ntaxa, unsigned flags)
{
    if (nstates < 2 || nstates > 64) { fail(nullptr, PB_ERR_INVALID_ARG, "nstates must be in 2..64"); return nullptr; }
    if (ncat < 1 || ncat > 64) { fail(nullptr, PB_ERR_INVALID_ARG, "ncat must be in 1..64"); return nullptr; }
    if (nsites < 1) { fail(nullptr, PB_ERR_INVALID_ARG, "nsites must be >= 1"); return nullptr; }
    if (ntaxa < 1) { fail(nullptr, PB_ERR_INVALID_ARG, "ntaxa must be >= 1"); return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        fail(nullptr, PB_ERR_CUDA, std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
        return nullptr;
    }
    if (device < 0 || device >= ndev) { fail(nullptr, PB_ERR_INVALID_ARG, "device index out of range"); return nullptr; }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        cuda_fail(nullptr, e, "cudaSetDevice");
        return nullptr;
    }
    if (prop.major != 10) {
        fail(nullptr, PB_ERR_CUDA, "device is not sm_100 (Blackwell B200): kernels are built for sm_100a only");
        return nullptr;
    }
    pb_ctx* c = new pb_ctx();
    c->device = device;
    c->n = nstates;
    c->ncat = ncat;
    c->nsites = nsites;
    c->ntaxa = ntaxa;
    c->flags = flags;
    c->spad = round_up(nsites, 1024);
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMallocHost(&c->h_total, 2 * sizeof(double))) != cudaSuccess) {
        cuda_fail(nullptr, e, "stream/pinned allocation");
        delete c;
        return nullptr;
    }
    c->stream = c->own_stream;
    if (const char* v = getenv("PB_FUSED_VARIANT")) {     // e.g. "256x2": threads per block x sites per thread
        int b = 0, q = 0;
        if (sscanf(v, "%dx%d", &b, &q) == 2 && (b == 128 || b == 256 || b == 512) && (q == 1 || q == 2)) {
            c->fused_block = b;
            c->fused_spt = q;
        }
    }
    if (const char* v = getenv("PB_FUSED_EXACT_LOG")) c->fused_exact_log = atoi(v) != 0;
    c->h_rates.assign(ncat, 1.0);
    c->h_weights.assign(ncat, 1.0 / ncat);
    if (c->d_bad.ensure(1) != cudaSuccess || cudaMemset(c->d_bad.p, 0, sizeof(int)) != cudaSuccess ||
        c->d_rates.ensure(ncat) != cudaSuccess || c->d_weights.ensure(ncat) != cudaSuccess ||
        cudaMemcpy(c->d_rates.p, c->h_rates.data(), ncat * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->d_weights.p, c->h_weights.data(), ncat * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
        fail(nullptr, PB_ERR_NOMEM, "device allocation failed");
        pb_destroy(c);
        return nullptr;
    }
    return c;
}

void pb_destroy(pb_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    {
        std::lock_guard<std::recursive_mutex> lock(g_const_mutex);      // a later context may get this address
        if (g_prog_owner[c->device & 63] == c) g_prog_owner[c->device & 63] = nullptr;
        if (g_pinner_owner[c->device & 63] == c) g_pinner_owner[c->device & 63] = nullptr;
    }
    if (c->stream && c->stream != c->own_stream) cudaStreamSynchronize(c->stream);   // work may still use our buffers
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->own_stream) { cudaStreamSynchronize(c->own_stream); cudaStreamDestroy(c->own_stream); }
    if (c->h_total) cudaFreeHost(c->h_total);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    for (auto& set : c->ev_ring)
        for (int i = 0; i < 4; i++) if (set[i]) cudaEventDestroy(set[i]);
    for (cudaEvent_t e : c->chunk_events) cudaEventDestroy(e);
    for (cudaEvent_t e : c->trace_events) cudaEventDestroy(e);
    for (cudaEvent_t e : c->kev) cudaEventDestroy(e);
    if (c->ev_root_done) cudaEventDestroy(c->ev_root_done);
    for (int l = 0; l < 4; l++) {
        if (c->lane_events[l]) cudaEventDestroy(c->lane_events[l]);
        if (c->lane_streams[l]) { cudaStreamSynchronize(c->lane_streams[l]); cudaStreamDestroy(c->lane_streams[l]); }
    }
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    c->d_lv_nodes.release(); c->d_lv_children.release(); c->d_fprog.release();
    c->d_tips.release(); c->d_masks.release(); c->d_pi.release(); c->d_rates.release(); c->d_weights.release();
    c->d_blen.release(); c->d_U.release(); c->d_lambda.release(); c->d_Uinv.release(); c->d_Q.release();
    c->d_P.release(); c->d_pool.release(); c->d_site_cat.release(); c->d_site_exp.release(); c->d_per_site.release();
    c->d_partials.release(); c->d_total.release();
    delete c;
}

#define PB_ENTER(c)                                                          \
    if (!(c)) return PB_ERR_INVALID_ARG;                                     \
    { cudaError_t e_ = cudaSetDevice((c)->device); if (e_ != cudaSuccess) return cuda_fail((c), e_, "cudaSetDevice"); }
// setters after which a captured evaluation (option "graph") no longer describes what would be launched
#define PB_RESHAPES(c) ((c)->config_epoch++)

int pb_set_stream(pb_ctx* c, void* s)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    if ((cudaStream_t)s != c->stream) PB_CUDA(c, cudaStreamSynchronize(c->stream));   // finish what the old stream holds
    c->stream = (cudaStream_t)s;      // 0 is a valid handle: the CUDA default stream
    return PB_OK;
}

int pb_use_private_stream(pb_ctx* c)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    if (c->stream != c->own_stream) PB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stream = c->own_stream;
    return PB_OK;
}

int pb_set_mode(pb_ctx* c, int mode)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    if (mode == PB_MODE_PHYLO_CTMC) { c->threshold = 1e-6; c->shift_at_root = 1; }
    else if (mode == PB_MODE_FELSENSTEIN) { c->threshold = 0.0000001; c->shift_at_root = 0; }
    else return fail(c, PB_ERR_INVALID_ARG, "unknown mode");
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

int pb_set_rescaling(pb_ctx* c, double threshold, int shift_at_root)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    c->threshold = threshold;
    c->shift_at_root = shift_at_root ? 1 : 0;
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

int pb_set_tree(pb_ctx* c, int nnodes, int root, const int32_t* child_ptr, const int32_t* child_idx,
                const int32_t* tip_row, const double* branch_len)
{
    PB_ENTER(c);
    if (!child_ptr || !child_idx || !tip_row) return fail(c, PB_ERR_INVALID_ARG, "tree: null array");
    if (nnodes < 2) return fail(c, PB_ERR_INVALID_ARG, "tree: needs at least one internal node and one leaf");
    if (root < 0 || root >= nnodes) return fail(c, PB_ERR_INVALID_ARG, "tree: root out of range");
    if (child_ptr[0] != 0) return fail(c, PB_ERR_INVALID_ARG, "tree: child_ptr[0] must be 0");
    for (int v = 0; v < nnodes; v++)
        if (child_ptr[v + 1] < child_ptr[v]) return fail(c, PB_ERR_INVALID_ARG, "tree: child_ptr must be non-decreasing");
    const int nedges = child_ptr[nnodes];
    if (nedges != nnodes - 1) return fail(c, PB_ERR_INVALID_ARG, "tree: a rooted tree over nnodes nodes has nnodes-1 branches");
    std::vector<int> parent(nnodes, -1);
    for (int v = 0; v < nnodes; v++)
        for (int k = child_ptr[v]; k < child_ptr[v + 1]; k++) {
            const int ch = child_idx[k];
            if (ch < 0 || ch >= nnodes || ch == root) return fail(c, PB_ERR_INVALID_ARG, "tree: bad child index");
            if (parent[ch] != -1) return fail(c, PB_ERR_INVALID_ARG, "tree: node has two parents");
            parent[ch] = v;
        }
    if (child_ptr[root + 1] == child_ptr[root]) return fail(c, PB_ERR_INVALID_ARG, "tree: the root must be an internal node");
    std::vector<char> seen(c->ntaxa, 0);
    for (int v = 0; v < nnodes; v++) {
        const bool leaf = child_ptr[v + 1] == child_ptr[v];
        if (leaf) {
            if (tip_row[v] < 0 || tip_row[v] >= c->ntaxa) return fail(c, PB_ERR_INVALID_ARG, "tree: leaf tip_row out of range");
            seen[tip_row[v]] = 1;
        }
    }
    c->nnodes = nnodes;
    c->root = root;
    c->child_ptr.assign(child_ptr, child_ptr + nnodes + 1);
    c->child_idx.assign(child_idx, child_idx + nedges);
    c->tip_row.assign(tip_row, tip_row + nnodes);
    c->parent = parent;
    PB_RESHAPES(c);
    c->have_tree = true;
    c->have_blen = false;
    c->p_dirty = true;
    c->all_dirty = true;
    c->fresh = false;
    if (c->psrc == PSRC_DIRECT) c->psrc = PSRC_NONE;   // matrices belonged to the old tree
    c->evaluated = false;
    int rc = compile_tree(c);
    if (rc != PB_OK) { c->have_tree = false; return rc; }
    bool reduced = false;
    if ((rc = want_reduced(c, &reduced)) != PB_OK) { c->have_tree = false; return rc; }   // plans the 4-state program (pb_path_name)
    if (branch_len) return pb_set_branch_lengths(c, branch_len);
    return PB_OK;
}

int pb_set_branch_lengths(pb_ctx* c, const double* bl)
{
    PB_ENTER(c);
    if (!c->have_tree) return fail(c, PB_ERR_STATE, "set the tree first");
    if (!bl) return fail(c, PB_ERR_INVALID_ARG, "branch_len is null");
    for (int v = 0; v < c->nnodes; v++)
        if (v != c->root && !(bl[v] >= 0.0)) return fail(c, PB_ERR_INVALID_ARG, "branch lengths must be >= 0");
    c->branch_len.assign(bl, bl + c->nnodes);
    c->branch_len[c->root] = 0.0;
    PB_CUDA(c, c->d_blen.ensure(c->nnodes));
    // pageable source: the runtime stages it before returning, so no synchronisation is needed here
    PB_CUDA(c, cudaMemcpyAsync(c->d_blen.p, c->branch_len.data(), c->nnodes * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->have_blen = true;
    c->p_dirty = true;
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

int pb_update_branch_length(pb_ctx* c, int node, double t)
{
    PB_ENTER(c);
    if (!c->have_tree || !c->have_blen) return fail(c, PB_ERR_STATE, "set the tree and all branch lengths first");
    if (node < 0 || node >= c->nnodes || node == c->root) return fail(c, PB_ERR_INVALID_ARG, "pb_update_branch_length: no such branch");
    if (!(t >= 0.0)) return fail(c, PB_ERR_INVALID_ARG, "branch lengths must be >= 0");
    if (c->psrc == PSRC_DIRECT) return fail(c, PB_ERR_STATE, "matrices were supplied directly: branch lengths are not in use");
    c->branch_len[node] = t;
    PB_CUDA(c, cudaMemcpyAsync(c->d_blen.p + node, &c->branch_len[node], sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->p_dirty = true;
    c->fresh = false;
    for (int v = c->parent[node]; v >= 0; v = c->parent[v]) c->dirty[v] = 1;    // the path to the root
    return PB_OK;
}

int pb_nodes_recomputed(const pb_ctx* c) { return c ? c->last_nodes_recomputed : 0; }

static int alloc_tips(pb_ctx* c)
{
    const size_t bytes = (size_t)c->ntaxa * c->spad;
    const bool fresh = c->d_tips.count < bytes || !c->d_tips.p;
    PB_CUDA(c, c->d_tips.ensure(bytes));
    if (fresh) PB_CUDA(c, cudaMemsetAsync(c->d_tips.p, 0, bytes, c->stream));
    return PB_OK;
}

static int alloc_tips2(pb_ctx* c)          // 2-bit upload buffer; padding columns read as state 0
{
    const size_t bytes = (size_t)c->ntaxa * (c->spad / 4);
    const bool fresh = c->d_tips2.count < bytes || !c->d_tips2.p;
    PB_CUDA(c, c->d_tips2.ensure(bytes));
    if (fresh) PB_CUDA(c, cudaMemsetAsync(c->d_tips2.p, 0, bytes, c->stream));
    return PB_OK;
}

int pb_set_tips(pb_ctx* c, const uint8_t* states, int64_t ld)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    if (!states || ld < c->nsites) return fail(c, PB_ERR_INVALID_ARG, "tips: null pointer or ld < nsites");
    int rc = alloc_tips(c);
    if (rc != PB_OK) return rc;
    PB_CUDA(c, cudaMemcpy2DAsync(c->d_tips.p, c->spad, states, ld, c->nsites, c->ntaxa, cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemsetAsync(c->d_bad.p, 0, sizeof(int), c->stream));
    // The caller's buffer may be page-locked (pb_pin_host_memory), in which case the copy above is truly
    // asynchronous: wait until it has been read, so that the caller may change or free it once we return.
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->have_tips = true;
    c->have_masks = false;
    c->tip_bytes_stale = false;
    c->pack_dirty = true;
    c->tips_checked = false;
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

int pb_set_tips_device(pb_ctx* c, const uint8_t* d_states, int64_t ld)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    if (!d_states || ld < c->nsites) return fail(c, PB_ERR_INVALID_ARG, "tips: null pointer or ld < nsites");
    int rc = alloc_tips(c);
    if (rc != PB_OK) return rc;
    PB_CUDA(c, cudaMemcpy2DAsync(c->d_tips.p, c->spad, d_states, ld, c->nsites, c->ntaxa, cudaMemcpyDeviceToDevice, c->stream));
    PB_CUDA(c, cudaMemsetAsync(c->d_bad.p, 0, sizeof(int), c->stream));
    c->have_tips = true;
    c->have_masks = false;
    c->tip_bytes_stale = false;
    c->pack_dirty = true;
    c->tips_checked = false;
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

int pb_set_tips_packed2(pb_ctx* c, const uint8_t* packed, int64_t ld)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    const int64_t nbytes = (c->nsites + 3) / 4;
    if (c->n != 4) return fail(c, PB_ERR_STATE, "2-bit packed tips are for 4-state alphabets");
    if (!packed || ld < nbytes) return fail(c, PB_ERR_INVALID_ARG, "packed tips: null pointer or ld < ceil(nsites / 4)");
    int rc = alloc_tips(c);
    if (rc != PB_OK) return rc;
    if ((rc = alloc_tips2(c)) != PB_OK) return rc;
    PB_CUDA(c, cudaMemcpy2DAsync(c->d_tips2.p, c->spad / 4, packed, ld, nbytes, c->ntaxa, cudaMemcpyHostToDevice, c->stream));
    pbk::expand_tips2<<<dim3(c->ntaxa, (unsigned)std::min<int64_t>((nbytes + 255) / 256, 2048)), 256, 0, c->stream>>>(
        c->d_tips2.p, c->spad / 4, c->d_tips.p, c->spad, 0, nbytes);
    c->launches++;
    PB_CUDA(c, cudaGetLastError());
    PB_CUDA(c, cudaMemsetAsync(c->d_bad.p, 0, sizeof(int), c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));     // the caller's (possibly pinned) buffer has been read
    c->have_tips = true;
    c->have_masks = false;
    c->tip_bytes_stale = false;
    c->pack_dirty = true;
    c->tips_checked = false;
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

int pb_set_site_weights(pb_ctx* c, const double* w)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    c->fresh = false;
    if (!w) { c->have_site_weights = false; return PB_OK; }
    for (int64_t s = 0; s < c->nsites; s++)
        if (!(w[s] >= 0.0)) return fail(c, PB_ERR_INVALID_ARG, "site weights must be >= 0");
    PB_CUDA(c, c->d_site_weights.ensure((size_t)c->spad));
    PB_CUDA(c, cudaMemcpyAsync(c->d_site_weights.p, w, sizeof(double) * c->nsites, cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->have_site_weights = true;
    return PB_OK;
}

int pb_set_tips_text(pb_ctx* c, const char* text, int64_t ld, int alphabet)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    const int width = alphabet == PB_ALPHABET_CODON ? 3 : 1;
    if (!text || ld < c->nsites * width) return fail(c, PB_ERR_INVALID_ARG, "text: null pointer or ld too small");
    const int expect = alphabet == PB_ALPHABET_DNA ? 4 : alphabet == PB_ALPHABET_AMINO_ACID ? 20 : alphabet == PB_ALPHABET_CODON ? 61 : -1;
    if (expect < 0) return fail(c, PB_ERR_INVALID_ARG, "unknown alphabet");
    if (expect != c->n) return fail(c, PB_ERR_INVALID_ARG, "alphabet size does not match nstates");
    unsigned char lut[256 + 64];
    memset(lut, 255, sizeof(lut));
    if (alphabet == PB_ALPHABET_DNA) {                       // lib/nucleotide.ml:8-13 (either case)
        const char* up = "ACGT"; const char* lo = "acgt";
        for (int i = 0; i < 4; i++) { lut[(unsigned char)up[i]] = i; lut[(unsigned char)lo[i]] = i; }
    } else if (alphabet == PB_ALPHABET_AMINO_ACID) {         // lib/amino_acid.ml:4 (upper case only)
        const char* aa = "ACDEFGHIKLMNPQRSTVWY";
        for (int i = 0; i < 20; i++) lut[(unsigned char)aa[i]] = i;
    } else {                                                 // lib/codon.ml:51-58,73: TCAG order, universal code, stops removed
        const char* nuc = "TCAG";
        for (int i = 0; i < 4; i++) lut[(unsigned char)nuc[i]] = i;
        const char* code = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
        int k = 0;
        for (int t = 0; t < 64; t++) lut[256 + t] = (code[t] == '*') ? 255 : (unsigned char)k++;
    }
    int rc = alloc_tips(c);
    if (rc != PB_OK) return rc;
    PB_CUDA(c, c->d_lut.ensure(sizeof(lut)));
    PB_CUDA(c, c->d_text.ensure((size_t)c->ntaxa * c->nsites * width));
    PB_CUDA(c, cudaMemcpyAsync(c->d_lut.p, lut, sizeof(lut), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpy2DAsync(c->d_text.p, c->nsites * width, text, ld, c->nsites * width, c->ntaxa,
                                 cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemsetAsync(c->d_bad.p, 0, sizeof(int), c->stream));
    dim3 grid(c->ntaxa, (unsigned)std::min<int64_t>((c->nsites + 255) / 256, 1024));
    pbk::encode_text<<<grid, 256, 0, c->stream>>>(c->d_text.p, c->nsites * width, width, c->d_lut.p, c->d_tips.p, c->spad,
                                                  c->nsites, c->d_bad.p);
    c->launches++;
    PB_CUDA(c, cudaGetLastError());
    int bad = 0;
    PB_CUDA(c, cudaMemcpyAsync(&bad, c->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (bad) {
        PB_CUDA(c, cudaMemsetAsync(c->d_bad.p, 0, sizeof(int), c->stream));
        c->have_tips = false;
        return fail(c, PB_ERR_INVALID_ARG, "alignment text has a character outside the alphabet (or a stop codon)");
    }
    c->have_tips = true;
    c->have_masks = false;
    c->tip_bytes_stale = false;
    c->pack_dirty = true;
    c->tips_checked = true;
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

int pb_set_tip_masks(pb_ctx* c, const uint64_t* masks, int64_t ld)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    if (!masks || ld < c->nsites) return fail(c, PB_ERR_INVALID_ARG, "masks: null pointer or ld < nsites");
    const size_t count = (size_t)c->ntaxa * c->spad;
    const bool fresh = c->d_masks.count < count || !c->d_masks.p;
    PB_CUDA(c, c->d_masks.ensure(count));
    if (fresh) PB_CUDA(c, cudaMemsetAsync(c->d_masks.p, 0, count * sizeof(uint64_t), c->stream));
    PB_CUDA(c, cudaMemcpy2DAsync(c->d_masks.p, c->spad * sizeof(uint64_t), masks, ld * sizeof(uint64_t),
                                 c->nsites * sizeof(uint64_t), c->ntaxa, cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemsetAsync(c->d_bad.p, 0, sizeof(int), c->stream));
    c->mask_radix = 0;
    if (c->n == 4) {
        // which of the 16 possible state sets occur: they are the alphabet of the site-resident form (fused4t)
        PB_CUDA(c, c->d_mask_present.ensure(1));
        PB_CUDA(c, c->d_mask_code.ensure(16));
        PB_CUDA(c, c->d_mask_set.ensure(16));
        PB_CUDA(c, cudaMemsetAsync(c->d_mask_present.p, 0, sizeof(unsigned), c->stream));
        pbk::mask_alphabet<<<c->sm_count * 8, 256, 0, c->stream>>>(c->d_masks.p, c->spad, c->nsites, c->ntaxa, c->d_mask_present.p);
        c->launches++;
        unsigned present = 0;
        PB_CUDA(c, cudaMemcpyAsync(&present, c->d_mask_present.p, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
        PB_CUDA(c, cudaStreamSynchronize(c->stream));
        present |= 1u;                                      // padding columns hold the empty set: always code 0
        for (int m = 0; m < 16; m++) {
            c->mask_code[m] = 0;
            if ((present >> m) & 1u) { c->mask_code[m] = (unsigned char)c->mask_radix; c->mask_set[c->mask_radix++] = (unsigned char)m; }
        }
        PB_CUDA(c, cudaMemcpyAsync(c->d_mask_code.p, c->mask_code, 16, cudaMemcpyHostToDevice, c->stream));
        PB_CUDA(c, cudaMemcpyAsync(c->d_mask_set.p, c->mask_set, 16, cudaMemcpyHostToDevice, c->stream));
    }
    PB_CUDA(c, cudaStreamSynchronize(c->stream));     // the caller's (possibly pinned) buffer has been read
    c->fresh = false;
    c->pack_dirty = true;
    c->have_masks = true;
    c->have_tips = true;
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

int pb_set_root_frequencies(pb_ctx* c, const double* pi)
{
    PB_ENTER(c);
    if (!pi) return fail(c, PB_ERR_INVALID_ARG, "pi is null");
    PB_CUDA(c, c->d_pi.ensure(c->n));
    PB_CUDA(c, cudaMemcpyAsync(c->d_pi.p, pi, c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->have_pi = true;
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

int pb_set_categories(pb_ctx* c, const double* rates, const double* weights)
{
    PB_ENTER(c);
    if (!rates || !weights) return fail(c, PB_ERR_INVALID_ARG, "categories: null pointer");
    for (int k = 0; k < c->ncat; k++)
        if (!(rates[k] >= 0.0) || !(weights[k] >= 0.0)) return fail(c, PB_ERR_INVALID_ARG, "category rates and weights must be >= 0");
    c->h_rates.assign(rates, rates + c->ncat);
    c->h_weights.assign(weights, weights + c->ncat);
    PB_CUDA(c, cudaMemcpyAsync(c->d_rates.p, c->h_rates.data(), c->ncat * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_weights.p, c->h_weights.data(), c->ncat * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->p_dirty = true;
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

int pb_set_eigensystem(pb_ctx* c, const double* U, const double* lambda, const double* Uinv)
{
    PB_ENTER(c);
    if (c->psrc != PSRC_EIGEN) PB_RESHAPES(c);      // (new VALUES of an eigensystem keep the captured launches valid)
    if (!U || !lambda || !Uinv) return fail(c, PB_ERR_INVALID_ARG, "eigensystem: null pointer");
    const int n = c->n;
    std::vector<double> ur((size_t)n * n), uir((size_t)n * n);
    to_row_major(n, 1, U, ur.data());
    to_row_major(n, 1, Uinv, uir.data());
    PB_CUDA(c, c->d_U.ensure((size_t)n * n));
    PB_CUDA(c, c->d_Uinv.ensure((size_t)n * n));
    PB_CUDA(c, c->d_lambda.ensure(n));
    PB_CUDA(c, cudaMemcpyAsync(c->d_U.p, ur.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_Uinv.p, uir.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->d_lambda.p, lambda, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->psrc = PSRC_EIGEN;
    c->p_dirty = true;
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

// Host-only introspection of what compile_tree / build_table_program produce for a tree (4 states): no device is
// touched, nothing is evaluated.  The tests interpret the programs on the CPU against the oracle.
int pb_debug_fused_program(int nnodes, int root, const int32_t* child_ptr, const int32_t* child_idx, const int32_t* tip_row,
                           int ntaxa, int max_leaves, int32_t* counts, int32_t* ops, int32_t* ops_tab, int cap_ops,
                           int32_t* tip_order, int32_t* inner_nodes, int32_t* leaf_nodes, int32_t* clades, int cap_clades)
{
    if (!child_ptr || !child_idx || !tip_row || !counts || !ops || !ops_tab || !tip_order || !inner_nodes || !leaf_nodes ||
        !clades || nnodes < 2 || root < 0 || root >= nnodes || ntaxa < 1 || child_ptr[0] != 0 || child_ptr[nnodes] != nnodes - 1)
        return fail(nullptr, PB_ERR_INVALID_ARG, "pb_debug_fused_program: bad arguments");
    std::unique_ptr<pb_ctx> c(new pb_ctx());
    c->plan_only = true;
    c->n = 4;
    c->ncat = 1;
    c->ntaxa = ntaxa;
    c->nsites = 1;
    c->spad = 1024;
    c->nnodes = nnodes;
    c->root = root;
    c->child_ptr.assign(child_ptr, child_ptr + nnodes + 1);
    c->child_idx.assign(child_idx, child_idx + nnodes - 1);
    c->tip_row.assign(tip_row, tip_row + nnodes);
    c->parent.assign(nnodes, -1);
    for (int v = 0; v < nnodes; v++)
        for (int k = child_ptr[v]; k < child_ptr[v + 1]; k++) {
            const int ch = child_idx[k];
            if (ch < 0 || ch >= nnodes || ch == root || c->parent[ch] != -1)
                return fail(nullptr, PB_ERR_INVALID_ARG, "pb_debug_fused_program: bad child index");
            c->parent[ch] = v;
        }
    int rc = compile_tree(c.get());
    if (rc == PB_OK) rc = build_table_program(c.get(), max_leaves & 0xff, (max_leaves & 0x100) != 0);
    if (rc != PB_OK) return fail(nullptr, rc, c->err);
    if ((int)c->fprog_c.size() > cap_ops || (int)c->clades.size() > cap_clades)
        return fail(nullptr, PB_ERR_INVALID_ARG, "pb_debug_fused_program: output buffers too small");
    static_assert(sizeof(FusedOp2) == 4 * sizeof(int32_t), "one instruction is four 32-bit words");
    std::memcpy(ops, c->fprog_c.data(), c->fprog_c.size() * sizeof(FusedOp2));
    std::memcpy(ops_tab, c->fprog_tab.data(), c->fprog_tab.size() * sizeof(FusedOp2));
    std::copy(c->tip_order.begin(), c->tip_order.end(), tip_order);
    std::copy(c->inner_nodes.begin(), c->inner_nodes.end(), inner_nodes);
    std::copy(c->leaf_nodes.begin(), c->leaf_nodes.end(), leaf_nodes);
    for (size_t i = 0; i < c->clades.size(); i++) {
        const pbk::CladeTab& t = c->clades[i];
        const int32_t row[6] = {t.op0, t.op1, t.sh0, t.k, t.ent0, t.pre};
        std::copy(row, row + 6, clades + 6 * i);
    }
    const int32_t cnt[8] = {(int32_t)c->fprog_c.size(), (int32_t)c->fprog_tab.size(), (int32_t)c->tip_order.size(), c->fwords,
                            std::max(c->fstack, 1), (int32_t)c->inner_nodes.size(), (int32_t)c->leaf_nodes.size(),
                            (int32_t)c->clades.size()};
    std::copy(cnt, cnt + 8, counts);
    return PB_OK;
}

// The same for the program over the REDUCED tree (kernel fused4t; build_reduced_program): its instructions, the
// clades' own instruction ranges (what fills their tables), the packed-tip slot order and the matrix lists.
int pb_debug_reduced_program(int nnodes, int root, const int32_t* child_ptr, const int32_t* child_idx, const int32_t* tip_row,
                             int ntaxa, int max_leaves, int32_t* counts, int32_t* ops, int cap_ops, int32_t* clade_ops,
                             int cap_clade_ops, int32_t* slot_rows, int32_t* slot_xor_rows, int cap_slots, int32_t* mat_nodes,
                             int32_t* leaf_nodes, int32_t* clades, int cap_clades)
{
    if (!slot_xor_rows) return fail(nullptr, PB_ERR_INVALID_ARG, "pb_debug_reduced_program: bad arguments");
    if (!child_ptr || !child_idx || !tip_row || !counts || !ops || !clade_ops || !slot_rows || !mat_nodes || !leaf_nodes ||
        !clades || nnodes < 2 || root < 0 || root >= nnodes || ntaxa < 1 || child_ptr[0] != 0 || child_ptr[nnodes] != nnodes - 1)
        return fail(nullptr, PB_ERR_INVALID_ARG, "pb_debug_reduced_program: bad arguments");
    std::unique_ptr<pb_ctx> c(new pb_ctx());
    c->plan_only = true;
    c->n = 4;
    c->ncat = 1;
    c->ntaxa = ntaxa;
    c->nsites = 1;
    c->spad = 1024;
    c->nnodes = nnodes;
    c->root = root;
    c->child_ptr.assign(child_ptr, child_ptr + nnodes + 1);
    c->child_idx.assign(child_idx, child_idx + nnodes - 1);
    c->tip_row.assign(tip_row, tip_row + nnodes);
    c->parent.assign(nnodes, -1);
    for (int v = 0; v < nnodes; v++)
        for (int k = child_ptr[v]; k < child_ptr[v + 1]; k++) {
            const int ch = child_idx[k];
            if (ch < 0 || ch >= nnodes || ch == root || c->parent[ch] != -1)
                return fail(nullptr, PB_ERR_INVALID_ARG, "pb_debug_reduced_program: bad child index");
            c->parent[ch] = v;
        }
    const int rc = build_reduced_program(c.get(), max_leaves);
    if (rc != PB_OK) return fail(nullptr, rc, c->err);
    if ((int)c->rprog.size() > cap_ops || (int)c->cprog.size() > cap_clade_ops || (int)c->rorder.size() > cap_slots ||
        (int)c->rclades.size() > cap_clades)
        return fail(nullptr, PB_ERR_INVALID_ARG, "pb_debug_reduced_program: output buffers too small");
    std::memcpy(ops, c->rprog.data(), c->rprog.size() * sizeof(FusedOp2));
    std::memcpy(clade_ops, c->cprog.data(), c->cprog.size() * sizeof(FusedOp2));
    std::copy(c->rorder.begin(), c->rorder.end(), slot_rows);
    std::copy(c->rxorder.begin(), c->rxorder.end(), slot_xor_rows);
    std::copy(c->rmat_nodes.begin(), c->rmat_nodes.end(), mat_nodes);
    std::copy(c->rleaf_nodes.begin(), c->rleaf_nodes.end(), leaf_nodes);
    for (size_t i = 0; i < c->rclades.size(); i++) {
        const pbk::CladeTab& t = c->rclades[i];
        const int32_t row[6] = {t.op0, t.op1, t.xor_coded, t.k, t.ent0, t.pre};
        std::copy(row, row + 6, clades + 6 * i);
    }
    const int32_t cnt[8] = {(int32_t)c->rprog.size(), (int32_t)c->cprog.size(), (int32_t)c->rorder.size(), c->rwords,
                            std::max(c->rstack, 1), (int32_t)c->rmat_nodes.size(), (int32_t)c->rleaf_nodes.size(),
                            (int32_t)c->rclades.size()};
    std::copy(cnt, cnt + 8, counts);
    return PB_OK;
}

// Host-only view of the dense-alphabet program over the tree whose cherries are tables (kernel fused_dmma3;
// build_dense_reduced_program), for tests of the planner.
int pb_debug_dense_reduced_program(int nnodes, int root, const int32_t* child_ptr, const int32_t* child_idx, const int32_t* tip_row,
                                   int ntaxa, int nstates, int32_t* counts, int32_t* ops, int cap_ops, int32_t* mlist, int cap_mlist,
                                   int32_t* tabs, int32_t* tab_nodes, int cap_tabs, int32_t* inner_nodes, int32_t* leaf_nodes)
{
    if (!child_ptr || !child_idx || !tip_row || !counts || !ops || !mlist || !tabs || !tab_nodes || !inner_nodes || !leaf_nodes ||
        nnodes < 2 || root < 0 || root >= nnodes || ntaxa < 1 || child_ptr[0] != 0 || child_ptr[nnodes] != nnodes - 1)
        return fail(nullptr, PB_ERR_INVALID_ARG, "pb_debug_dense_reduced_program: bad arguments");
    std::unique_ptr<pb_ctx> c(new pb_ctx());
    c->plan_only = true;
    c->n = nstates & 0xff;
    c->fused_dmma_triple = (nstates & 0x100) != 0;          // (option "fused_dmma_triple")
    c->ncat = 1;
    c->ntaxa = ntaxa;
    c->nsites = 1;
    c->spad = 1024;
    c->nnodes = nnodes;
    c->root = root;
    c->child_ptr.assign(child_ptr, child_ptr + nnodes + 1);
    c->child_idx.assign(child_idx, child_idx + nnodes - 1);
    c->tip_row.assign(tip_row, tip_row + nnodes);
    c->parent.assign(nnodes, -1);
    for (int v = 0; v < nnodes; v++)
        for (int k = child_ptr[v]; k < child_ptr[v + 1]; k++) {
            const int ch = child_idx[k];
            if (ch < 0 || ch >= nnodes || ch == root || c->parent[ch] != -1)
                return fail(nullptr, PB_ERR_INVALID_ARG, "pb_debug_dense_reduced_program: bad child index");
            c->parent[ch] = v;
        }
    int rc = compile_tree(c.get());                           // (the compact matrix indices)
    if (rc == PB_OK) rc = build_dense_reduced_program(c.get());
    if (rc != PB_OK) return fail(nullptr, rc, c->err);
    if ((int)c->dprog_r.size() > cap_ops || (int)c->mlist_r.size() > cap_mlist || (int)c->dtabs.size() > cap_tabs)
        return fail(nullptr, PB_ERR_INVALID_ARG, "pb_debug_dense_reduced_program: output buffers too small");
    std::memcpy(ops, c->dprog_r.data(), c->dprog_r.size() * sizeof(FusedOp2));
    std::copy(c->mlist_r.begin(), c->mlist_r.end(), mlist);
    std::memcpy(tabs, c->dtabs.data(), c->dtabs.size() * sizeof(int4));
    std::memcpy(tab_nodes, c->dtab_nodes.data(), c->dtab_nodes.size() * sizeof(pbk::DenseTabNodes));
    std::copy(c->inner_nodes.begin(), c->inner_nodes.end(), inner_nodes);
    std::copy(c->leaf_nodes.begin(), c->leaf_nodes.end(), leaf_nodes);
    const int32_t cnt[8] = {(int32_t)c->dprog_r.size(), (int32_t)c->mlist_r.size(), (int32_t)c->dtabs.size(), c->dstack_r,
                            (int32_t)c->inner_nodes.size(), (int32_t)c->leaf_nodes.size(), c->dstack, (int32_t)c->mlist.size()};
    std::copy(cnt, cnt + 8, counts);
    return PB_OK;
}

// Page-lock a caller-owned host buffer (an OCaml Bigarray payload lives outside the OCaml heap and does not move)
// so that the chunked copies of pb_loglik_host / pb_loglik_host_packed2 overlap with the pruning.
int pb_pin_host_memory(void* p, size_t bytes)
{
    if (!p || bytes == 0) return fail(nullptr, PB_ERR_INVALID_ARG, "pb_pin_host_memory: null pointer or empty range");
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(nullptr, e, "cudaHostRegister"); }
    return PB_OK;
}

int pb_unpin_host_memory(void* p)
{
    if (!p) return fail(nullptr, PB_ERR_INVALID_ARG, "pb_unpin_host_memory: null pointer");
    const cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(nullptr, e, "cudaHostUnregister"); }
    return PB_OK;
}

int pb_host_reversible_eigensystem(int n, const double* Q, const double* pi, double* U, double* lambda, double* Uinv)
{
    if (n < 1 || n > 64 || !Q || !pi || !U || !lambda || !Uinv) return PB_ERR_INVALID_ARG;
    std::string why;
    return pbh::reversible_eigensystem(n, Q, pi, U, lambda, Uinv, &why) ? PB_OK : PB_ERR_NUMERIC;
}

int pb_set_rate_matrix(pb_ctx* c, const double* Q, const double* pi)
{
    PB_ENTER(c);
    if (!Q || !pi) return fail(c, PB_ERR_INVALID_ARG, "rate matrix: null pointer");
    const int n = c->n;
    std::vector<double> U((size_t)n * n), Ui((size_t)n * n), lam(n);
    std::string why;
    if (!pbh::reversible_eigensystem(n, Q, pi, U.data(), lam.data(), Ui.data(), &why))
        return fail(c, PB_ERR_NUMERIC, "pb_set_rate_matrix: " + why);
    return pb_set_eigensystem(c, U.data(), lam.data(), Ui.data());
}

int pb_set_rate_matrix_general(pb_ctx* c, const double* Q)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    if (!Q) return fail(c, PB_ERR_INVALID_ARG, "rate matrix: null pointer");
    const int n = c->n;
    std::vector<double> qr((size_t)n * n);
    to_row_major(n, 1, Q, qr.data());
    PB_CUDA(c, c->d_Q.ensure((size_t)n * n));
    PB_CUDA(c, cudaMemcpyAsync(c->d_Q.p, qr.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->psrc = PSRC_EXPM;
    c->p_dirty = true;
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

int pb_set_pmatrices(pb_ctx* c, const double* P)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    if (!c->have_tree) return fail(c, PB_ERR_STATE, "set the tree first");
    if (!P) return fail(c, PB_ERR_INVALID_ARG, "P is null");
    const int n = c->n;
    const size_t count = (size_t)c->ncat * c->nnodes;
    std::vector<double> pr(count * n * n);
    to_row_major(n, count, P, pr.data());
    PB_CUDA(c, c->d_P.ensure(count * n * n));
    PB_CUDA(c, cudaMemcpyAsync(c->d_P.p, pr.data(), sizeof(double) * pr.size(), cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->psrc = PSRC_DIRECT;
    c->p_dirty = false;
    c->all_dirty = true;
    c->fresh = false;
    return PB_OK;
}

// Option "graph": replay (capturing it first when needed) one evaluation of the reduced 4-state program as a CUDA graph.
// Only after a plain evaluation with the same configuration has allocated, planned and packed everything
// (warm_epoch), and only while nothing has reshaped the launches since the capture (graph_epoch).
static int enqueue_graph(pb_ctx* c, bool* done)
{
    *done = false;
    bool reduced = false;
    int rc = want_reduced(c, &reduced);
    if (rc != PB_OK) return rc;
    if (!reduced || path_of(c) != PATH_FUSED4 || c->have_masks || c->pack_dirty || c->psrc == PSRC_NONE ||
        c->psrc == PSRC_DIRECT || !c->have_blen || c->warm_epoch != c->config_epoch)
        return PB_OK;
    cudaEvent_t& bank_free = g_const_event[c->device & 63];
    if (c->graph_epoch != c->config_epoch || !c->graph_exec) {
        if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
        const int64_t before = c->launches;
        cudaGraph_t graph = nullptr;
        PB_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
        c->capturing = true;
        c->p_dirty = true;                                    // the captured evaluation always rebuilds P(t)
        rc = build_pmatrices(c);
        if (rc == PB_OK) rc = run_pruning(c);
        if (rc == PB_OK) rc = run_root_reduction(c);
        if (rc == PB_OK) {
            cudaMemcpyAsync(c->h_total, c->d_total.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
            cudaMemcpyAsync(c->h_total + 1, c->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
        }
        c->capturing = false;
        const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        if (rc != PB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(c, e, "cudaStreamEndCapture"); }
        const cudaError_t e2 = cudaGraphInstantiate(&c->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e2 != cudaSuccess) { c->graph_exec = nullptr; return cuda_fail(c, e2, "cudaGraphInstantiate"); }
        c->graph_launches = c->launches - before;
        c->launches = before;
        c->graph_epoch = c->config_epoch;
    }
    {
        std::lock_guard<std::recursive_mutex> lock(g_const_mutex);      // the constant bank is ours for the whole graph
        if (bank_free) PB_CUDA(c, cudaStreamWaitEvent(c->stream, bank_free, 0));
        PB_CUDA(c, cudaGraphLaunch(c->graph_exec, c->stream));
        if (bank_free) PB_CUDA(c, cudaEventRecord(bank_free, c->stream));
        g_prog_owner[c->device & 63] = c;                     // (the graph uploads its program every time)
        g_pinner_owner[c->device & 63] = nullptr;             // (... and its matrices; what the bank holds afterwards is the last group's)
        g_prog_version[c->device & 63] = (c->fprog_version * 64 + (uint64_t)c->r_leaves * 4 + 2) * 32;
    }
    c->launches += c->graph_launches;
    c->graph_replays++;
    c->p_dirty = false;
    c->evaluated = true;
    c->fresh = true;
    *done = true;
    return PB_OK;
}

int pb_loglik_enqueue(pb_ctx* c)
{
    PB_ENTER(c);
    if (!c->have_tree) return fail(c, PB_ERR_STATE, "tree not set");
    if (!c->have_tips) return fail(c, PB_ERR_STATE, "tip states not set");
    if (!c->have_pi) return fail(c, PB_ERR_STATE, "root frequencies not set");
    if (c->use_graph && !c->timing && !(c->flags & (PB_FLAG_KEEP_CLV | PB_FLAG_LEVEL_KERNELS))) {
        bool done = false;
        const int rc = enqueue_graph(c, &done);
        if (rc != PB_OK) return rc;
        if (done) return PB_OK;
    }
    if (c->timing) { c->ev = c->ev_ring[c->timed_evals++ % pb_ctx::TIMING_RING]; c->kev_n = 0; PB_CUDA(c, cudaEventRecord(c->ev[0], c->stream)); }
    int rc = build_pmatrices(c);
    if (rc != PB_OK) return rc;
    if (c->timing) PB_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    if ((rc = run_pruning(c)) != PB_OK) return rc;
    if (c->timing) PB_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    if ((rc = run_root_reduction(c)) != PB_OK) return rc;
    if (c->timing) PB_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->h_total, c->d_total.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->h_total + 1, c->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    c->evaluated = true;
    c->fresh = true;
    c->warm_epoch = c->config_epoch;         // everything is allocated, planned and packed for this configuration
    return PB_OK;
}

int pb_loglik_fetch(pb_ctx* c, double* total, double* per_site)
{
    PB_ENTER(c);
    if (!c->evaluated) return fail(c, PB_ERR_STATE, "nothing enqueued");
    if (per_site)
        PB_CUDA(c, cudaMemcpyAsync(per_site, c->d_per_site.p, sizeof(double) * c->nsites, cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (*reinterpret_cast<const int*>(c->h_total + 1) != 0)
        return fail(c, PB_ERR_INVALID_ARG, "tip state out of range (>= nstates) in the alignment");
    if (total) *total = *c->h_total;
    return PB_OK;
}

// Column bounds of the chunks of a pipelined host evaluation over `spad` (a multiple of 1024) site columns.
// round > 0: the sites one ROUND of the persistent pruning blocks covers.  A launch costs whole rounds, so chunks are
// whole rounds (cut down to the 1024 sites = 256 packed bytes the copies are aligned to), and what the pipeline cannot
// hide is the work behind the LAST copy, so the schedule is laid from the end: a ramp of short last chunks (t, t, 2t,
// 4t, ... rounds, t = tail_sites in rounds, at least one; tail_sites < 0: no ramp) while shorter than a regular chunk,
// regular chunks of about spad / chunks before it, what is left over first.  round == 0: `chunks` equal chunks.
// tail_sites == -2: the ramp-up schedule of a compute-bound pipeline (below).  At most 64 chunks in every case.
static std::vector<int64_t> host_chunk_bounds(int64_t spad, int chunks, int64_t tail_sites, int64_t round)
{
    std::vector<int64_t> bounds;
    chunks = std::max(1, std::min(chunks, 64));
    if (round > 0 && tail_sites == -2) {
        // The schedule of a COMPUTE-bound pipeline (the dense alphabets: the copy takes a fraction of the pruning time):
        // what cannot be hidden is the FIRST copy, so the chunks grow from one round — 1, 2, 4, ... rounds; every copy
        // is done long before the chunks ahead of it are pruned.
        auto rounds = [&](int64_t r) { return std::max<int64_t>(1024, r * round / 1024 * 1024); };
        int64_t lo = 0;
        bounds.push_back(0);
        for (int64_t t = 1; spad - lo > rounds(t) + rounds(t) / 2 && bounds.size() < 64; t *= 2) bounds.push_back(lo += rounds(t));
        bounds.push_back(spad);
        return bounds;
    }
    if (round > 0) {
        auto rounds = [&](int64_t r) { return std::max<int64_t>(1024, r * round / 1024 * 1024); };
        const int64_t per = std::max<int64_t>(1, (spad / chunks + round / 2) / round);
        int64_t hi = spad;
        bounds.push_back(hi);
        if (tail_sites >= 0) {
            int64_t t = std::max<int64_t>(1, tail_sites / round);
            for (int step = 0; t < per && hi > rounds(t) + rounds(1) / 2 && bounds.size() < 60; step++) {
                bounds.push_back(hi -= rounds(t));
                if (step >= 1) t *= 2;
            }
        }
        while (hi > rounds(per) + rounds(1) / 2 && bounds.size() < 64) bounds.push_back(hi -= rounds(per));
        bounds.push_back(0);
        std::reverse(bounds.begin(), bounds.end());
    } else {
        const int64_t chunk = round_up((spad + chunks - 1) / chunks, 1024);
        for (int64_t s0 = 0; s0 < spad; s0 += chunk) bounds.push_back(s0);
        bounds.push_back(spad);
    }
    return bounds;
}

int pb_debug_host_schedule(int64_t spad, int host_chunks, int64_t host_tail_sites, int64_t round_sites, int64_t* bounds,
                           int cap, int* nbounds)
{
    if (spad < 1024 || spad % 1024 || host_chunks < 1 || round_sites < 0 || !bounds || !nbounds) return PB_ERR_INVALID_ARG;
    const std::vector<int64_t> b = host_chunk_bounds(spad, host_chunks, host_tail_sites, round_sites);
    if ((int)b.size() > cap) return PB_ERR_INVALID_ARG;
    std::copy(b.begin(), b.end(), bounds);
    *nbounds = (int)b.size();
    return PB_OK;
}

// The pipelined host evaluation of the dense alphabets (20 / 61 states, kernel fused_dmma3): the byte-per-site
// alignment crosses PCIe in column chunks (whole rounds of the persistent CTAs, host_chunk_bounds) on the copy stream
// while the chunks before it are pruned and reduced; matrix panels and cherry tables are built behind the first copy.
// One compute stream: the kernel's spill area (d_gstack) belongs to one launch at a time.
static int loglik_host_dense(pb_ctx* c, const uint8_t* states, int64_t ld, double* total, double* per_site, bool fetch)
{
    int rc = alloc_tips(c);
    if (rc != PB_OK) return rc;
    if (!c->copy_stream) PB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!c->ev_start) PB_CUDA(c, cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
    if (c->timing) { c->ev = c->ev_ring[c->timed_evals++ % pb_ctx::TIMING_RING]; c->kev_n = 0; PB_CUDA(c, cudaEventRecord(c->ev[0], c->stream)); }
    if ((rc = build_pmatrices(c)) != PB_OK) return rc;
    if (c->timing) PB_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    const int64_t nblocks = (c->nsites + 255) / 256;
    PB_CUDA(c, c->d_site_cat.ensure((size_t)c->ncat * c->spad));
    PB_CUDA(c, c->d_site_exp.ensure((size_t)c->ncat * c->spad));
    PB_CUDA(c, c->d_per_site.ensure((size_t)c->spad));
    PB_CUDA(c, c->d_partials.ensure((size_t)nblocks));
    PB_CUDA(c, c->d_total.ensure(1));
    PB_CUDA(c, cudaMemsetAsync(c->d_bad.p, 0, sizeof(int), c->stream));
    PB_CUDA(c, cudaEventRecord(c->ev_start, c->stream));          // earlier work may still read d_tips
    PB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_start, 0));
    size_t tr = 0;                                                // option "host_trace": next timing event
    auto trace = [&](cudaStream_t st) -> cudaError_t {
        if (!c->host_trace) return cudaSuccess;
        if (tr == c->trace_events.size()) {
            cudaEvent_t e;
            cudaError_t err = cudaEventCreate(&e);
            if (err != cudaSuccess) return err;
            c->trace_events.push_back(e);
        }
        return cudaEventRecord(c->trace_events[tr++], st);
    };
    c->trace_sites.clear();
    PB_CUDA(c, trace(c->stream));
    const std::vector<int64_t> bounds = host_chunk_bounds(c->spad, c->host_chunks, c->host_chunk_align ? -2 : c->host_tail_sites,
                                                          c->host_chunk_align ? pbk::fused_dmma3_round_sites(c->n, c->sm_count) : 0);
    const int nchunks = (int)bounds.size() - 1;
    while ((int)c->chunk_events.size() < nchunks) {
        cudaEvent_t e;
        PB_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->chunk_events.push_back(e);
    }
    if ((rc = launch_fused_dmma_path(c, 0, 0, true)) != PB_OK) return rc;      // panels and tables, behind the first copy
    for (int k = 0; k < nchunks; k++) {
        const int64_t s0 = bounds[k], s1 = bounds[k + 1];
        const int64_t w = std::min(s1, c->nsites) - s0;
        if (w > 0)
            PB_CUDA(c, cudaMemcpy2DAsync(c->d_tips.p + s0, c->spad, states + s0, ld, w, c->ntaxa, cudaMemcpyHostToDevice,
                                         c->copy_stream));
        PB_CUDA(c, cudaEventRecord(c->chunk_events[k], c->copy_stream));
        PB_CUDA(c, trace(c->copy_stream));
        c->trace_sites.push_back(s0);
        c->trace_sites.push_back(s1 - s0);
        PB_CUDA(c, cudaStreamWaitEvent(c->stream, c->chunk_events[k], 0));
        if ((rc = launch_fused_dmma_path(c, s0, s1, false)) != PB_OK) return rc;
        const int64_t b0 = s0 / 256, b1 = std::min<int64_t>(s1 / 256, nblocks);
        if (b1 > b0) {
            pbk::root_combine<<<(unsigned)(b1 - b0), 256, 0, c->stream>>>(
                c->d_site_cat.p, site_exp_of(c), c->ncat, c->d_weights.p, c->have_site_weights ? c->d_site_weights.p : nullptr,
                c->nsites, c->spad, b0, c->d_per_site.p, c->d_partials.p);
            c->launches++;
        }
        PB_CUDA(c, trace(c->stream));
    }
    // (the kernel clamps a state outside the alphabet; the flag makes the call fail instead of returning that value)
    pbk::validate_tips<<<c->sm_count * 4, 256, 0, c->stream>>>(c->d_tips.p, (int64_t)c->ntaxa * c->spad, c->n, c->d_bad.p);
    c->launches++;
    if (c->timing) PB_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    pbk::final_sum<<<1, 256, 0, c->stream>>>(c->d_partials.p, nblocks, c->d_total.p);
    c->launches++;
    if (c->timing) PB_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    PB_CUDA(c, cudaGetLastError());
    PB_CUDA(c, cudaMemcpyAsync(c->h_total, c->d_total.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->h_total + 1, c->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(c, trace(c->stream));
    c->have_tips = true;
    c->have_masks = false;
    c->pack_dirty = true;
    c->tips_checked = true;
    c->all_dirty = true;
    c->evaluated = true;
    c->fresh = true;
    c->tip_bytes_stale = false;
    return fetch ? pb_loglik_fetch(c, total, per_site) : PB_OK;
}

static int loglik_host_impl(pb_ctx* c, const uint8_t* states, int64_t ld, bool packed2, double* total, double* per_site,
                            bool fetch = true)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    if (packed2 && c->n != 4) return fail(c, PB_ERR_STATE, "2-bit packed tips are for 4-state alphabets");
    if (!states || ld < (packed2 ? (c->nsites + 3) / 4 : c->nsites))
        return fail(c, PB_ERR_INVALID_ARG, packed2 ? "packed tips: null pointer or ld < ceil(nsites / 4)" : "tips: null pointer or ld < nsites");
    if (!c->have_tree) return fail(c, PB_ERR_STATE, "tree not set");
    if (!c->have_pi) return fail(c, PB_ERR_STATE, "root frequencies not set");
    bool reduced = false;                                         // which 4-state program runs (and orders the packed tips)
    {
        const bool had_masks = c->have_masks;
        c->have_masks = false;                                    // this call replaces whatever the context held
        const int rc0 = want_reduced(c, &reduced);
        c->have_masks = had_masks;
        if (rc0 != PB_OK) return rc0;
    }
    {
        const bool had_masks = c->have_masks;
        c->have_masks = false;                                    // (the path of plain tips: this call replaces the masks)
        const bool dense = path_of(c) == PATH_FUSED_DMMA && !packed2 && c->fused_dmma_version == 3 && c->host_chunks > 1 &&
                           c->spad >= 8192;
        if (dense) return loglik_host_dense(c, states, ld, total, per_site, fetch);
        c->have_masks = had_masks;
    }
    if (path_of(c) != PATH_FUSED4 || c->host_chunks == 1 || c->spad < 8192) {
        int rc = packed2 ? pb_set_tips_packed2(c, states, ld) : pb_set_tips(c, states, ld);
        if (rc != PB_OK) return rc;
        if ((rc = pb_loglik_enqueue(c)) != PB_OK) return rc;
        return fetch ? pb_loglik_fetch(c, total, per_site) : PB_OK;
    }
    // Pipeline: the alignment crosses PCIe in column chunks on a copy stream while the previous
    // chunk is packed, pruned and root-reduced on the compute stream.
    int rc = alloc_tips(c);
    if (rc != PB_OK) return rc;
    if (packed2 && (rc = alloc_tips2(c)) != PB_OK) return rc;
    if (!c->copy_stream) PB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!c->ev_start) PB_CUDA(c, cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
    while ((int)c->chunk_events.size() < c->host_chunks) {
        cudaEvent_t e;
        PB_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->chunk_events.push_back(e);
    }
    if (c->timing) { c->ev = c->ev_ring[c->timed_evals++ % pb_ctx::TIMING_RING]; c->kev_n = 0; PB_CUDA(c, cudaEventRecord(c->ev[0], c->stream)); }
    if ((rc = build_pmatrices(c)) != PB_OK) return rc;
    if (c->timing) PB_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    const int64_t nblocks = (c->nsites + 255) / 256;
    const int nwords = reduced ? c->rwords : c->fwords;
    const int* d_order = reduced ? c->d_rorder.p : c->d_tip_order.p;
    const int* d_xorder = reduced ? c->d_rxorder.p : nullptr;
    const int nslots = reduced ? (int)c->rorder.size() : (int)c->tip_order.size();
    PB_CUDA(c, c->d_site_cat.ensure((size_t)c->ncat * c->spad));
    PB_CUDA(c, c->d_site_exp.ensure((size_t)c->ncat * c->spad));
    PB_CUDA(c, c->d_packed.ensure((size_t)(nwords + 2) * c->spad));
    PB_CUDA(c, c->d_per_site.ensure((size_t)c->spad));
    PB_CUDA(c, c->d_partials.ensure((size_t)nblocks));
    PB_CUDA(c, c->d_total.ensure(1));
    PB_CUDA(c, cudaMemsetAsync(c->d_bad.p, 0, sizeof(int), c->stream));
    PB_CUDA(c, cudaEventRecord(c->ev_start, c->stream));          // earlier work may still read d_tips
    PB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_start, 0));
    size_t tr = 0;                                                // option "host_trace": next timing event
    auto trace = [&](cudaStream_t st) -> cudaError_t {
        if (!c->host_trace) return cudaSuccess;
        if (tr == c->trace_events.size()) {
            cudaEvent_t e;
            cudaError_t err = cudaEventCreate(&e);
            if (err != cudaSuccess) return err;
            c->trace_events.push_back(e);
        }
        return cudaEventRecord(c->trace_events[tr++], st);
    };
    c->trace_sites.clear();
    PB_CUDA(c, trace(c->stream));
    // 2-bit packed: the first `host_lead_cats` categories (default: all) are pruned chunk by chunk behind the copy in one
    // launch per chunk, the others (if any) run once over all the sites.
    const bool first_cat_only = packed2 && c->ncat > 1 && (reduced || fused_const_ok(c));
    const int lead_cats = std::max(1, std::min(c->host_lead_cats, c->ncat));
    const bool root_per_chunk = !first_cat_only || lead_cats == c->ncat;
    // Column bounds of the chunks.  A launch of the persistent kernel costs whole ROUNDS of its resident blocks (they
    // are dealt over the categories of the launch: a round covers sm_count x blocks per SM / categories blocks of
    // fb x fq sites), so chunks are whole rounds (cut to the 1024 sites = 256 packed bytes the copies are aligned to).
    // What the pipeline cannot hide is the work behind the LAST copy: the schedule is laid from the end — a short
    // last chunk (option "host_tail_sites"; default one round), regular chunks before it, what is left over first.
    int64_t round_sites = 0;
    if (reduced && c->host_chunk_align) {
        int fb = 0, fq = 0;
        fused4t_shape(c, &fb, &fq);
        const int cats_per_launch = fused4t_cats_per_launch(c, first_cat_only ? lead_cats : c->ncat);
        const int64_t blocks_x = std::max<int64_t>(1, ((int64_t)c->sm_count * std::max(c->fused4t_occ, 1) + cats_per_launch - 1) / cats_per_launch);
        round_sites = blocks_x * fb * fq;
    }
    const std::vector<int64_t> bounds = host_chunk_bounds(c->spad, c->host_chunks, c->host_tail_sites, round_sites);
    const int nchunks = (int)bounds.size() - 1;
    auto grow = [&](std::vector<cudaEvent_t>& v) -> cudaError_t {
        while ((int)v.size() < nchunks) {
            cudaEvent_t e;
            cudaError_t err = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            if (err != cudaSuccess) return err;
            v.push_back(e);
        }
        return cudaSuccess;
    };
    PB_CUDA(c, grow(c->chunk_events));
    // Two lanes: the chunks alternate between two streams, each packing, pruning and reducing its chunk in order.
    // Nothing orders chunk k+1 after chunk k, so its blocks take over the SMs as the persistent blocks of chunk k
    // retire — no ramp between launches — and the small pack / reduction kernels run in the gaps.
    const bool side = c->host_side_streams != 0;
    // The lanes' launches only READ the constant bank when one load holds every category they ask for: they neither wait
    // for nor move the bank's event; the bank is ours (mutex) from its preparation until the event is recorded behind
    // the joined lanes.
    const int want_cats = first_cat_only ? lead_cats : c->ncat;
    const bool hold_bank = side && reduced && c->fused_const_keep && fused4t_cats_per_launch(c, want_cats) == want_cats;
    std::unique_lock<std::recursive_mutex> bank_lock(g_const_mutex, std::defer_lock);
    if (hold_bank) bank_lock.lock();
    // Nothing of this depends on the alignment: matrices for the constant bank and the clade tables are built while
    // the first chunk is on its way.
    if (reduced && (rc = launch_fused4t(c, 0, 0, true, 0, want_cats)) != PB_OK) return rc;
    const int nlanes = side ? std::max(2, std::min(c->host_lanes, 4)) : 1;
    cudaStream_t lanes[4] = {c->stream, c->stream, c->stream, c->stream};
    if (side) {
        if (!c->ev_root_done) PB_CUDA(c, cudaEventCreateWithFlags(&c->ev_root_done, cudaEventDisableTiming));
        for (int l = 0; l < nlanes; l++) {
            if (!c->lane_streams[l]) PB_CUDA(c, cudaStreamCreateWithFlags(&c->lane_streams[l], cudaStreamNonBlocking));
            if (!c->lane_events[l]) PB_CUDA(c, cudaEventCreateWithFlags(&c->lane_events[l], cudaEventDisableTiming));
            lanes[l] = c->lane_streams[l];
        }
        PB_CUDA(c, cudaEventRecord(c->ev_root_done, c->stream));       // (matrices, tables, the constant bank; earlier work)
        for (int l = 0; l < nlanes; l++) PB_CUDA(c, cudaStreamWaitEvent(lanes[l], c->ev_root_done, 0));
    }
    cudaStream_t const main_stream = c->stream;
    struct Restore { pb_ctx* c; cudaStream_t s; ~Restore() { c->stream = s; c->bank_held = false; } } restore{c, main_stream};
    for (int k = 0; k < nchunks; k++) {
        const int64_t s0 = bounds[k], s1 = bounds[k + 1];
        const int64_t w = std::min(s1, c->nsites) - s0;
        cudaStream_t lane = lanes[k % nlanes];
        if (w > 0 && !packed2)
            PB_CUDA(c, cudaMemcpy2DAsync(c->d_tips.p + s0, c->spad, states + s0, ld, w, c->ntaxa,
                                         cudaMemcpyHostToDevice, c->copy_stream));
        if (w > 0 && packed2)                                // chunk bounds are multiples of 1024 sites = 256 bytes
            PB_CUDA(c, cudaMemcpy2DAsync(c->d_tips2.p + s0 / 4, c->spad / 4, states + s0 / 4, ld, (w + 3) / 4, c->ntaxa,
                                         cudaMemcpyHostToDevice, c->copy_stream));
        PB_CUDA(c, cudaEventRecord(c->chunk_events[k], c->copy_stream));
        PB_CUDA(c, trace(c->copy_stream));
        c->trace_sites.push_back(s0);
        c->trace_sites.push_back(s1 - s0);
        PB_CUDA(c, cudaStreamWaitEvent(lane, c->chunk_events[k], 0));
        if (packed2) {
            // The program-order words come straight from the 2-bit upload.  The byte-per-site matrix that other
            // consumers read (level kernels, pb_get_clv, a re-pack after pb_set_tree) is expanded when one of them
            // asks for it (ensure_tip_bytes): an evaluation loop over host alignments never pays for it.
            pbk::pack_tips4_from2<<<dim3((unsigned)(((s1 - s0) / 16 + 127) / 128), (unsigned)std::min(nwords, 65535)), 128, 0, lane>>>(
                c->d_tips2.p, c->spad / 4, d_order, d_xorder, nslots, nwords, c->d_packed.p, c->spad, s0 / 4, s1 / 4);
            c->launches++;
        } else {
            pbk::pack_tips4<<<(unsigned)((s1 - s0 + 255) / 256), 256, 0, lane>>>(
                c->d_tips.p, c->spad, d_order, d_xorder, nslots, nwords, c->d_packed.p, c->spad, s0, s1, c->d_bad.p);
            c->launches++;
        }
        c->stream = lane;                                    // (the launchers enqueue on the context's stream)
        c->bank_held = hold_bank;
        if (first_cat_only) {                                // categories beyond lead_cats follow once, over all the sites
            rc = reduced ? launch_fused4t(c, s0, s1, false, 0, lead_cats) : launch_fused4c(c, s0, s1, s0 == 0, 0, lead_cats);
        } else
            rc = reduced ? launch_fused4t(c, s0, s1, false) : launch_fused4(c, s0, s1, false);
        c->stream = main_stream;
        c->bank_held = false;
        if (rc != PB_OK) return rc;
        const int64_t b0 = s0 / 256, b1 = std::min<int64_t>(s1 / 256, nblocks);
        if (root_per_chunk && b1 > b0) {
            pbk::root_combine<<<(unsigned)(b1 - b0), 256, 0, lane>>>(
                c->d_site_cat.p, site_exp_of(c), c->ncat, c->d_weights.p, c->have_site_weights ? c->d_site_weights.p : nullptr, c->nsites,
                c->spad, b0, c->d_per_site.p, c->d_partials.p);
            c->launches++;
        }
        PB_CUDA(c, trace(lane));
    }
    if (side) {
        for (int l = 0; l < nlanes; l++) {
            PB_CUDA(c, cudaEventRecord(c->lane_events[l], lanes[l]));
            PB_CUDA(c, cudaStreamWaitEvent(c->stream, c->lane_events[l], 0));
        }
        if (hold_bank) {
            PB_CUDA(c, cudaEventRecord(g_const_event[c->device & 63], c->stream));
            bank_lock.unlock();
        }
    }
    if (!root_per_chunk) {
        if ((rc = reduced ? launch_fused4t(c, 0, c->spad, false, lead_cats, c->ncat)
                          : launch_fused4c(c, 0, c->spad, false, lead_cats, c->ncat)) != PB_OK) return rc;
        pbk::root_combine<<<(unsigned)nblocks, 256, 0, c->stream>>>(
            c->d_site_cat.p, site_exp_of(c), c->ncat, c->d_weights.p, c->have_site_weights ? c->d_site_weights.p : nullptr, c->nsites,
            c->spad, 0, c->d_per_site.p, c->d_partials.p);
        c->launches++;
    }
    if (c->timing) PB_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    pbk::final_sum<<<1, 256, 0, c->stream>>>(c->d_partials.p, nblocks, c->d_total.p);
    c->launches++;
    if (c->timing) PB_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    PB_CUDA(c, cudaGetLastError());
    PB_CUDA(c, cudaMemcpyAsync(c->h_total, c->d_total.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(c, cudaMemcpyAsync(c->h_total + 1, c->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(c, trace(c->stream));
    c->have_tips = true;
    c->have_masks = false;
    c->pack_dirty = false;
    c->packed_for_reduced = reduced;
    c->tips_checked = true;
    c->all_dirty = true;
    c->evaluated = true;
    c->fresh = true;
    c->tip_bytes_stale = packed2;
    return fetch ? pb_loglik_fetch(c, total, per_site) : PB_OK;
}

int pb_debug_host_trace(pb_ctx* c, double* out, int cap)
{
    PB_ENTER(c);
    const int nchunks = (int)c->trace_sites.size() / 2;
    if (!c->host_trace || nchunks == 0 || (int)c->trace_events.size() < 2 * nchunks + 2)
        return fail(c, PB_ERR_STATE, "option host_trace is off or no pipelined host evaluation has run");
    if (!out || cap < 2 + 4 * nchunks) return fail(c, PB_ERR_INVALID_ARG, "trace: buffer too small");
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    auto ms = [&](int i, double* v) -> cudaError_t {
        float f = 0.f;
        cudaError_t e = cudaEventElapsedTime(&f, c->trace_events[0], c->trace_events[i]);
        *v = f;
        return e;
    };
    out[0] = nchunks;
    for (int k = 0; k < nchunks; k++) {
        out[1 + 4 * k] = (double)c->trace_sites[2 * k];
        out[2 + 4 * k] = (double)c->trace_sites[2 * k + 1];
        PB_CUDA(c, ms(1 + 2 * k, &out[3 + 4 * k]));
        PB_CUDA(c, ms(2 + 2 * k, &out[4 + 4 * k]));
    }
    PB_CUDA(c, ms(1 + 2 * nchunks, &out[1 + 4 * nchunks]));
    return PB_OK;
}

int pb_loglik_host_enqueue(pb_ctx* c, const uint8_t* states, int64_t ld, int packed2)
{
    return loglik_host_impl(c, states, ld, packed2 != 0, nullptr, nullptr, false);
}

int pb_loglik_host(pb_ctx* c, const uint8_t* states, int64_t ld, double* total, double* per_site)
{
    return loglik_host_impl(c, states, ld, false, total, per_site);
}

int pb_loglik_host_packed2(pb_ctx* c, const uint8_t* packed, int64_t ld, double* total, double* per_site)
{
    return loglik_host_impl(c, packed, ld, true, total, per_site);
}

int pb_loglik(pb_ctx* c, double* total, double* per_site)
{
    int rc = pb_loglik_enqueue(c);
    if (rc != PB_OK) return rc;
    return pb_loglik_fetch(c, total, per_site);
}

int pb_result_device_ptrs(pb_ctx* c, const double** d_total, const double** d_per_site)
{
    PB_ENTER(c);
    if (!c->evaluated || !c->fresh) return fail(c, PB_ERR_STATE, "nothing evaluated since the inputs last changed");
    if (d_total) *d_total = c->d_total.p;
    if (d_per_site) *d_per_site = c->d_per_site.p;
    return PB_OK;
}

int pb_get_clv(pb_ctx* c, int node, int cat, double* v, double* carry)
{
    PB_ENTER(c);
    if (!(c->flags & PB_FLAG_KEEP_CLV)) return fail(c, PB_ERR_STATE, "pb_get_clv needs PB_FLAG_KEEP_CLV");
    if (!c->evaluated || !c->fresh) return fail(c, PB_ERR_STATE, "nothing evaluated since the inputs last changed");
    if (node < 0 || node >= c->nnodes || cat < 0 || cat >= c->ncat || !v || !carry)
        return fail(c, PB_ERR_INVALID_ARG, "pb_get_clv: bad node/category/pointer");
    const int n = c->n;
    if (c->child_ptr[node] == c->child_ptr[node + 1] && !c->have_masks) {      // a leaf's row of the byte-per-site matrix
        const int rc = ensure_tip_bytes(c);
        if (rc != PB_OK) return rc;
    }
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->child_ptr[node] == c->child_ptr[node + 1] && c->have_masks) {   // leaf: set indicator, carry 0 (NaN if missing)
        std::vector<uint64_t> row(c->nsites);
        PB_CUDA(c, cudaMemcpy(row.data(), c->d_masks.p + (size_t)c->tip_row[node] * c->spad,
                              c->nsites * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        for (int64_t s = 0; s < c->nsites; s++) {
            for (int i = 0; i < n; i++) v[s * n + i] = (double)((row[s] >> i) & 1u);
            carry[s] = row[s] ? 0.0 : NAN;
        }
        return PB_OK;
    }
    if (c->child_ptr[node] == c->child_ptr[node + 1]) {      // leaf: indicator, carry 0
        std::vector<uint8_t> row(c->nsites);
        PB_CUDA(c, cudaMemcpy(row.data(), c->d_tips.p + (size_t)c->tip_row[node] * c->spad, c->nsites, cudaMemcpyDeviceToHost));
        for (int64_t s = 0; s < c->nsites; s++) {
            for (int i = 0; i < n; i++) v[s * n + i] = (i == row[s]) ? 1.0 : 0.0;
            carry[s] = 0.0;
        }
        return PB_OK;
    }
    const int slot = c->slot_of_node[node];
    std::vector<double> planes((size_t)(n + 1) * c->nsites);
    const double* src = c->d_pool.p + (size_t)slot * c->slot_stride + (size_t)cat * (n + 1) * c->spad;
    PB_CUDA(c, cudaMemcpy2D(planes.data(), c->nsites * sizeof(double), src, c->spad * sizeof(double),
                            c->nsites * sizeof(double), n + 1, cudaMemcpyDeviceToHost));
    for (int64_t s = 0; s < c->nsites; s++) {
        for (int i = 0; i < n; i++) v[s * n + i] = planes[(size_t)i * c->nsites + s];
        carry[s] = planes[(size_t)n * c->nsites + s];
    }
    return PB_OK;
}

int pb_conditional_simulation(pb_ctx* c, uint64_t seed, uint8_t* states, int64_t ld)
{
    PB_ENTER(c);
    if (!(c->flags & PB_FLAG_KEEP_CLV)) return fail(c, PB_ERR_STATE, "pb_conditional_simulation needs PB_FLAG_KEEP_CLV");
    if (!c->evaluated || !c->fresh) return fail(c, PB_ERR_STATE, "nothing evaluated since the inputs last changed");
    if (c->ncat != 1) return fail(c, PB_ERR_STATE, "pb_conditional_simulation: one rate category only (as the reference)");
    if (!states || ld < c->nsites) return fail(c, PB_ERR_INVALID_ARG, "pb_conditional_simulation: bad output buffer");
    // parents before children
    std::vector<pbk::SimNode> pre;
    pre.reserve(c->nnodes);
    std::vector<int> stack{c->root};
    while (!stack.empty()) {
        const int v = stack.back();
        stack.pop_back();
        const bool leaf = c->child_ptr[v] == c->child_ptr[v + 1];
        pre.push_back({v, c->parent[v], leaf ? -1 : c->slot_of_node[v], leaf ? c->tip_row[v] : -1});
        for (int k = c->child_ptr[v + 1] - 1; k >= c->child_ptr[v]; k--) stack.push_back(c->child_idx[k]);
    }
    if (!c->have_masks) {
        const int rc = ensure_tip_bytes(c);
        if (rc != PB_OK) return rc;
    }
    PB_CUDA(c, c->d_sim_nodes.ensure(pre.size()));
    PB_CUDA(c, c->d_states.ensure((size_t)c->nnodes * c->spad));
    PB_CUDA(c, cudaMemcpyAsync(c->d_sim_nodes.p, pre.data(), pre.size() * sizeof(pbk::SimNode), cudaMemcpyHostToDevice, c->stream));
    pbk::conditional_simulation<<<(unsigned)((c->nsites + 255) / 256), 256, 0, c->stream>>>(
        c->d_sim_nodes.p, (int)pre.size(), c->nnodes, c->n, c->d_P.p, c->d_pi.p, c->d_pool.p, c->slot_stride, c->spad,
        c->d_tips.p, c->have_masks ? c->d_masks.p : nullptr, c->nsites, seed, c->d_states.p);
    c->launches++;
    PB_CUDA(c, cudaGetLastError());
    PB_CUDA(c, cudaMemcpy2DAsync(states, (size_t)ld, c->d_states.p, (size_t)c->spad, (size_t)c->nsites, c->nnodes,
                                 cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));            // `pre` is read by the asynchronous upload
    return PB_OK;
}

int pb_get_pmatrices(pb_ctx* c, double* P)
{
    PB_ENTER(c);
    if (!P) return fail(c, PB_ERR_INVALID_ARG, "P is null");
    if (!c->have_tree) return fail(c, PB_ERR_STATE, "tree not set");
    int rc = build_pmatrices(c);
    if (rc != PB_OK) return rc;
    const int n = c->n;
    const size_t count = (size_t)c->ncat * c->nnodes;
    std::vector<double> pr(count * n * n);
    PB_CUDA(c, cudaMemcpyAsync(pr.data(), c->d_P.p, sizeof(double) * pr.size(), cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(c, cudaStreamSynchronize(c->stream));
    to_row_major(n, count, pr.data(), P);   // transpose is its own inverse
    if (c->psrc != PSRC_DIRECT)
        for (int k = 0; k < c->ncat; k++)   // the root has no branch: report identity
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) P[((size_t)k * c->nnodes + c->root) * n * n + j * n + i] = (i == j) ? 1.0 : 0.0;
    return PB_OK;
}

int pb_set_option(pb_ctx* c, const char* name, double value)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    if (!name) return fail(c, PB_ERR_INVALID_ARG, "option name is null");
    const std::string k(name);
    const int iv = (int)value;
    if (k == "fused_block") {
        if (iv != 128 && iv != 256 && iv != 512) return fail(c, PB_ERR_INVALID_ARG, "fused_block must be 128, 256 or 512");
        c->fused_block = iv;
    } else if (k == "fused_spt") {
        if (iv != 1 && iv != 2 && iv != 4) return fail(c, PB_ERR_INVALID_ARG, "fused_spt must be 1, 2 or 4");
        c->fused_spt = iv;
        c->fused_spt_set = true;
    } else if (k == "fused_const_cats") {
        if (iv < 1 || iv > 8) return fail(c, PB_ERR_INVALID_ARG, "fused_const_cats must be in 1..8");
        c->fused_const_cats = iv;
        c->fused_const_cats_set = true;
    } else if (k == "fused_const") {
        c->fused_const = iv != 0;
    } else if (k == "fused_reduced") {
        c->fused_reduced = iv != 0;
    } else if (k == "fused_masks") {
        c->fused_masks = iv != 0;
    } else if (k == "fused_occ") {
        if (iv != 0 && iv != 4) return fail(c, PB_ERR_INVALID_ARG, "fused_occ must be 0 or 4");
        c->fused_occ = iv;
    } else if (k == "fused_xor") {
        c->fused_xor = iv != 0;
        c->r_leaves = -1;                  // replan (and repack) with the other coding
    } else if (k == "fused_triple") {
        c->fused_triple = iv != 0;
    } else if (k == "fused_preapply") {
        if (iv != 0 && !FUSED_PREAPPLY) return fail(c, PB_ERR_INVALID_ARG, "fused_preapply: this build has no FUSED_PREAPPLY kernels");
        c->fused_preapply = iv != 0;
    } else if (k == "fused_clade_force") {
        c->fused_clade_force = iv != 0;
    } else if (k == "fused_clade_leaves") {
        if (iv < 2 || iv > 8) return fail(c, PB_ERR_INVALID_ARG, "fused_clade_leaves must be in 2..8");
        c->fused_clade_leaves = iv;
    } else if (k == "fused_cherry") {
        c->fused_cherry = iv != 0;
    } else if (k == "fused_pow2") {
        c->fused_pow2 = iv != 0;
    } else if (k == "fused_exact_log") {
        c->fused_exact_log = iv != 0;
    } else if (k == "level4_async") {
        c->level4_async = iv != 0;
    } else if (k == "level4_waves") {
        if (iv < 1 || iv > 64) return fail(c, PB_ERR_INVALID_ARG, "level4_waves must be in 1..64");
        c->level4_waves = iv;
    } else if (k == "level4_min_blocks") {
        if (iv < 3 || iv > 6) return fail(c, PB_ERR_INVALID_ARG, "level4_min_blocks must be 3..6");
        c->level4_minb = iv;
    } else if (k == "fused_dmma") {
        c->fused_dmma = value != 0.0;
    } else if (k == "fused_dmma_cherry") {
        c->fused_dmma_cherry = iv != 0;
    } else if (k == "fused_dmma_builder_waves") {
        if (iv < 0 || iv > 8) return fail(c, PB_ERR_INVALID_ARG, "fused_dmma_builder_waves must be in 0..8");
        c->dense_builder_waves = iv;
    } else if (k == "fused_dmma_triple") {
        c->fused_dmma_triple = iv != 0;
        c->dense_reduced_built = false;
    } else if (k == "fused_dmma_version") {
        if (iv < 1 || iv > 3) return fail(c, PB_ERR_INVALID_ARG, "fused_dmma_version must be 1, 2 or 3");
        c->fused_dmma_version = iv;
    } else if (k == "fused_dmma_slots") {
        if (iv != 0 && (iv < 2 || iv > 14)) return fail(c, PB_ERR_INVALID_ARG, "fused_dmma_slots must be 0 or 2..14");
        c->fused_dmma_slots = iv;
    } else if (k == "fused_dmma_spill") {
        c->fused_dmma_spill = iv != 0;
    } else if (k == "fused_dmma_stagger") {
        if (iv < 0) return fail(c, PB_ERR_INVALID_ARG, "fused_dmma_stagger must be 0..1000000 ns");
        c->fused_dmma_stagger = iv;
    } else if (k == "dmma_version") {
        if (iv != 1 && iv != 2) return fail(c, PB_ERR_INVALID_ARG, "dmma_version must be 1 or 2");
        c->dmma_version = iv;
    } else if (k == "dmma_tip_levels_v1") {
        c->dmma_tip_levels_v1 = iv != 0;
    } else if (k == "dmma_per_level") {
        c->dmma_per_level = iv != 0;
    } else if (k == "dmma_group_tiles") {
        if (iv < 8 || iv > 1024 || iv % 8) return fail(c, PB_ERR_INVALID_ARG, "dmma_group_tiles must be a multiple of 8 in 8..1024");
        c->dmma_seq_tiles = iv;
    } else if (k == "fused_const_group") {
        c->fused_const_group = iv != 0;
    } else if (k == "host_lead_cats") {
        if (iv < 1 || iv > 64) return fail(c, PB_ERR_INVALID_ARG, "host_lead_cats must be in 1..64");
        c->host_lead_cats = iv;
    } else if (k == "graph") {
        c->use_graph = iv != 0;
    } else if (k == "fused_fold") {
        c->fused_fold = iv != 0;
        c->tables_stale = true;
    } else if (k == "fused_const_keep") {
        c->fused_const_keep = iv != 0;
    } else if (k == "fused_split") {
        if (iv < 1 || iv > 64) return fail(c, PB_ERR_INVALID_ARG, "fused_split must be in 1..64");
        c->fused_split = iv;
    } else if (k == "host_tail_sites") {
        if (iv < -1) return fail(c, PB_ERR_INVALID_ARG, "host_tail_sites must be >= -1");
        c->host_tail_sites = iv;
    } else if (k == "host_lanes") {
        if (iv < 2 || iv > 4) return fail(c, PB_ERR_INVALID_ARG, "host_lanes must be in 2..4");
        c->host_lanes = iv;
    } else if (k == "host_side_streams") {
        c->host_side_streams = iv != 0;
    } else if (k == "host_trace") {
        c->host_trace = iv != 0;
    } else if (k == "host_chunk_align") {
        c->host_chunk_align = iv != 0;
    } else if (k == "host_chunks") {
        if (iv < 1 || iv > 64) return fail(c, PB_ERR_INVALID_ARG, "host_chunks must be in 1..64");
        c->host_chunks = iv;
    } else if (k == "dense_scalar") {
        c->force_dense = iv != 0;
    } else {
        return fail(c, PB_ERR_INVALID_ARG, "unknown option: " + k);
    }
    if (c->fused_block == 512 && c->fused_spt > 1 && !c->fused_const) c->fused_spt = 1;
    if (c->fused_block == 256 && c->fused_spt == 4 && c->fused_exact_log) c->fused_spt = 2;
    if (c->fused_block == 512 && c->fused_spt == 4) c->fused_spt = 2;
    bool reduced = false;
    return want_reduced(c, &reduced);          // the options decide which 4-state program runs (pb_path_name)
}

int pb_set_timing(pb_ctx* c, int on)
{
    PB_ENTER(c);
    PB_RESHAPES(c);
    if (on && !c->ev_ring[0][0])
        for (auto& set : c->ev_ring)
            for (int i = 0; i < 4; i++) PB_CUDA(c, cudaEventCreate(&set[i]));
    c->timing = on != 0;
    c->kernel_timing = on == 2;          // (2: also an event pair around every launch of the dominant kernel)
    return PB_OK;
}

int pb_get_timing_at(pb_ctx* c, int back, double* ms)
{
    PB_ENTER(c);
    if (!c->timing || c->timed_evals == 0 || !ms) return fail(c, PB_ERR_STATE, "timing is off or nothing was evaluated");
    if (back < 0 || back >= pb_ctx::TIMING_RING || back >= c->timed_evals)
        return fail(c, PB_ERR_INVALID_ARG, "pb_get_timing_at: that evaluation is not (or no longer) recorded");
    cudaEvent_t* set = c->ev_ring[(c->timed_evals - 1 - back) % pb_ctx::TIMING_RING];
    PB_CUDA(c, cudaEventSynchronize(set[3]));
    for (int i = 0; i < 3; i++) {
        float t = 0.f;
        PB_CUDA(c, cudaEventElapsedTime(&t, set[i], set[i + 1]));
        ms[i] = t;
    }
    return PB_OK;
}

int pb_get_timing(pb_ctx* c, double* ms) { return pb_get_timing_at(c, 0, ms); }

int pb_get_kernel_timing(pb_ctx* c, double* ms, int* launches)
{
    PB_ENTER(c);
    if (!c->timing || !c->kernel_timing || c->timed_evals == 0 || !ms)
        return fail(c, PB_ERR_STATE, "kernel timing is off (pb_set_timing(ctx, 2)) or nothing was evaluated");
    double sum = 0.0;
    for (size_t i = 0; i + 1 < c->kev_n; i += 2) {
        float t = 0.f;
        PB_CUDA(c, cudaEventSynchronize(c->kev[i + 1]));
        PB_CUDA(c, cudaEventElapsedTime(&t, c->kev[i], c->kev[i + 1]));
        sum += t;
    }
    *ms = sum;
    if (launches) *launches = (int)(c->kev_n / 2);
    return PB_OK;
}

int pb_measure_fp64_peak(int device, int kind, double* tflops)
{
    if (!tflops || (kind != 0 && kind != 1)) return PB_ERR_INVALID_ARG;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PB_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 4, threads = 256, iters = 20000;
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads) != cudaSuccess) return PB_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {                  // (the first repetition warms the clocks up)
        cudaEventRecord(e0);
        if (kind == 0) pbk::fp64_peak_dfma<<<blocks, threads>>>(out, iters);
        else pbk::fp64_peak_dmma<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (err != cudaSuccess || !(best > 0.f)) return PB_ERR_CUDA;
    // flops per thread and iteration: 8 chains x 2 (one FMA), or 8 chains x 512 / 32 (an 8 x 8 x 4 MMA is 512 flops per warp)
    const double per_thread_iter = kind == 0 ? 16.0 : 8.0 * 512.0 / 32.0;
    *tflops = per_thread_iter * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    return PB_OK;
}

int64_t pb_kernel_launches(const pb_ctx* c) { return c ? c->launches : 0; }

int64_t pb_graph_replays(const pb_ctx* c) { return c ? c->graph_replays : 0; }

const char* pb_path_name(const pb_ctx* c)
{
    if (!c) return "";
    bool reduced = false;
    want_reduced(const_cast<pb_ctx*>(c), &reduced);       // (re)plans the 4-state program for the current options and inputs
    switch (path_of(c)) {
    case PATH_FUSED4: return "fused4";
    case PATH_LEVEL4: return "level4";
    case PATH_LEVEL_DMMA: return "level_dmma";
    case PATH_FUSED_DMMA: return "fused_dmma";
    default: return "level_dense";
    }
}

static int standalone_setup(int device, int n, int count)
{
    if (n < 1 || n > 64 || count < 1) return fail(nullptr, PB_ERR_INVALID_ARG, "bad n or count");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(nullptr, PB_ERR_CUDA, "no CUDA device (no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, PB_ERR_INVALID_ARG, "device index out of range");
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    return PB_OK;
}

int pb_transition_matrices_eigen(int device, int n, int count, const double* U, const double* lambda,
                                 const double* Uinv, const double* t, double* out)
{
    int rc = standalone_setup(device, n, count);
    if (rc != PB_OK) return rc;
    if (!U || !lambda || !Uinv || !t || !out) return fail(nullptr, PB_ERR_INVALID_ARG, "null pointer");
    std::vector<double> ur((size_t)n * n), uir((size_t)n * n), res((size_t)count * n * n);
    to_row_major(n, 1, U, ur.data());
    to_row_major(n, 1, Uinv, uir.data());
    DevBuf<double> dU, dUi, dl, dt, dP;
    pb_ctx* none = nullptr;
    PB_CUDA(none, dU.ensure((size_t)n * n)); PB_CUDA(none, dUi.ensure((size_t)n * n)); PB_CUDA(none, dl.ensure(n));
    PB_CUDA(none, dt.ensure(count)); PB_CUDA(none, dP.ensure((size_t)count * n * n));
    PB_CUDA(none, cudaMemcpy(dU.p, ur.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    PB_CUDA(none, cudaMemcpy(dUi.p, uir.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    PB_CUDA(none, cudaMemcpy(dl.p, lambda, sizeof(double) * n, cudaMemcpyHostToDevice));
    PB_CUDA(none, cudaMemcpy(dt.p, t, sizeof(double) * count, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(n + 2 * n * n) * sizeof(double);
    PB_CUDA(none, cudaFuncSetAttribute(pbk::pmat_eigen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pbk::pmat_eigen<<<dim3(count, 1), 256, smem>>>(n, count, -1, dU.p, dl.p, dUi.p, dt.p, nullptr, dP.p);
    PB_CUDA(none, cudaGetLastError());
    PB_CUDA(none, cudaMemcpy(res.data(), dP.p, sizeof(double) * res.size(), cudaMemcpyDeviceToHost));
    to_row_major(n, count, res.data(), out);
    dU.release(); dUi.release(); dl.release(); dt.release(); dP.release();
    return PB_OK;
}

int pb_transition_matrices_expm(int device, int n, int count, const double* Q, const double* t, double* out)
{
    int rc = standalone_setup(device, n, count);
    if (rc != PB_OK) return rc;
    if (!Q || !t || !out) return fail(nullptr, PB_ERR_INVALID_ARG, "null pointer");
    cudaDeviceProp prop;
    pb_ctx* none = nullptr;
    PB_CUDA(none, cudaGetDeviceProperties(&prop, device));
    std::vector<double> qr((size_t)n * n), res((size_t)count * n * n);
    to_row_major(n, 1, Q, qr.data());
    DevBuf<double> dQ, dt, dP;
    PB_CUDA(none, dQ.ensure((size_t)n * n)); PB_CUDA(none, dt.ensure(count)); PB_CUDA(none, dP.ensure((size_t)count * n * n));
    PB_CUDA(none, cudaMemcpy(dQ.p, qr.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    PB_CUDA(none, cudaMemcpy(dt.p, t, sizeof(double) * count, cudaMemcpyHostToDevice));
    if (pbk::launch_expm_batched(n, count, 1, -1, dQ.p, dt.p, nullptr, dP.p, prop.sharedMemPerBlockOptin, 0) != 0)
        return fail(nullptr, PB_ERR_CUDA, "expm kernel: not enough shared memory for this state count");
    PB_CUDA(none, cudaGetLastError());
    PB_CUDA(none, cudaMemcpy(res.data(), dP.p, sizeof(double) * res.size(), cudaMemcpyDeviceToHost));
    to_row_major(n, count, res.data(), out);
    dQ.release(); dt.release(); dP.release();
    return PB_OK;
}

}  // extern "C"
