// pb_sharded_*: one host thread drives the GPUs of one box (include/phylo_b200.h, "multi-GPU").
//
// Sites are independent given the tree and the model (the reference sums per-site values,
// lib/felsenstein.ml:73-75), so the alignment is cut into contiguous column shards, one pb_ctx per device;
// tree, branch lengths, model and categories are replicated (P(t) is rebuilt redundantly on every device:
// cheaper than a broadcast) and the only exchange is the all-reduce of one double per evaluation.  NCCL is
// bound at run time (dlopen of libnccl.so.2: in a process that already holds NCCL, e.g. under torch, that is
// the loaded library), so the product library itself has no link-time dependency on it; a missing or failing
// NCCL surfaces as PB_ERR_NCCL from pb_create_sharded / pb_sharded_loglik*, never as a crash.
#include "../../include/phylo_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <string>
#include <vector>

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;

    bool load()
    {
        if (handle) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { why = std::string("cannot load NCCL: ") + dlerror(); return false; }
#define PB_SYM(field, sym)                                                                    \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, sym));                            \
    if (!field) { why = std::string("NCCL symbol missing: ") + sym; handle = nullptr; return false; }
        PB_SYM(CommInitAll, "ncclCommInitAll")
        PB_SYM(CommDestroy, "ncclCommDestroy")
        PB_SYM(AllReduce, "ncclAllReduce")
        PB_SYM(GroupStart, "ncclGroupStart")
        PB_SYM(GroupEnd, "ncclGroupEnd")
        PB_SYM(GetErrorString, "ncclGetErrorString")
#undef PB_SYM
        return true;
    }
};

NcclApi g_nccl;
thread_local std::string g_sharded_create_error;

}  // namespace

struct pb_sharded {
    int ndev = 0;
    std::vector<int> devices;
    std::vector<pb_ctx*> ctx;
    std::vector<int64_t> lo, hi;          // shard d holds columns [lo[d], hi[d])
    std::vector<ncclComm_t> comms;
    std::vector<cudaStream_t> streams;    // private stream of every shard (the collective is enqueued behind its evaluation)
    int64_t nsites = 0;
    int ntaxa = 0, nstates = 0;
    int reduce = PB_REDUCE_NCCL;
    double* h_total = nullptr;            // pinned
    mutable std::string err;
};

namespace {

int sfail(pb_sharded* s, int code, const std::string& msg)
{
    if (s) s->err = msg; else g_sharded_create_error = msg;
    return code;
}

// A failure inside one shard's context: carry its message up.
int shard_fail(pb_sharded* s, int d, int rc)
{
    return sfail(s, rc, "device " + std::to_string(s->devices[d]) + ": " + pb_last_error(s->ctx[d]));
}

#define PB_EACH(s, call)                                   \
    for (int d = 0; d < (s)->ndev; d++) {                  \
        const int rc_ = (call);                            \
        if (rc_ != PB_OK) return shard_fail((s), d, rc_);  \
    }

// enqueue phase done on every device: combine the device-resident totals and fetch
int combine_and_fetch(pb_sharded* s, double* total, double* per_site)
{
    std::vector<const double*> d_total(s->ndev, nullptr);
    for (int d = 0; d < s->ndev; d++) {
        const int rc = pb_result_device_ptrs(s->ctx[d], &d_total[d], nullptr);
        if (rc != PB_OK) return shard_fail(s, d, rc);
    }
    if (s->ndev > 1 && s->reduce == PB_REDUCE_NCCL) {
        // in place on every device's total, behind that device's evaluation on its own stream
        ncclResult_t r = g_nccl.GroupStart();
        for (int d = 0; d < s->ndev && r == ncclSuccess; d++) {
            cudaSetDevice(s->devices[d]);
            r = g_nccl.AllReduce(d_total[d], const_cast<double*>(d_total[d]), 1, ncclDouble, ncclSum, s->comms[d], s->streams[d]);
        }
        const ncclResult_t r2 = g_nccl.GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess)
            return sfail(s, PB_ERR_NCCL, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r != ncclSuccess ? r : r2));
    }
    // every shard: wait for its stream (evaluation + collective), range check of its tips, per-site values
    double ordered = 0.0;
    for (int d = 0; d < s->ndev; d++) {
        double local = 0.0;
        const int rc = pb_loglik_fetch(s->ctx[d], &local, per_site ? per_site + s->lo[d] : nullptr);
        if (rc != PB_OK) return shard_fail(s, d, rc);
        ordered += local;                 // the local totals, summed in shard order: bit-reproducible whatever NCCL does
    }
    if (s->ndev == 1 || s->reduce == PB_REDUCE_ORDERED) {
        if (total) *total = ordered;
        return PB_OK;
    }
    cudaSetDevice(s->devices[0]);
    const cudaError_t e = cudaMemcpy(s->h_total, d_total[0], sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return sfail(s, PB_ERR_CUDA, std::string("reading the reduced total: ") + cudaGetErrorString(e));
    if (total) *total = *s->h_total;
    return PB_OK;
}

}  // namespace

extern "C" {

const char* pb_sharded_last_error(const pb_sharded* s) { return s ? s->err.c_str() : g_sharded_create_error.c_str(); }

pb_sharded* pb_create_sharded(int ndev, const int* devices, int nstates, int ncat, int64_t nsites_total, int ntaxa,
                              unsigned flags)
{
    if (ndev < 1 || ndev > 64 || !devices) { sfail(nullptr, PB_ERR_INVALID_ARG, "pb_create_sharded: ndev must be in 1..64"); return nullptr; }
    if (nsites_total < ndev) { sfail(nullptr, PB_ERR_INVALID_ARG, "pb_create_sharded: more devices than sites"); return nullptr; }
    for (int a = 0; a < ndev; a++)
        for (int b = a + 1; b < ndev; b++)
            if (devices[a] == devices[b]) {     // (NCCL would refuse it too: one rank per device)
                sfail(nullptr, PB_ERR_NCCL, "pb_create_sharded: a device is listed twice (NCCL needs one rank per device)");
                return nullptr;
            }
    pb_sharded* s = new pb_sharded();
    s->ndev = ndev;
    s->devices.assign(devices, devices + ndev);
    s->nsites = nsites_total;
    s->ntaxa = ntaxa;
    s->nstates = nstates;
    for (int d = 0; d < ndev; d++) {
        // contiguous column blocks whose bounds are multiples of 4 sites (a 2-bit packed host alignment shards on bytes)
        const int64_t lo = nsites_total * d / ndev / 4 * 4;
        const int64_t hi = d + 1 == ndev ? nsites_total : nsites_total * (d + 1) / ndev / 4 * 4;
        s->lo.push_back(lo);
        s->hi.push_back(hi);
    }
    for (int d = 0; d < ndev; d++) {
        if (s->hi[d] <= s->lo[d]) { sfail(nullptr, PB_ERR_INVALID_ARG, "pb_create_sharded: a shard would be empty"); pb_sharded_destroy(s); return nullptr; }
        pb_ctx* c = pb_create(devices[d], nstates, ncat, s->hi[d] - s->lo[d], ntaxa, flags);
        if (!c) {
            const std::string why = pb_last_error(nullptr);
            const bool arg = why.find("must be") != std::string::npos || why.find("out of range") != std::string::npos;
            sfail(nullptr, arg ? PB_ERR_INVALID_ARG : PB_ERR_CUDA, "device " + std::to_string(devices[d]) + ": " + why);
            pb_sharded_destroy(s);
            return nullptr;
        }
        s->ctx.push_back(c);
        const double* unused = nullptr;
        (void)unused;
    }
    cudaSetDevice(devices[0]);
    if (cudaMallocHost(&s->h_total, sizeof(double)) != cudaSuccess) {
        sfail(nullptr, PB_ERR_NOMEM, "pinned allocation failed");
        pb_sharded_destroy(s);
        return nullptr;
    }
    // the shard's collective runs on a stream we own and the context evaluates on
    for (int d = 0; d < ndev; d++) {
        cudaStream_t st = nullptr;
        cudaSetDevice(devices[d]);
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess || pb_set_stream(s->ctx[d], st) != PB_OK) {
            sfail(nullptr, PB_ERR_CUDA, "stream creation failed on device " + std::to_string(devices[d]));
            pb_sharded_destroy(s);
            return nullptr;
        }
        s->streams.push_back(st);
    }
    if (ndev > 1) {
        if (!g_nccl.load()) { sfail(nullptr, PB_ERR_NCCL, g_nccl.why); pb_sharded_destroy(s); return nullptr; }
        s->comms.assign(ndev, nullptr);
        const ncclResult_t r = g_nccl.CommInitAll(s->comms.data(), ndev, devices);
        if (r != ncclSuccess) {
            s->comms.clear();
            sfail(nullptr, PB_ERR_NCCL, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r));
            pb_sharded_destroy(s);
            return nullptr;
        }
    }
    return s;
}

void pb_sharded_destroy(pb_sharded* s)
{
    if (!s) return;
    for (size_t d = 0; d < s->ctx.size(); d++) pb_destroy(s->ctx[d]);       // waits for the shard's stream
    for (ncclComm_t c : s->comms)
        if (c) g_nccl.CommDestroy(c);
    for (size_t d = 0; d < s->streams.size(); d++) {
        cudaSetDevice(s->devices[d]);
        cudaStreamDestroy(s->streams[d]);
    }
    if (s->h_total) cudaFreeHost(s->h_total);
    delete s;
}

int pb_sharded_ndev(const pb_sharded* s) { return s ? s->ndev : 0; }

pb_ctx* pb_sharded_ctx(pb_sharded* s, int d) { return (s && d >= 0 && d < s->ndev) ? s->ctx[d] : nullptr; }

int pb_sharded_shard_bounds(const pb_sharded* s, int d, int64_t* lo, int64_t* hi)
{
    if (!s || d < 0 || d >= s->ndev) return PB_ERR_INVALID_ARG;
    if (lo) *lo = s->lo[d];
    if (hi) *hi = s->hi[d];
    return PB_OK;
}

int pb_sharded_set_reduce(pb_sharded* s, int how)
{
    if (!s) return PB_ERR_INVALID_ARG;
    if (how != PB_REDUCE_NCCL && how != PB_REDUCE_ORDERED) return sfail(s, PB_ERR_INVALID_ARG, "pb_sharded_set_reduce: unknown mode");
    s->reduce = how;
    return PB_OK;
}

// ---- replicated inputs ------------------------------------------------------------------------------------
int pb_sharded_set_mode(pb_sharded* s, int mode)
{
    if (!s) return PB_ERR_INVALID_ARG;
    PB_EACH(s, pb_set_mode(s->ctx[d], mode));
    return PB_OK;
}

int pb_sharded_set_option(pb_sharded* s, const char* name, double value)
{
    if (!s) return PB_ERR_INVALID_ARG;
    PB_EACH(s, pb_set_option(s->ctx[d], name, value));
    return PB_OK;
}

int pb_sharded_set_tree(pb_sharded* s, int nnodes, int root, const int32_t* child_ptr, const int32_t* child_idx,
                        const int32_t* tip_row, const double* branch_len)
{
    if (!s) return PB_ERR_INVALID_ARG;
    PB_EACH(s, pb_set_tree(s->ctx[d], nnodes, root, child_ptr, child_idx, tip_row, branch_len));
    return PB_OK;
}

int pb_sharded_set_branch_lengths(pb_sharded* s, const double* branch_len)
{
    if (!s) return PB_ERR_INVALID_ARG;
    PB_EACH(s, pb_set_branch_lengths(s->ctx[d], branch_len));
    return PB_OK;
}

int pb_sharded_update_branch_length(pb_sharded* s, int node, double t)
{
    if (!s) return PB_ERR_INVALID_ARG;
    PB_EACH(s, pb_update_branch_length(s->ctx[d], node, t));
    return PB_OK;
}

int pb_sharded_set_root_frequencies(pb_sharded* s, const double* pi)
{
    if (!s) return PB_ERR_INVALID_ARG;
    PB_EACH(s, pb_set_root_frequencies(s->ctx[d], pi));
    return PB_OK;
}

int pb_sharded_set_categories(pb_sharded* s, const double* rates, const double* weights)
{
    if (!s) return PB_ERR_INVALID_ARG;
    PB_EACH(s, pb_set_categories(s->ctx[d], rates, weights));
    return PB_OK;
}

int pb_sharded_set_eigensystem(pb_sharded* s, const double* U, const double* lambda, const double* Uinv)
{
    if (!s) return PB_ERR_INVALID_ARG;
    PB_EACH(s, pb_set_eigensystem(s->ctx[d], U, lambda, Uinv));
    return PB_OK;
}

int pb_sharded_set_rate_matrix(pb_sharded* s, const double* Q, const double* pi)
{
    if (!s) return PB_ERR_INVALID_ARG;
    if (!Q || !pi) return sfail(s, PB_ERR_INVALID_ARG, "rate matrix: null pointer");
    // one host eigen-decomposition, installed on every device
    const int n = s->nstates;
    std::vector<double> U((size_t)n * n), Ui((size_t)n * n), lam(n);
    const int rc = pb_host_reversible_eigensystem(n, Q, pi, U.data(), lam.data(), Ui.data());
    if (rc != PB_OK) return sfail(s, rc, "pb_sharded_set_rate_matrix: Q is not reversible with respect to pi");
    return pb_sharded_set_eigensystem(s, U.data(), lam.data(), Ui.data());
}

int pb_sharded_set_rate_matrix_general(pb_sharded* s, const double* Q)
{
    if (!s) return PB_ERR_INVALID_ARG;
    PB_EACH(s, pb_set_rate_matrix_general(s->ctx[d], Q));
    return PB_OK;
}

int pb_sharded_set_pmatrices(pb_sharded* s, const double* P)
{
    if (!s) return PB_ERR_INVALID_ARG;
    PB_EACH(s, pb_set_pmatrices(s->ctx[d], P));
    return PB_OK;
}

// ---- sharded inputs ---------------------------------------------------------------------------------------
int pb_sharded_set_tips(pb_sharded* s, const uint8_t* states, int64_t ld)
{
    if (!s) return PB_ERR_INVALID_ARG;
    if (!states || ld < s->nsites) return sfail(s, PB_ERR_INVALID_ARG, "tips: null pointer or ld < nsites");
    PB_EACH(s, pb_set_tips(s->ctx[d], states + s->lo[d], ld));
    return PB_OK;
}

int pb_sharded_set_tips_packed2(pb_sharded* s, const uint8_t* packed, int64_t ld)
{
    if (!s) return PB_ERR_INVALID_ARG;
    if (!packed || ld < (s->nsites + 3) / 4) return sfail(s, PB_ERR_INVALID_ARG, "packed tips: null pointer or ld < ceil(nsites / 4)");
    PB_EACH(s, pb_set_tips_packed2(s->ctx[d], packed + s->lo[d] / 4, ld));      // shard bounds are multiples of 4 sites
    return PB_OK;
}

int pb_sharded_set_tip_masks(pb_sharded* s, const uint64_t* masks, int64_t ld)
{
    if (!s) return PB_ERR_INVALID_ARG;
    if (!masks || ld < s->nsites) return sfail(s, PB_ERR_INVALID_ARG, "masks: null pointer or ld < nsites");
    PB_EACH(s, pb_set_tip_masks(s->ctx[d], masks + s->lo[d], ld));
    return PB_OK;
}

// ---- evaluation -------------------------------------------------------------------------------------------
int pb_sharded_loglik(pb_sharded* s, double* total, double* per_site)
{
    if (!s) return PB_ERR_INVALID_ARG;
    PB_EACH(s, pb_loglik_enqueue(s->ctx[d]));          // every device works before anyone waits
    return combine_and_fetch(s, total, per_site);
}

int pb_sharded_loglik_host(pb_sharded* s, const uint8_t* states, int64_t ld, int packed2, double* total, double* per_site)
{
    if (!s) return PB_ERR_INVALID_ARG;
    if (!states) return sfail(s, PB_ERR_INVALID_ARG, "tips: null pointer");
    PB_EACH(s, pb_loglik_host_enqueue(s->ctx[d], states + (packed2 ? s->lo[d] / 4 : s->lo[d]), ld, packed2));
    return combine_and_fetch(s, total, per_site);
}

}  // extern "C"
