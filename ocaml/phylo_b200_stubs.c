/*
 * OCaml C stubs over include/phylo_b200.h — the reference-side binding a maintainer of
 * biocaml/phylogenetics would add (lib/dune: foreign_stubs + c_library_flags -lphylo_b200).
 *
 * The build image has no OCaml toolchain (no caml headers), so this file is never built against the real
 * OCaml runtime here.  It IS compiled (-Wall -Wextra -Werror) against a mock of the runtime's C interface
 * (tests/ocaml_mock/caml/) and executed on the GPU by tests/ocaml_mock/stubs_driver.c, which replays the call
 * sequences of ocaml/phylo_b200.ml with Bigarray descriptors built in C (tests/test_ocaml_stubs.py).  See
 * INTEGRATION.md.
 *
 * Buffers: Linear_algebra.mat / vec are Lacaml.D.mat / vec = float64 Bigarrays in Fortran layout
 * (lib/linear_algebra.ml:151-152); Caml_ba_data_val gives the column-major double* the ABI takes
 * as is.  Bigarray payloads live outside the OCaml heap and do not move, so the runtime lock can
 * be released around the blocking calls.
 */
#include <string.h>

#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <caml/threads.h>

#include "phylo_b200.h"

#define Ctx_val(v) (*((pb_ctx **)Data_custom_val(v)))

static void ctx_finalize(value v)
{
    if (Ctx_val(v)) { pb_destroy(Ctx_val(v)); Ctx_val(v) = NULL; }
}

static struct custom_operations ctx_ops = {
    "org.biocaml.phylogenetics.phylo_b200.ctx", ctx_finalize, custom_compare_default, custom_hash_default,
    custom_serialize_default, custom_deserialize_default, custom_compare_ext_default, custom_fixed_length_default
};

/* PB_ERR_INVALID_ARG -> Invalid_argument, everything else -> Failure (the reference's conventions:
 * invalid_arg in lib/linear_algebra.ml:268,280,311, failwith in lib/rate_matrix.ml:48). */
static void check(pb_ctx *ctx, int rc)
{
    if (rc == PB_OK) return;
    if (rc == PB_ERR_INVALID_ARG) caml_invalid_argument(pb_last_error(ctx));
    caml_failwith(pb_last_error(ctx));
}

/* The C ABI carries no sizes: every Bigarray handed to a stub is measured against the context BEFORE its pointer is
 * passed on, and a short one is an Invalid_argument, never an out-of-bounds access. */
typedef struct { int64_t nstates, ncat, nsites, ntaxa, nnodes, flags; } dims_t;
static dims_t dims_of(pb_ctx *c)
{
    int64_t d[6] = {0, 0, 0, 0, 0, 0};
    dims_t r;
    pb_get_dims(c, d);
    r.nstates = d[0]; r.ncat = d[1]; r.nsites = d[2]; r.ntaxa = d[3]; r.nnodes = d[4]; r.flags = d[5];
    return r;
}
static void need(int cond, const char *msg)
{
    if (!cond) caml_invalid_argument(msg);
}
static int64_t len1(value ba) { return (int64_t)Caml_ba_array_val(ba)->dim[0]; }
static int is2(value ba, int64_t d0, int64_t d1)       /* a 2-dimensional Bigarray of at least d0 x d1 */
{
    struct caml_ba_array *b = Caml_ba_array_val(ba);
    return b->num_dims == 2 && (int64_t)b->dim[0] >= d0 && (int64_t)b->dim[1] >= d1;
}
static int is_square(value ba, int64_t n)
{
    struct caml_ba_array *b = Caml_ba_array_val(ba);
    return b->num_dims == 2 && (int64_t)b->dim[0] == n && (int64_t)b->dim[1] == n;
}

CAMLprim value pb_stub_create(value device, value nstates, value ncat, value nsites, value ntaxa, value flags)
{
    CAMLparam5(device, nstates, ncat, nsites, ntaxa);
    CAMLxparam1(flags);
    CAMLlocal1(v);
    pb_ctx *ctx = pb_create(Int_val(device), Int_val(nstates), Int_val(ncat), (int64_t)Long_val(nsites),
                            Int_val(ntaxa), (unsigned)Int_val(flags));
    if (!ctx) check(NULL, pb_last_status(NULL) ? pb_last_status(NULL) : PB_ERR_CUDA);   /* invalid sizes -> Invalid_argument */
    v = caml_alloc_custom(&ctx_ops, sizeof(pb_ctx *), 0, 1);
    Ctx_val(v) = ctx;
    CAMLreturn(v);
}
CAMLprim value pb_stub_create_bytecode(value *a, int n) { (void)n; return pb_stub_create(a[0], a[1], a[2], a[3], a[4], a[5]); }

CAMLprim value pb_stub_set_mode(value ctx, value mode)
{
    check(Ctx_val(ctx), pb_set_mode(Ctx_val(ctx), Int_val(mode)));
    return Val_unit;
}

CAMLprim value pb_stub_set_rescaling(value ctx, value threshold, value shift_at_root)
{
    check(Ctx_val(ctx), pb_set_rescaling(Ctx_val(ctx), Double_val(threshold), Bool_val(shift_at_root)));
    return Val_unit;
}

/* child_ptr, child_idx, tip_row: (int32, int32_elt, c_layout) Bigarray.Array1; branch_len: float64 Array1 */
CAMLprim value pb_stub_set_tree(value ctx, value root, value child_ptr, value child_idx, value tip_row, value blen)
{
    CAMLparam5(ctx, root, child_ptr, child_idx, tip_row);
    CAMLxparam1(blen);
    int nnodes = (int)Caml_ba_array_val(tip_row)->dim[0];
    need(nnodes >= 2 && len1(child_ptr) >= (int64_t)nnodes + 1 && len1(child_idx) >= (int64_t)nnodes - 1 &&
         len1(blen) >= nnodes, "Phylo_b200.set_tree: child_ptr needs nnodes + 1, child_idx nnodes - 1, branch_lengths nnodes entries");
    check(Ctx_val(ctx), pb_set_tree(Ctx_val(ctx), nnodes, Int_val(root), (const int32_t *)Caml_ba_data_val(child_ptr),
                                    (const int32_t *)Caml_ba_data_val(child_idx),
                                    (const int32_t *)Caml_ba_data_val(tip_row), (const double *)Caml_ba_data_val(blen)));
    CAMLreturn(Val_unit);
}
CAMLprim value pb_stub_set_tree_bytecode(value *a, int n) { (void)n; return pb_stub_set_tree(a[0], a[1], a[2], a[3], a[4], a[5]); }

CAMLprim value pb_stub_set_branch_lengths(value ctx, value blen)
{
    need(len1(blen) >= dims_of(Ctx_val(ctx)).nnodes, "Phylo_b200.set_branch_lengths: one entry per node");
    check(Ctx_val(ctx), pb_set_branch_lengths(Ctx_val(ctx), (const double *)Caml_ba_data_val(blen)));
    return Val_unit;
}

/* tips: (int, int8_unsigned_elt, c_layout) Bigarray.Array2 of shape [ntaxa][nsites] */
CAMLprim value pb_stub_set_tips(value ctx, value tips)
{
    CAMLparam2(ctx, tips);
    pb_ctx *c = Ctx_val(ctx);
    const uint8_t *p = (const uint8_t *)Caml_ba_data_val(tips);
    int64_t ld = Caml_ba_array_val(tips)->dim[1];
    int rc;
    need(is2(tips, dims_of(c).ntaxa, dims_of(c).nsites), "Phylo_b200.set_tips: the tip matrix must be ntaxa x nsites");
    caml_release_runtime_system();      /* 128 MB at 1M sites x 128 taxa: do not hold the lock */
    rc = pb_set_tips(c, p, ld);
    caml_acquire_runtime_system();
    check(c, rc);
    CAMLreturn(Val_unit);
}

/* tips: int8_unsigned Array2 [ntaxa][ceil(nsites/4)], nucleotides 2-bit packed (site 4b+j in bits 2j..2j+1 of byte b) */
CAMLprim value pb_stub_set_tips_packed2(value ctx, value tips)
{
    CAMLparam2(ctx, tips);
    pb_ctx *c = Ctx_val(ctx);
    const uint8_t *p = (const uint8_t *)Caml_ba_data_val(tips);
    int64_t ld = Caml_ba_array_val(tips)->dim[1];
    int rc;
    need(is2(tips, dims_of(c).ntaxa, (dims_of(c).nsites + 3) / 4), "Phylo_b200.set_tips_packed2: the matrix must be ntaxa x ceil(nsites / 4)");
    caml_release_runtime_system();
    rc = pb_set_tips_packed2(c, p, ld);
    caml_acquire_runtime_system();
    check(c, rc);
    CAMLreturn(Val_unit);
}

CAMLprim value pb_stub_set_root_frequencies(value ctx, value pi)
{
    need(len1(pi) >= dims_of(Ctx_val(ctx)).nstates, "Phylo_b200.set_root_frequencies: one entry per state");
    check(Ctx_val(ctx), pb_set_root_frequencies(Ctx_val(ctx), (const double *)Caml_ba_data_val(pi)));
    return Val_unit;
}

CAMLprim value pb_stub_set_categories(value ctx, value rates, value weights)
{
    need(len1(rates) >= dims_of(Ctx_val(ctx)).ncat && len1(weights) >= dims_of(Ctx_val(ctx)).ncat,
         "Phylo_b200.set_categories: one rate and one weight per category");
    check(Ctx_val(ctx), pb_set_categories(Ctx_val(ctx), (const double *)Caml_ba_data_val(rates),
                                          (const double *)Caml_ba_data_val(weights)));
    return Val_unit;
}

/* u, uinv: Lacaml.D.mat (Fortran layout = the ABI's column-major); lambda: Lacaml.D.vec */
CAMLprim value pb_stub_set_eigensystem(value ctx, value u, value lambda, value uinv)
{
    const int64_t n = dims_of(Ctx_val(ctx)).nstates;
    need(is_square(u, n) && is_square(uinv, n) && len1(lambda) >= n, "Phylo_b200.set_eigensystem: nstates x nstates matrices, nstates eigenvalues");
    check(Ctx_val(ctx), pb_set_eigensystem(Ctx_val(ctx), (const double *)Caml_ba_data_val(u),
                                           (const double *)Caml_ba_data_val(lambda),
                                           (const double *)Caml_ba_data_val(uinv)));
    return Val_unit;
}

CAMLprim value pb_stub_set_rate_matrix(value ctx, value q, value pi)
{
    const int64_t n = dims_of(Ctx_val(ctx)).nstates;
    need(is_square(q, n) && len1(pi) >= n, "Phylo_b200.set_rate_matrix: an nstates x nstates matrix and nstates frequencies");
    check(Ctx_val(ctx), pb_set_rate_matrix(Ctx_val(ctx), (const double *)Caml_ba_data_val(q),
                                           (const double *)Caml_ba_data_val(pi)));
    return Val_unit;
}

CAMLprim value pb_stub_set_rate_matrix_general(value ctx, value q)
{
    need(is_square(q, dims_of(Ctx_val(ctx)).nstates), "Phylo_b200.set_rate_matrix_general: an nstates x nstates matrix");
    check(Ctx_val(ctx), pb_set_rate_matrix_general(Ctx_val(ctx), (const double *)Caml_ba_data_val(q)));
    return Val_unit;
}

/* p: float64 Bigarray.Array1 holding [ncat][nnodes] column-major n x n matrices back to back */
CAMLprim value pb_stub_set_pmatrices(value ctx, value p)
{
    const dims_t d = dims_of(Ctx_val(ctx));
    need(d.nnodes > 0 && len1(p) >= d.ncat * d.nnodes * d.nstates * d.nstates,
         "Phylo_b200.set_pmatrices: set the tree first; ncat * nnodes matrices of nstates x nstates");
    check(Ctx_val(ctx), pb_set_pmatrices(Ctx_val(ctx), (const double *)Caml_ba_data_val(p)));
    return Val_unit;
}

/* per_site: float64 Array1 of length nsites, or an empty one to skip the per-site copy */
CAMLprim value pb_stub_loglik(value ctx, value per_site)
{
    CAMLparam2(ctx, per_site);
    pb_ctx *c = Ctx_val(ctx);
    double total = 0.0;
    double *ps = Caml_ba_array_val(per_site)->dim[0] > 0 ? (double *)Caml_ba_data_val(per_site) : NULL;
    int rc;
    need(!ps || len1(per_site) >= dims_of(c).nsites, "Phylo_b200.loglik: per_site must be empty or hold nsites values");
    caml_release_runtime_system();
    rc = pb_loglik(c, &total, ps);
    caml_acquire_runtime_system();
    check(c, rc);
    CAMLreturn(caml_copy_double(total));
}

/* v: float64 Array2 [nsites][nstates] (c_layout); carry: float64 Array1 [nsites] */
CAMLprim value pb_stub_get_clv(value ctx, value node, value cat, value v, value carry)
{
    const dims_t d = dims_of(Ctx_val(ctx));
    need(is2(v, d.nsites, d.nstates) && len1(carry) >= d.nsites, "Phylo_b200.conditional_likelihood: v must be nsites x nstates, carry nsites");
    check(Ctx_val(ctx), pb_get_clv(Ctx_val(ctx), Int_val(node), Int_val(cat), (double *)Caml_ba_data_val(v),
                                   (double *)Caml_ba_data_val(carry)));
    return Val_unit;
}

/* states: int8_unsigned Array2 [nnodes][nsites] (c_layout); seed: Int64 drawn by the caller from its Gsl.Rng.t */
CAMLprim value pb_stub_conditional_simulation(value ctx, value seed, value states)
{
    const dims_t d = dims_of(Ctx_val(ctx));
    need(d.nnodes > 0 && is2(states, d.nnodes, d.nsites), "Phylo_b200.conditional_simulation: states must be nnodes x nsites");
    check(Ctx_val(ctx), pb_conditional_simulation(Ctx_val(ctx), (uint64_t)Int64_val(seed),
                                                  (uint8_t *)Caml_ba_data_val(states),
                                                  (int64_t)Caml_ba_array_val(states)->dim[1]));
    return Val_unit;
}

/* Stand-alone P(t): u, uinv Lacaml mats; lambda, ts vecs; out: float64 Array1 of count*n*n (column-major each) */
CAMLprim value pb_stub_transition_matrices_eigen(value device, value u, value lambda, value uinv, value ts, value out)
{
    int n = (int)Caml_ba_array_val(lambda)->dim[0], count = (int)Caml_ba_array_val(ts)->dim[0];
    need(is_square(u, n) && is_square(uinv, n) && len1(out) >= (int64_t)count * n * n,
         "Phylo_b200.transition_matrices_eigen: n x n matrices and room for count * n * n results");
    int rc = pb_transition_matrices_eigen(Int_val(device), n, count, (const double *)Caml_ba_data_val(u),
                                          (const double *)Caml_ba_data_val(lambda),
                                          (const double *)Caml_ba_data_val(uinv),
                                          (const double *)Caml_ba_data_val(ts), (double *)Caml_ba_data_val(out));
    check(NULL, rc);
    return Val_unit;
}
CAMLprim value pb_stub_transition_matrices_eigen_bytecode(value *a, int n)
{
    (void)n;
    return pb_stub_transition_matrices_eigen(a[0], a[1], a[2], a[3], a[4], a[5]);
}

CAMLprim value pb_stub_transition_matrices_expm(value device, value q, value ts, value out)
{
    int n = (int)Caml_ba_array_val(q)->dim[0], count = (int)Caml_ba_array_val(ts)->dim[0];
    need(is_square(q, n) && len1(out) >= (int64_t)count * n * n,
         "Phylo_b200.transition_matrices_expm: a square matrix and room for count * n * n results");
    check(NULL, pb_transition_matrices_expm(Int_val(device), n, count, (const double *)Caml_ba_data_val(q),
                                            (const double *)Caml_ba_data_val(ts), (double *)Caml_ba_data_val(out)));
    return Val_unit;
}

/* ---- the rest of include/phylo_b200.h ---------------------------------------------------------------- */

/* masks: (int64, int64_elt, c_layout) Array2 [ntaxa][nsites]; bit i set = state i compatible, 0 = missing
 * (Phylo_ctmc.Missing_values / Ambiguous, lib/phylo_ctmc.ml:145-172, 232-302) */
CAMLprim value pb_stub_set_tip_masks(value ctx, value masks)
{
    need(is2(masks, dims_of(Ctx_val(ctx)).ntaxa, dims_of(Ctx_val(ctx)).nsites), "Phylo_b200.set_tip_masks: the mask matrix must be ntaxa x nsites");
    check(Ctx_val(ctx), pb_set_tip_masks(Ctx_val(ctx), (const uint64_t *)Caml_ba_data_val(masks),
                                         (int64_t)Caml_ba_array_val(masks)->dim[1]));
    return Val_unit;
}

/* text: (char, int8_unsigned_elt, c_layout) Array2 [ntaxa][nsites * width]; alphabet: PB_ALPHABET_* */
CAMLprim value pb_stub_set_tips_text(value ctx, value text, value alphabet)
{
    need(is2(text, dims_of(Ctx_val(ctx)).ntaxa, dims_of(Ctx_val(ctx)).nsites * (Int_val(alphabet) == PB_ALPHABET_CODON ? 3 : 1)),
         "Phylo_b200.set_tips_text: ntaxa rows of nsites symbols");
    check(Ctx_val(ctx), pb_set_tips_text(Ctx_val(ctx), (const char *)Caml_ba_data_val(text),
                                         (int64_t)Caml_ba_array_val(text)->dim[1], Int_val(alphabet)));
    return Val_unit;
}

/* w: float64 Array1 [nsites], or an empty one to remove the weights */
CAMLprim value pb_stub_set_site_weights(value ctx, value w)
{
    const double *p = Caml_ba_array_val(w)->dim[0] > 0 ? (const double *)Caml_ba_data_val(w) : NULL;
    need(!p || len1(w) >= dims_of(Ctx_val(ctx)).nsites, "Phylo_b200.set_site_weights: empty, or one weight per site");
    check(Ctx_val(ctx), pb_set_site_weights(Ctx_val(ctx), p));
    return Val_unit;
}

CAMLprim value pb_stub_update_branch_length(value ctx, value node, value t)
{
    check(Ctx_val(ctx), pb_update_branch_length(Ctx_val(ctx), Int_val(node), Double_val(t)));
    return Val_unit;
}

CAMLprim value pb_stub_nodes_recomputed(value ctx) { return Val_int(pb_nodes_recomputed(Ctx_val(ctx))); }

CAMLprim value pb_stub_set_option(value ctx, value name, value v)
{
    check(Ctx_val(ctx), pb_set_option(Ctx_val(ctx), String_val(name), Double_val(v)));
    return Val_unit;
}

/* p: float64 Array1 of ncat * nnodes * n * n */
CAMLprim value pb_stub_get_pmatrices(value ctx, value p)
{
    const dims_t d = dims_of(Ctx_val(ctx));
    need(d.nnodes > 0 && len1(p) >= d.ncat * d.nnodes * d.nstates * d.nstates, "Phylo_b200.get_pmatrices: room for ncat * nnodes matrices");
    check(Ctx_val(ctx), pb_get_pmatrices(Ctx_val(ctx), (double *)Caml_ba_data_val(p)));
    return Val_unit;
}

CAMLprim value pb_stub_path_name(value ctx) { return caml_copy_string(pb_path_name(Ctx_val(ctx))); }
CAMLprim value pb_stub_kernel_launches(value ctx) { return Val_long((long)pb_kernel_launches(Ctx_val(ctx))); }

/* One-call evaluation from a host alignment, copy and pruning overlapped (pb_loglik_host / _packed2): the calls the
 * end-to-end throughput is quoted on.  tips as in pb_stub_set_tips / pb_stub_set_tips_packed2; pin the Bigarray once
 * with pb_stub_pin for the copies to be asynchronous. */
static value loglik_host(value ctx, value tips, value per_site, int packed2)
{
    CAMLparam3(ctx, tips, per_site);
    pb_ctx *c = Ctx_val(ctx);
    const uint8_t *p = (const uint8_t *)Caml_ba_data_val(tips);
    int64_t ld = Caml_ba_array_val(tips)->dim[1];
    double *ps = Caml_ba_array_val(per_site)->dim[0] > 0 ? (double *)Caml_ba_data_val(per_site) : NULL;
    double total = 0.0;
    int rc;
    need(is2(tips, dims_of(c).ntaxa, packed2 ? (dims_of(c).nsites + 3) / 4 : dims_of(c).nsites),
         "Phylo_b200.loglik_host: the tip matrix must be ntaxa x nsites (ntaxa x ceil(nsites / 4) when packed)");
    need(!ps || len1(per_site) >= dims_of(c).nsites, "Phylo_b200.loglik_host: per_site must be empty or hold nsites values");
    caml_release_runtime_system();
    rc = packed2 ? pb_loglik_host_packed2(c, p, ld, &total, ps) : pb_loglik_host(c, p, ld, &total, ps);
    caml_acquire_runtime_system();
    check(c, rc);
    CAMLreturn(caml_copy_double(total));
}
CAMLprim value pb_stub_loglik_host(value ctx, value tips, value per_site) { return loglik_host(ctx, tips, per_site, 0); }
CAMLprim value pb_stub_loglik_host_packed2(value ctx, value tips, value per_site) { return loglik_host(ctx, tips, per_site, 1); }

/* Page-lock / release the payload of a Bigarray (it lives outside the OCaml heap and does not move). */
CAMLprim value pb_stub_pin(value ba)
{
    struct caml_ba_array *b = Caml_ba_array_val(ba);
    size_t bytes = caml_ba_byte_size(b);
    check(NULL, pb_pin_host_memory(b->data, bytes));
    return Val_unit;
}
CAMLprim value pb_stub_unpin(value ba)
{
    check(NULL, pb_unpin_host_memory(Caml_ba_data_val(ba)));
    return Val_unit;
}

CAMLprim value pb_stub_get_dims(value ctx)
{
    CAMLparam1(ctx);
    CAMLlocal1(t);
    const dims_t d = dims_of(Ctx_val(ctx));
    t = caml_alloc_tuple(6);
    Store_field(t, 0, Val_long(d.nstates)); Store_field(t, 1, Val_long(d.ncat)); Store_field(t, 2, Val_long(d.nsites));
    Store_field(t, 3, Val_long(d.ntaxa)); Store_field(t, 4, Val_long(d.nnodes)); Store_field(t, 5, Val_long(d.flags));
    CAMLreturn(t);
}

/* ---- multi-GPU: pb_create_sharded (one host thread, one context per device, NCCL all-reduce of the scalar) ---- */

#define Sharded_val(v) (*((pb_sharded **)Data_custom_val(v)))

static void sharded_finalize(value v)
{
    if (Sharded_val(v)) { pb_sharded_destroy(Sharded_val(v)); Sharded_val(v) = NULL; }
}

static struct custom_operations sharded_ops = {
    "org.biocaml.phylogenetics.phylo_b200.sharded", sharded_finalize, custom_compare_default, custom_hash_default,
    custom_serialize_default, custom_deserialize_default, custom_compare_ext_default, custom_fixed_length_default
};

static void scheck(pb_sharded *s, int rc)
{
    if (rc == PB_OK) return;
    if (rc == PB_ERR_INVALID_ARG) caml_invalid_argument(pb_sharded_last_error(s));
    caml_failwith(pb_sharded_last_error(s));             /* PB_ERR_NCCL, PB_ERR_CUDA, ... */
}

/* devices: (int32, int32_elt, c_layout) Array1 */
CAMLprim value pb_stub_create_sharded(value devices, value nstates, value ncat, value nsites, value ntaxa, value flags)
{
    CAMLparam5(devices, nstates, ncat, nsites, ntaxa);
    CAMLxparam1(flags);
    CAMLlocal1(v);
    pb_sharded *s = pb_create_sharded((int)len1(devices), (const int *)Caml_ba_data_val(devices), Int_val(nstates), Int_val(ncat),
                                      (int64_t)Long_val(nsites), Int_val(ntaxa), (unsigned)Int_val(flags));
    if (!s) {
        const char *msg = pb_sharded_last_error(NULL);
        if (strstr(msg, "pb_create_sharded: ndev") || strstr(msg, "more devices than sites") || strstr(msg, "must be"))
            caml_invalid_argument(msg);
        caml_failwith(msg);
    }
    v = caml_alloc_custom(&sharded_ops, sizeof(pb_sharded *), 0, 1);
    Sharded_val(v) = s;
    CAMLreturn(v);
}
CAMLprim value pb_stub_create_sharded_bytecode(value *a, int n) { (void)n; return pb_stub_create_sharded(a[0], a[1], a[2], a[3], a[4], a[5]); }

CAMLprim value pb_stub_sharded_ndev(value s) { return Val_int(pb_sharded_ndev(Sharded_val(s))); }

CAMLprim value pb_stub_sharded_set_reduce(value s, value how)
{
    scheck(Sharded_val(s), pb_sharded_set_reduce(Sharded_val(s), Int_val(how)));
    return Val_unit;
}

CAMLprim value pb_stub_sharded_set_rescaling(value s, value threshold, value shift_at_root)
{
    pb_sharded *sh = Sharded_val(s);
    for (int d = 0; d < pb_sharded_ndev(sh); d++)
        check(pb_sharded_ctx(sh, d), pb_set_rescaling(pb_sharded_ctx(sh, d), Double_val(threshold), Bool_val(shift_at_root)));
    return Val_unit;
}

CAMLprim value pb_stub_sharded_set_tree(value s, value root, value child_ptr, value child_idx, value tip_row, value blen)
{
    CAMLparam5(s, root, child_ptr, child_idx, tip_row);
    CAMLxparam1(blen);
    int nnodes = (int)len1(tip_row);
    need(nnodes >= 2 && len1(child_ptr) >= (int64_t)nnodes + 1 && len1(child_idx) >= (int64_t)nnodes - 1 && len1(blen) >= nnodes,
         "Phylo_b200.Sharded.set_tree: child_ptr needs nnodes + 1, child_idx nnodes - 1, branch_lengths nnodes entries");
    scheck(Sharded_val(s), pb_sharded_set_tree(Sharded_val(s), nnodes, Int_val(root), (const int32_t *)Caml_ba_data_val(child_ptr),
                                               (const int32_t *)Caml_ba_data_val(child_idx),
                                               (const int32_t *)Caml_ba_data_val(tip_row), (const double *)Caml_ba_data_val(blen)));
    CAMLreturn(Val_unit);
}
CAMLprim value pb_stub_sharded_set_tree_bytecode(value *a, int n) { (void)n; return pb_stub_sharded_set_tree(a[0], a[1], a[2], a[3], a[4], a[5]); }

CAMLprim value pb_stub_sharded_set_branch_lengths(value s, value blen)
{
    need(len1(blen) >= dims_of(pb_sharded_ctx(Sharded_val(s), 0)).nnodes, "Phylo_b200.Sharded.set_branch_lengths: one entry per node");
    scheck(Sharded_val(s), pb_sharded_set_branch_lengths(Sharded_val(s), (const double *)Caml_ba_data_val(blen)));
    return Val_unit;
}

CAMLprim value pb_stub_sharded_set_root_frequencies(value s, value pi)
{
    need(len1(pi) >= dims_of(pb_sharded_ctx(Sharded_val(s), 0)).nstates, "Phylo_b200.Sharded.set_root_frequencies: one entry per state");
    scheck(Sharded_val(s), pb_sharded_set_root_frequencies(Sharded_val(s), (const double *)Caml_ba_data_val(pi)));
    return Val_unit;
}

CAMLprim value pb_stub_sharded_set_categories(value s, value rates, value weights)
{
    const int64_t ncat = dims_of(pb_sharded_ctx(Sharded_val(s), 0)).ncat;
    need(len1(rates) >= ncat && len1(weights) >= ncat, "Phylo_b200.Sharded.set_categories: one rate and one weight per category");
    scheck(Sharded_val(s), pb_sharded_set_categories(Sharded_val(s), (const double *)Caml_ba_data_val(rates),
                                                     (const double *)Caml_ba_data_val(weights)));
    return Val_unit;
}

CAMLprim value pb_stub_sharded_set_eigensystem(value s, value u, value lambda, value uinv)
{
    const int64_t n = dims_of(pb_sharded_ctx(Sharded_val(s), 0)).nstates;
    need(is_square(u, n) && is_square(uinv, n) && len1(lambda) >= n, "Phylo_b200.Sharded.set_eigensystem: nstates x nstates matrices, nstates eigenvalues");
    scheck(Sharded_val(s), pb_sharded_set_eigensystem(Sharded_val(s), (const double *)Caml_ba_data_val(u),
                                                      (const double *)Caml_ba_data_val(lambda), (const double *)Caml_ba_data_val(uinv)));
    return Val_unit;
}

CAMLprim value pb_stub_sharded_set_rate_matrix(value s, value q, value pi)
{
    const int64_t n = dims_of(pb_sharded_ctx(Sharded_val(s), 0)).nstates;
    need(is_square(q, n) && len1(pi) >= n, "Phylo_b200.Sharded.set_rate_matrix: an nstates x nstates matrix and nstates frequencies");
    scheck(Sharded_val(s), pb_sharded_set_rate_matrix(Sharded_val(s), (const double *)Caml_ba_data_val(q), (const double *)Caml_ba_data_val(pi)));
    return Val_unit;
}

CAMLprim value pb_stub_sharded_set_pmatrices(value s, value p)
{
    const dims_t d = dims_of(pb_sharded_ctx(Sharded_val(s), 0));
    need(d.nnodes > 0 && len1(p) >= d.ncat * d.nnodes * d.nstates * d.nstates,
         "Phylo_b200.Sharded.set_pmatrices: set the tree first; ncat * nnodes matrices of nstates x nstates");
    scheck(Sharded_val(s), pb_sharded_set_pmatrices(Sharded_val(s), (const double *)Caml_ba_data_val(p)));
    return Val_unit;
}

static int64_t sharded_nsites(pb_sharded *s)
{
    int64_t lo = 0, hi = 0;
    pb_sharded_shard_bounds(s, pb_sharded_ndev(s) - 1, &lo, &hi);
    return hi;
}

/* tips: the FULL alignment, int8_unsigned Array2 [ntaxa][nsites_total]; every device copies its own columns */
CAMLprim value pb_stub_sharded_set_tips(value s, value tips)
{
    CAMLparam2(s, tips);
    pb_sharded *sh = Sharded_val(s);
    int rc;
    need(is2(tips, dims_of(pb_sharded_ctx(sh, 0)).ntaxa, sharded_nsites(sh)), "Phylo_b200.Sharded.set_tips: the tip matrix must be ntaxa x nsites");
    caml_release_runtime_system();
    rc = pb_sharded_set_tips(sh, (const uint8_t *)Caml_ba_data_val(tips), (int64_t)Caml_ba_array_val(tips)->dim[1]);
    caml_acquire_runtime_system();
    scheck(sh, rc);
    CAMLreturn(Val_unit);
}

/* per_site: float64 Array1 of nsites_total, or empty */
CAMLprim value pb_stub_sharded_loglik(value s, value per_site)
{
    CAMLparam2(s, per_site);
    pb_sharded *sh = Sharded_val(s);
    double total = 0.0;
    double *ps = len1(per_site) > 0 ? (double *)Caml_ba_data_val(per_site) : NULL;
    int rc;
    need(!ps || len1(per_site) >= sharded_nsites(sh), "Phylo_b200.Sharded.loglik: per_site must be empty or hold nsites values");
    caml_release_runtime_system();
    rc = pb_sharded_loglik(sh, &total, ps);
    caml_acquire_runtime_system();
    scheck(sh, rc);
    CAMLreturn(caml_copy_double(total));
}

/* one call per evaluation from a host alignment (packed2: ntaxa x ceil(nsites / 4)) */
CAMLprim value pb_stub_sharded_loglik_host(value s, value tips, value packed2, value per_site)
{
    CAMLparam4(s, tips, packed2, per_site);
    pb_sharded *sh = Sharded_val(s);
    const int pk = Bool_val(packed2);
    double total = 0.0;
    double *ps = len1(per_site) > 0 ? (double *)Caml_ba_data_val(per_site) : NULL;
    int rc;
    need(is2(tips, dims_of(pb_sharded_ctx(sh, 0)).ntaxa, pk ? (sharded_nsites(sh) + 3) / 4 : sharded_nsites(sh)),
         "Phylo_b200.Sharded.loglik_host: the tip matrix must be ntaxa x nsites (ntaxa x ceil(nsites / 4) when packed)");
    need(!ps || len1(per_site) >= sharded_nsites(sh), "Phylo_b200.Sharded.loglik_host: per_site must be empty or hold nsites values");
    caml_release_runtime_system();
    rc = pb_sharded_loglik_host(sh, (const uint8_t *)Caml_ba_data_val(tips), (int64_t)Caml_ba_array_val(tips)->dim[1], pk, &total, ps);
    caml_acquire_runtime_system();
    scheck(sh, rc);
    CAMLreturn(caml_copy_double(total));
}
