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

CAMLprim value pb_stub_create(value device, value nstates, value ncat, value nsites, value ntaxa, value flags)
{
    CAMLparam5(device, nstates, ncat, nsites, ntaxa);
    CAMLxparam1(flags);
    CAMLlocal1(v);
    pb_ctx *ctx = pb_create(Int_val(device), Int_val(nstates), Int_val(ncat), (int64_t)Long_val(nsites),
                            Int_val(ntaxa), (unsigned)Int_val(flags));
    if (!ctx) caml_failwith(pb_last_error(NULL));
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
    check(Ctx_val(ctx), pb_set_tree(Ctx_val(ctx), nnodes, Int_val(root), (const int32_t *)Caml_ba_data_val(child_ptr),
                                    (const int32_t *)Caml_ba_data_val(child_idx),
                                    (const int32_t *)Caml_ba_data_val(tip_row), (const double *)Caml_ba_data_val(blen)));
    CAMLreturn(Val_unit);
}
CAMLprim value pb_stub_set_tree_bytecode(value *a, int n) { (void)n; return pb_stub_set_tree(a[0], a[1], a[2], a[3], a[4], a[5]); }

CAMLprim value pb_stub_set_branch_lengths(value ctx, value blen)
{
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
    caml_release_runtime_system();
    rc = pb_set_tips_packed2(c, p, ld);
    caml_acquire_runtime_system();
    check(c, rc);
    CAMLreturn(Val_unit);
}

CAMLprim value pb_stub_set_root_frequencies(value ctx, value pi)
{
    check(Ctx_val(ctx), pb_set_root_frequencies(Ctx_val(ctx), (const double *)Caml_ba_data_val(pi)));
    return Val_unit;
}

CAMLprim value pb_stub_set_categories(value ctx, value rates, value weights)
{
    check(Ctx_val(ctx), pb_set_categories(Ctx_val(ctx), (const double *)Caml_ba_data_val(rates),
                                          (const double *)Caml_ba_data_val(weights)));
    return Val_unit;
}

/* u, uinv: Lacaml.D.mat (Fortran layout = the ABI's column-major); lambda: Lacaml.D.vec */
CAMLprim value pb_stub_set_eigensystem(value ctx, value u, value lambda, value uinv)
{
    check(Ctx_val(ctx), pb_set_eigensystem(Ctx_val(ctx), (const double *)Caml_ba_data_val(u),
                                           (const double *)Caml_ba_data_val(lambda),
                                           (const double *)Caml_ba_data_val(uinv)));
    return Val_unit;
}

CAMLprim value pb_stub_set_rate_matrix(value ctx, value q, value pi)
{
    check(Ctx_val(ctx), pb_set_rate_matrix(Ctx_val(ctx), (const double *)Caml_ba_data_val(q),
                                           (const double *)Caml_ba_data_val(pi)));
    return Val_unit;
}

CAMLprim value pb_stub_set_rate_matrix_general(value ctx, value q)
{
    check(Ctx_val(ctx), pb_set_rate_matrix_general(Ctx_val(ctx), (const double *)Caml_ba_data_val(q)));
    return Val_unit;
}

/* p: float64 Bigarray.Array1 holding [ncat][nnodes] column-major n x n matrices back to back */
CAMLprim value pb_stub_set_pmatrices(value ctx, value p)
{
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
    caml_release_runtime_system();
    rc = pb_loglik(c, &total, ps);
    caml_acquire_runtime_system();
    check(c, rc);
    CAMLreturn(caml_copy_double(total));
}

/* v: float64 Array2 [nsites][nstates] (c_layout); carry: float64 Array1 [nsites] */
CAMLprim value pb_stub_get_clv(value ctx, value node, value cat, value v, value carry)
{
    check(Ctx_val(ctx), pb_get_clv(Ctx_val(ctx), Int_val(node), Int_val(cat), (double *)Caml_ba_data_val(v),
                                   (double *)Caml_ba_data_val(carry)));
    return Val_unit;
}

/* states: int8_unsigned Array2 [nnodes][nsites] (c_layout); seed: Int64 drawn by the caller from its Gsl.Rng.t */
CAMLprim value pb_stub_conditional_simulation(value ctx, value seed, value states)
{
    check(Ctx_val(ctx), pb_conditional_simulation(Ctx_val(ctx), (uint64_t)Int64_val(seed),
                                                  (uint8_t *)Caml_ba_data_val(states),
                                                  (int64_t)Caml_ba_array_val(states)->dim[1]));
    return Val_unit;
}

/* Stand-alone P(t): u, uinv Lacaml mats; lambda, ts vecs; out: float64 Array1 of count*n*n (column-major each) */
CAMLprim value pb_stub_transition_matrices_eigen(value device, value u, value lambda, value uinv, value ts, value out)
{
    int n = (int)Caml_ba_array_val(lambda)->dim[0], count = (int)Caml_ba_array_val(ts)->dim[0];
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
    check(NULL, pb_transition_matrices_expm(Int_val(device), n, count, (const double *)Caml_ba_data_val(q),
                                            (const double *)Caml_ba_data_val(ts), (double *)Caml_ba_data_val(out)));
    return Val_unit;
}

/* ---- the rest of include/phylo_b200.h ---------------------------------------------------------------- */

/* masks: (int64, int64_elt, c_layout) Array2 [ntaxa][nsites]; bit i set = state i compatible, 0 = missing
 * (Phylo_ctmc.Missing_values / Ambiguous, lib/phylo_ctmc.ml:145-172, 232-302) */
CAMLprim value pb_stub_set_tip_masks(value ctx, value masks)
{
    check(Ctx_val(ctx), pb_set_tip_masks(Ctx_val(ctx), (const uint64_t *)Caml_ba_data_val(masks),
                                         (int64_t)Caml_ba_array_val(masks)->dim[1]));
    return Val_unit;
}

/* text: (char, int8_unsigned_elt, c_layout) Array2 [ntaxa][nsites * width]; alphabet: PB_ALPHABET_* */
CAMLprim value pb_stub_set_tips_text(value ctx, value text, value alphabet)
{
    check(Ctx_val(ctx), pb_set_tips_text(Ctx_val(ctx), (const char *)Caml_ba_data_val(text),
                                         (int64_t)Caml_ba_array_val(text)->dim[1], Int_val(alphabet)));
    return Val_unit;
}

/* w: float64 Array1 [nsites], or an empty one to remove the weights */
CAMLprim value pb_stub_set_site_weights(value ctx, value w)
{
    const double *p = Caml_ba_array_val(w)->dim[0] > 0 ? (const double *)Caml_ba_data_val(w) : NULL;
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
