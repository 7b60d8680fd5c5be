/*
 * Executes ocaml/phylo_b200_stubs.c on the GPU against a MOCK of the OCaml runtime (tests/ocaml_mock/caml/):
 * test infrastructure only.  It plays the part of ocaml/phylo_b200.ml — builds the Bigarray descriptors the
 * OCaml side would pass (c_layout int32 / float64 / int8_unsigned arrays, Fortran-layout Lacaml matrices),
 * calls the stubs in the order `pruning_sites` and `Felsenstein(..).evaluate` call them, and checks
 *   - results against the RNG-free fixtures of the reference (lib/mCMC.ml:22,36-38) and closed forms
 *     (the same values as tests/golden/known_answers.json and tests/cpp/host_mirror_test.cpp),
 *   - the error mapping: PB_ERR_INVALID_ARG -> Invalid_argument, anything else -> Failure,
 *   - that the runtime lock is released and re-acquired in pairs and never held released across a raise,
 *   - that the custom block's finaliser destroys the context exactly once.
 * The real OCaml runtime is not involved; what this proves is that the C code of the stubs is correct with
 * respect to include/phylo_b200.h and behaves on a device.
 */
#include <math.h>
#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <caml/threads.h>

#include "phylo_b200.h"

/* ---- the mock runtime ------------------------------------------------------------------------------- */

enum { EXN_NONE = 0, EXN_FAILURE = 1, EXN_INVALID_ARGUMENT = 2 };
static jmp_buf exn_jmp;
static int exn_armed = 0, exn_kind = EXN_NONE;
static char exn_msg[512];
static int lock_released = 0, lock_releases = 0, lock_errors = 0, finalised = 0;

static void raise_exn(int kind, const char *msg) __attribute__((noreturn));
static void raise_exn(int kind, const char *msg)
{
    if (lock_released) { lock_errors++; lock_released = 0; }   /* raising without the runtime lock: fatal in OCaml */
    exn_kind = kind;
    snprintf(exn_msg, sizeof exn_msg, "%s", msg ? msg : "(null)");
    if (!exn_armed) { fprintf(stderr, "uncaught mock exception %d: %s\n", kind, exn_msg); exit(2); }
    longjmp(exn_jmp, 1);
}
void caml_failwith(const char *msg) { raise_exn(EXN_FAILURE, msg); }
void caml_invalid_argument(const char *msg) { raise_exn(EXN_INVALID_ARGUMENT, msg); }
void caml_release_runtime_system(void) { if (lock_released) lock_errors++; lock_released = 1; lock_releases++; }
void caml_acquire_runtime_system(void) { if (!lock_released) lock_errors++; lock_released = 0; }

value caml_copy_double(double d)
{
    double *p = malloc(sizeof *p);
    *p = d;
    return (value)p;
}
value caml_copy_int64(int64_t i)
{
    int64_t *p = malloc(sizeof *p);
    *p = i;
    return (value)p;
}
value caml_copy_string(const char *str)
{
    char *p = malloc(strlen(str) + 1);
    strcpy(p, str);
    return (value)p;
}
uintnat caml_ba_byte_size(struct caml_ba_array *b)
{
    static const int elt[] = {4, 8, 1, 1, 2, 2, 4, 8};          /* enum caml_ba_kind order */
    uintnat n = elt[b->flags & CAML_BA_KIND_MASK];
    for (intnat i = 0; i < b->num_dims; i++) n *= (uintnat)b->dim[i];
    return n;
}
value caml_alloc_tuple(uintnat n) { return (value)calloc(n, sizeof(value)); }
value caml_alloc_custom(struct custom_operations *ops, uintnat size, uintnat mem, uintnat max)
{
    (void)mem; (void)max;
    struct caml_mock_custom *b = calloc(1, sizeof *b + size);
    b->ops = ops;
    return (value)b;
}

/* what `Bigarray.Array1.create` / `Array2.create` / a Lacaml matrix look like from C */
static value ba1(void *data, intnat n, int kind)
{
    struct caml_ba_array *b = calloc(1, sizeof *b);
    b->data = data; b->num_dims = 1; b->flags = kind | CAML_BA_C_LAYOUT; b->dim[0] = n;
    return (value)b;
}
static value ba2(void *data, intnat d0, intnat d1, int kind, int layout)
{
    struct caml_ba_array *b = calloc(1, sizeof *b);
    b->data = data; b->num_dims = 2; b->flags = kind | layout; b->dim[0] = d0; b->dim[1] = d1;
    return (value)b;
}

/* ---- the stubs under test (ocaml/phylo_b200_stubs.c) ------------------------------------------------ */

value pb_stub_create(value, value, value, value, value, value);
value pb_stub_create_bytecode(value *, int);
value pb_stub_set_mode(value, value);
value pb_stub_set_rescaling(value, value, value);
value pb_stub_set_tree(value, value, value, value, value, value);
value pb_stub_set_tree_bytecode(value *, int);
value pb_stub_set_branch_lengths(value, value);
value pb_stub_set_tips(value, value);
value pb_stub_set_tips_packed2(value, value);
value pb_stub_set_root_frequencies(value, value);
value pb_stub_set_categories(value, value, value);
value pb_stub_set_eigensystem(value, value, value, value);
value pb_stub_set_rate_matrix(value, value, value);
value pb_stub_set_rate_matrix_general(value, value);
value pb_stub_set_pmatrices(value, value);
value pb_stub_loglik(value, value);
value pb_stub_get_clv(value, value, value, value, value);
value pb_stub_conditional_simulation(value, value, value);
value pb_stub_transition_matrices_eigen(value, value, value, value, value, value);
value pb_stub_transition_matrices_eigen_bytecode(value *, int);
value pb_stub_transition_matrices_expm(value, value, value, value);
value pb_stub_set_tip_masks(value, value);
value pb_stub_set_tips_text(value, value, value);
value pb_stub_set_site_weights(value, value);
value pb_stub_update_branch_length(value, value, value);
value pb_stub_nodes_recomputed(value);
value pb_stub_set_option(value, value, value);
value pb_stub_get_pmatrices(value, value);
value pb_stub_path_name(value);
value pb_stub_kernel_launches(value);
value pb_stub_loglik_host(value, value, value);
value pb_stub_loglik_host_packed2(value, value, value);
value pb_stub_pin(value);
value pb_stub_unpin(value);
value pb_stub_get_dims(value);
value pb_stub_create_sharded(value, value, value, value, value, value);
value pb_stub_create_sharded_bytecode(value *, int);
value pb_stub_sharded_ndev(value);
value pb_stub_sharded_set_reduce(value, value);
value pb_stub_sharded_set_rescaling(value, value, value);
value pb_stub_sharded_set_tree(value, value, value, value, value, value);
value pb_stub_sharded_set_tree_bytecode(value *, int);
value pb_stub_sharded_set_branch_lengths(value, value);
value pb_stub_sharded_set_root_frequencies(value, value);
value pb_stub_sharded_set_categories(value, value, value);
value pb_stub_sharded_set_eigensystem(value, value, value, value);
value pb_stub_sharded_set_rate_matrix(value, value, value);
value pb_stub_sharded_set_pmatrices(value, value);
value pb_stub_sharded_set_tips(value, value);
value pb_stub_sharded_loglik(value, value);
value pb_stub_sharded_loglik_host(value, value, value, value);

/* ---- checks ----------------------------------------------------------------------------------------- */

static int failures = 0;
static void expect_close(const char *what, double got, double want, double tol)
{
    const int ok = fabs(got - want) <= tol * fmax(fabs(want), 1e-300);
    printf("%-66s got %.15g want %.15g %s\n", what, got, want, ok ? "ok" : "FAIL");
    if (!ok) failures++;
}
static void expect_true(const char *what, int cond)
{
    printf("%-66s %s\n", what, cond ? "ok" : "FAIL");
    if (!cond) failures++;
}
/* runs `call`, expecting the mock exception `kind` */
#define EXPECT_RAISES(what, kind, call)                                                        \
    do {                                                                                       \
        exn_kind = EXN_NONE; exn_armed = 1;                                                    \
        if (!setjmp(exn_jmp)) { (void)(call); }                                                \
        exn_armed = 0;                                                                         \
        printf("%-66s %s (%s)\n", what, exn_kind == (kind) ? "ok" : "FAIL", exn_kind ? exn_msg : "no exception"); \
        if (exn_kind != (kind)) failures++;                                                    \
    } while (0)

/* the tree ((L0:a,L1:b):c,(L2:d,L3:e):f) numbered in pre-order as Phylo_b200.flatten does */
static int32_t child_ptr[8] = {0, 2, 4, 4, 4, 6, 6, 6};
static int32_t child_idx[6] = {1, 4, 2, 3, 5, 6};
static int32_t tip_row[7] = {-1, -1, 0, 1, -1, 2, 3};

static value make_ctx(int nstates, int ncat, long nsites, int ntaxa, int flags)
{
    return pb_stub_create(Val_int(0), Val_int(nstates), Val_int(ncat), Val_long(nsites), Val_int(ntaxa), Val_int(flags));
}
static void set_tree4(value ctx, double *bl)
{
    pb_stub_set_tree(ctx, Val_int(0), ba1(child_ptr, 8, CAML_BA_INT32), ba1(child_idx, 6, CAML_BA_INT32),
                     ba1(tip_row, 7, CAML_BA_INT32), ba1(bl, 7, CAML_BA_FLOAT64));
}
static void finalise(value ctx) { Custom_ops_val(ctx)->finalize(ctx); finalised++; }

int main(void)
{
    static double quarter[4] = {0.25, 0.25, 0.25, 0.25};
    double empty = 0.0;

    /* 1. Felsenstein(K80 kappa = 2).felsenstein on the fixture of lib/mCMC.ml:22,36-38:
     *    ((T0:0.1,T1:0.1):0.1,(T2:3.0,T3:0.1):0.1), columns A,A,A,T -> -5.704499092002674;
     *    call order of Phylo_b200.Felsenstein(..).evaluate */
    {
        double bl[7] = {0.0, 0.1, 0.1, 0.1, 0.1, 3.0, 0.1};
        double q[16];                       /* column-major; nucleotides A,C,G,T (lib/nucleotide.ml:3-6) */
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                q[j * 4 + i] = i == j ? -1.0 : ((i ^ j) == 2 ? 0.5 : 0.25);    /* A<->G, C<->T transitions: kappa/(kappa+2) */
        uint8_t tips[4] = {0, 0, 0, 3};
        value ctx = make_ctx(4, 1, 1, 4, 0);
        set_tree4(ctx, bl);
        pb_stub_set_rate_matrix(ctx, ba2(q, 4, 4, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT), ba1(quarter, 4, CAML_BA_FLOAT64));
        pb_stub_set_tips(ctx, ba2(tips, 4, 1, CAML_BA_UINT8, CAML_BA_C_LAYOUT));
        pb_stub_set_rescaling(ctx, caml_copy_double(0.0000001), Val_false);
        pb_stub_set_root_frequencies(ctx, ba1(quarter, 4, CAML_BA_FLOAT64));
        expect_close("Felsenstein(K80).felsenstein, mCMC fixture, via the stubs",
                     Double_val(pb_stub_loglik(ctx, ba1(&empty, 0, CAML_BA_FLOAT64))), -5.704499092002674, 1e-13);
        bl[5] = 2.5;                        /* tests/test_MCMC.ml:11-12 */
        pb_stub_set_branch_lengths(ctx, ba1(bl, 7, CAML_BA_FLOAT64));
        expect_close("  the same with the long branch at 2.5 (set_branch_lengths)",
                     Double_val(pb_stub_loglik(ctx, ba1(&empty, 0, CAML_BA_FLOAT64))), -5.708825294632623, 1e-13);
        finalise(ctx);
        finalise(ctx);                      /* a second finalisation must be harmless */
        finalised--;
    }

    /* 2. Phylo_ctmc.pruning, JC69 on ((A:0.8,C:1.2):0.4,(G:0.6,T:0.3):0.1) -> -6.157406549426741 at site 0;
     *    call order of Phylo_b200.pruning_sites: caller-built matrices, 2-bit packed tips, per-site output */
    {
        enum { S = 6 };
        double bl[7] = {0.0, 0.4, 0.8, 1.2, 0.1, 0.6, 0.3}, zero[7] = {0};
        static double p[7 * 16];
        for (int v = 1; v < 7; v++) {
            const double e = exp(-4.0 * bl[v] / 3.0);
            for (int k = 0; k < 16; k++) p[v * 16 + k] = (k % 5 == 0) ? 0.25 + 0.75 * e : 0.25 - 0.25 * e;
        }
        const uint8_t cols[S][4] = {{0, 1, 2, 3}, {0, 0, 0, 0}, {3, 3, 1, 1}, {2, 0, 2, 0}, {1, 1, 1, 2}, {3, 2, 1, 0}};
        uint8_t bytes[4][S], packed[4][2] = {{0}};
        for (int r = 0; r < 4; r++)
            for (int s = 0; s < S; s++) {
                bytes[r][s] = cols[s][r];
                packed[r][s / 4] |= (uint8_t)(cols[s][r] << (2 * (s & 3)));
            }
        double per_site[2][S], total[2];
        for (int pass = 0; pass < 2; pass++) {
            value ctx = make_ctx(4, 1, S, 4, 0);
            set_tree4(ctx, zero);
            pb_stub_set_pmatrices(ctx, ba1(p, 7 * 16, CAML_BA_FLOAT64));
            if (pass == 0) pb_stub_set_tips_packed2(ctx, ba2(packed, 4, 2, CAML_BA_UINT8, CAML_BA_C_LAYOUT));
            else pb_stub_set_tips(ctx, ba2(bytes, 4, S, CAML_BA_UINT8, CAML_BA_C_LAYOUT));
            pb_stub_set_mode(ctx, Val_int(PB_MODE_PHYLO_CTMC));
            pb_stub_set_root_frequencies(ctx, ba1(quarter, 4, CAML_BA_FLOAT64));
            total[pass] = Double_val(pb_stub_loglik(ctx, ba1(per_site[pass], S, CAML_BA_FLOAT64)));
            finalise(ctx);
        }
        expect_close("Phylo_ctmc.pruning JC69 4 leaves A,C,G,T (packed tips), via the stubs", per_site[0][0], -6.157406549426741, 1e-13);
        double sum = 0.0;
        for (int s = 0; s < S; s++) sum += per_site[0][s];
        expect_close("  total = sum of the per-site values", total[0], sum, 1e-14);
        expect_true("  2-bit packed tips and one byte per site agree bit for bit",
                    total[0] == total[1] && !memcmp(per_site[0], per_site[1], sizeof per_site[0]));
    }

    /* 3. conditional likelihoods and conditional simulation (keep_clv), JC69 pair closed form at the root */
    {
        double bl[7] = {0.0, 0.4, 0.8, 1.2, 0.1, 0.6, 0.3};
        double q[16];
        for (int k = 0; k < 16; k++) q[k] = (k % 5 == 0) ? -1.0 : 1.0 / 3.0;
        uint8_t tips[4][2] = {{0, 2}, {1, 2}, {2, 2}, {3, 1}};
        value ctx = make_ctx(4, 1, 2, 4, PB_FLAG_KEEP_CLV);
        set_tree4(ctx, bl);
        pb_stub_set_rate_matrix_general(ctx, ba2(q, 4, 4, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT));   /* Pade route */
        pb_stub_set_tips(ctx, ba2(tips, 4, 2, CAML_BA_UINT8, CAML_BA_C_LAYOUT));
        pb_stub_set_mode(ctx, Val_int(PB_MODE_PHYLO_CTMC));
        pb_stub_set_root_frequencies(ctx, ba1(quarter, 4, CAML_BA_FLOAT64));
        double ps[2];
        pb_stub_loglik(ctx, ba1(ps, 2, CAML_BA_FLOAT64));
        expect_close("pruning through the Pade route (set_rate_matrix_general)", ps[0], -6.157406549426741, 1e-12);
        double v[2][4], carry[2];
        pb_stub_get_clv(ctx, Val_int(1), Val_int(0), ba2(v, 2, 4, CAML_BA_FLOAT64, CAML_BA_C_LAYOUT), ba1(carry, 2, CAML_BA_FLOAT64));
        /* node 1 = (A:0.8, C:1.2): v[i] = P_0.8(i, A) P_1.2(i, C), times exp(carry) */
        const double e8 = exp(-4.0 * 0.8 / 3.0), e12 = exp(-4.0 * 1.2 / 3.0);
        expect_close("conditional_likelihoods: node (A:0.8,C:1.2), state A", v[0][0] * exp(carry[0]),
                     (0.25 + 0.75 * e8) * (0.25 - 0.25 * e12), 1e-13);
        expect_close("conditional_likelihoods: node (A:0.8,C:1.2), state G", v[0][2] * exp(carry[0]),
                     (0.25 - 0.25 * e8) * (0.25 - 0.25 * e12), 1e-13);
        uint8_t states[7][2];
        memset(states, 0xff, sizeof states);
        pb_stub_conditional_simulation(ctx, caml_copy_int64(12345), ba2(states, 7, 2, CAML_BA_UINT8, CAML_BA_C_LAYOUT));
        int leaves_kept = states[2][0] == 0 && states[3][0] == 1 && states[5][0] == 2 && states[6][0] == 3 &&
                          states[2][1] == 2 && states[3][1] == 2 && states[5][1] == 2 && states[6][1] == 1;
        int inner_ok = 1;
        for (int s = 0; s < 2; s++) inner_ok &= states[0][s] < 4 && states[1][s] < 4 && states[4][s] < 4;
        expect_true("conditional_simulation: observed leaves keep their states", leaves_kept);
        expect_true("conditional_simulation: inner nodes hold states", inner_ok);
        EXPECT_RAISES("get_clv on a node that does not exist -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_get_clv(ctx, Val_int(99), Val_int(0), ba2(v, 2, 4, CAML_BA_FLOAT64, CAML_BA_C_LAYOUT),
                                      ba1(carry, 2, CAML_BA_FLOAT64)));
        finalise(ctx);
    }

    /* 4. Gamma categories through set_categories + set_eigensystem: equal rates 1 reproduce the plain value */
    {
        double bl[7] = {0.0, 0.4, 0.8, 1.2, 0.1, 0.6, 0.3};
        double q[16], u[16], lambda[4], uinv[16];
        for (int k = 0; k < 16; k++) q[k] = (k % 5 == 0) ? -1.0 : 1.0 / 3.0;
        expect_true("pb_host_reversible_eigensystem", pb_host_reversible_eigensystem(4, q, quarter, u, lambda, uinv) == PB_OK);
        uint8_t tips[4] = {0, 1, 2, 3};
        double rates[2] = {1.0, 1.0}, weights[2] = {0.5, 0.5};
        value ctx = make_ctx(4, 2, 1, 4, 0);
        set_tree4(ctx, bl);
        pb_stub_set_eigensystem(ctx, ba2(u, 4, 4, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT), ba1(lambda, 4, CAML_BA_FLOAT64),
                                ba2(uinv, 4, 4, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT));
        pb_stub_set_categories(ctx, ba1(rates, 2, CAML_BA_FLOAT64), ba1(weights, 2, CAML_BA_FLOAT64));
        pb_stub_set_tips(ctx, ba2(tips, 4, 1, CAML_BA_UINT8, CAML_BA_C_LAYOUT));
        pb_stub_set_mode(ctx, Val_int(PB_MODE_PHYLO_CTMC));
        pb_stub_set_root_frequencies(ctx, ba1(quarter, 4, CAML_BA_FLOAT64));
        expect_close("two rate categories of rate 1 (set_eigensystem, set_categories)",
                     Double_val(pb_stub_loglik(ctx, ba1(&empty, 0, CAML_BA_FLOAT64))), -6.157406549426741, 1e-13);

        /* stand-alone P(t): Site_evolution_model (eigen) and Rate_matrix.transition_probability_matrix (Pade) */
        double ts[2] = {0.1, 0.7}, out_e[32], out_p[32];
        value args[6] = {Val_int(0), ba2(u, 4, 4, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT), ba1(lambda, 4, CAML_BA_FLOAT64),
                         ba2(uinv, 4, 4, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT), ba1(ts, 2, CAML_BA_FLOAT64),
                         ba1(out_e, 32, CAML_BA_FLOAT64)};
        pb_stub_transition_matrices_eigen_bytecode(args, 6);
        pb_stub_transition_matrices_expm(Val_int(0), ba2(q, 4, 4, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT),
                                         ba1(ts, 2, CAML_BA_FLOAT64), ba1(out_p, 32, CAML_BA_FLOAT64));
        expect_close("transition_matrices_eigen: JC69 P(0.7)(A,A)", out_e[16], 0.25 + 0.75 * exp(-4.0 * 0.7 / 3.0), 1e-14);
        expect_close("transition_matrices_expm:  JC69 P(0.7)(A,C)", out_p[16 + 4], 0.25 - 0.25 * exp(-4.0 * 0.7 / 3.0), 1e-13);
        expect_close("transition_matrices_expm:  JC69 P(0.1)(G,G)", out_p[10], 0.25 + 0.75 * exp(-4.0 * 0.1 / 3.0), 1e-13);
        finalise(ctx);
    }

    /* 5. error mapping (invalid_arg / failwith) */
    {
        EXPECT_RAISES("create with 1 state -> Invalid_argument (pb_create returns NULL, pb_last_status says why)", EXN_INVALID_ARGUMENT,
                      make_ctx(1, 1, 10, 4, 0));
        EXPECT_RAISES("create on a device that does not exist -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_create(Val_int(4096), Val_int(4), Val_int(1), Val_long(10), Val_int(4), Val_int(0)));
        value ctx = make_ctx(4, 1, 3, 4, 0);
        EXPECT_RAISES("set_mode 7 -> Invalid_argument", EXN_INVALID_ARGUMENT, pb_stub_set_mode(ctx, Val_int(7)));
        EXPECT_RAISES("loglik before set_tree -> Failure", EXN_FAILURE, pb_stub_loglik(ctx, ba1(&empty, 0, CAML_BA_FLOAT64)));
        double bl[7] = {0.0, 0.4, 0.8, 1.2, 0.1, 0.6, 0.3};
        double q[16];
        for (int k = 0; k < 16; k++) q[k] = (k % 5 == 0) ? -1.0 : 1.0 / 3.0;
        uint8_t bad[4][3] = {{0, 1, 2}, {0, 1, 2}, {0, 9, 2}, {0, 1, 2}};       /* state 9 of 4 */
        value a[6] = {ctx, Val_int(0), ba1(child_ptr, 8, CAML_BA_INT32), ba1(child_idx, 6, CAML_BA_INT32),
                      ba1(tip_row, 7, CAML_BA_INT32), ba1(bl, 7, CAML_BA_FLOAT64)};
        pb_stub_set_tree_bytecode(a, 6);
        pb_stub_set_rate_matrix(ctx, ba2(q, 4, 4, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT), ba1(quarter, 4, CAML_BA_FLOAT64));
        pb_stub_set_root_frequencies(ctx, ba1(quarter, 4, CAML_BA_FLOAT64));
        const int releases_before = lock_releases;
        exn_kind = EXN_NONE; exn_armed = 1;
        if (!setjmp(exn_jmp)) {
            pb_stub_set_tips(ctx, ba2(bad, 4, 3, CAML_BA_UINT8, CAML_BA_C_LAYOUT));
            pb_stub_loglik(ctx, ba1(&empty, 0, CAML_BA_FLOAT64));
        }
        exn_armed = 0;
        expect_true("a tip state >= nstates -> Invalid_argument (from set_tips or loglik)", exn_kind == EXN_INVALID_ARGUMENT);
        expect_true("  the runtime lock was released around the blocking calls", lock_releases > releases_before);
        finalise(ctx);
    }

    /* 6. the rest of the header: state-set masks (Missing_values / Ambiguous), alignment text, site weights, options,
     *    introspection, the one-call pipelined host evaluation on a pinned Bigarray, partial re-evaluation */
    {
        enum { S = 12000 };                                  /* enough columns for the chunked copy/compute pipeline */
        double bl[7] = {0.0, 0.4, 0.8, 1.2, 0.1, 0.6, 0.3};
        double q[16];
        for (int k = 0; k < 16; k++) q[k] = (k % 5 == 0) ? -1.0 : 1.0 / 3.0;
        static uint8_t bytes[4][S], packed[4][S / 4], text[4][S];
        static int64_t masks[4][S];
        static double per_site[3][S], w[S];
        memset(packed, 0, sizeof packed);
        for (int r = 0; r < 4; r++)
            for (int s = 0; s < S; s++) {
                const int st = s == 0 ? r : (r * 5 + s * 7 + (s >> 3)) & 3;     /* column 0 is A,C,G,T */
                bytes[r][s] = (uint8_t)st;
                packed[r][s / 4] |= (uint8_t)(st << (2 * (s & 3)));
                text[r][s] = (uint8_t)"ACGT"[st];
                masks[r][s] = (int64_t)1 << st;
                w[s] = 1.0 + (s % 3);
            }
        for (int r = 0; r < 4; r++) { masks[r][1] = 0; masks[r][2] = 15; }         /* all missing; all fully ambiguous */
        value ctx = make_ctx(4, 1, S, 4, 0);
        set_tree4(ctx, bl);
        pb_stub_set_rate_matrix(ctx, ba2(q, 4, 4, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT), ba1(quarter, 4, CAML_BA_FLOAT64));
        pb_stub_set_mode(ctx, Val_int(PB_MODE_PHYLO_CTMC));
        pb_stub_set_root_frequencies(ctx, ba1(quarter, 4, CAML_BA_FLOAT64));
        value vpacked = ba2(packed, 4, S / 4, CAML_BA_UINT8, CAML_BA_C_LAYOUT);
        pb_stub_pin(vpacked);
        const double t_host = Double_val(pb_stub_loglik_host_packed2(ctx, vpacked, ba1(per_site[0], S, CAML_BA_FLOAT64)));
        pb_stub_unpin(vpacked);
        expect_close("loglik_host_packed2 on a pinned Bigarray: site 0", per_site[0][0], -6.157406549426741, 1e-13);
        expect_true("  path_name / kernel_launches", !strcmp(String_val(pb_stub_path_name(ctx)), "fused4") &&
                                                     Long_val(pb_stub_kernel_launches(ctx)) > 0);
        const double t_bytes = Double_val(pb_stub_loglik_host(ctx, ba2(bytes, 4, S, CAML_BA_UINT8, CAML_BA_C_LAYOUT),
                                                              ba1(per_site[1], S, CAML_BA_FLOAT64)));
        expect_true("  loglik_host (one byte per site) agrees per site, bit for bit", !memcmp(per_site[0], per_site[1], sizeof per_site[0]));
        expect_close("  and in total", t_bytes, t_host, 1e-13);
        pb_stub_set_tips_text(ctx, ba2(text, 4, S, CAML_BA_UINT8, CAML_BA_C_LAYOUT), Val_int(PB_ALPHABET_DNA));
        pb_stub_set_option(ctx, caml_copy_string("fused_clade_leaves"), caml_copy_double(2.0));
        const double t_text = Double_val(pb_stub_loglik(ctx, ba1(per_site[2], S, CAML_BA_FLOAT64)));
        expect_true("set_tips_text + set_option: the same per-site values", fabs(t_text - t_bytes) <= 1e-13 * fabs(t_bytes) && !memcmp(per_site[1], per_site[2], sizeof per_site[1]));
        EXPECT_RAISES("set_option with an unknown name -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_option(ctx, caml_copy_string("no_such_option"), caml_copy_double(1.0)));
        pb_stub_set_site_weights(ctx, ba1(w, S, CAML_BA_FLOAT64));
        double weighted = 0.0;
        for (int s = 0; s < S; s++) weighted += w[s] * per_site[1][s];
        expect_close("set_site_weights: total = sum w[s] logL[s]", Double_val(pb_stub_loglik(ctx, ba1(&empty, 0, CAML_BA_FLOAT64))), weighted, 1e-12);
        pb_stub_set_site_weights(ctx, ba1(&empty, 0, CAML_BA_FLOAT64));
        pb_stub_set_tip_masks(ctx, ba2(masks, 4, S, CAML_BA_INT64, CAML_BA_C_LAYOUT));
        pb_stub_loglik(ctx, ba1(per_site[2], S, CAML_BA_FLOAT64));
        expect_close("set_tip_masks: singleton sets = observed states", per_site[2][0], -6.157406549426741, 1e-13);
        expect_true("  an all-missing column and a fully ambiguous one give log-likelihood 0", per_site[2][1] == 0.0 && fabs(per_site[2][2]) < 1e-13);
        expect_close("  every other column as with observed states", per_site[2][S - 1], per_site[1][S - 1], 1e-13);
        finalise(ctx);

        /* partial re-evaluation: one branch changes, only its path to the root is recomputed */
        value keep = make_ctx(4, 1, S, 4, PB_FLAG_KEEP_CLV);
        set_tree4(keep, bl);
        pb_stub_set_rate_matrix(keep, ba2(q, 4, 4, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT), ba1(quarter, 4, CAML_BA_FLOAT64));
        pb_stub_set_mode(keep, Val_int(PB_MODE_PHYLO_CTMC));
        pb_stub_set_root_frequencies(keep, ba1(quarter, 4, CAML_BA_FLOAT64));
        pb_stub_set_tips(keep, ba2(bytes, 4, S, CAML_BA_UINT8, CAML_BA_C_LAYOUT));
        pb_stub_loglik(keep, ba1(&empty, 0, CAML_BA_FLOAT64));
        pb_stub_update_branch_length(keep, Val_int(5), caml_copy_double(0.9));
        const double t_part = Double_val(pb_stub_loglik(keep, ba1(&empty, 0, CAML_BA_FLOAT64)));
        const int recomputed = Int_val(pb_stub_nodes_recomputed(keep));
        bl[5] = 0.9;
        pb_stub_set_branch_lengths(keep, ba1(bl, 7, CAML_BA_FLOAT64));
        const double t_full = Double_val(pb_stub_loglik(keep, ba1(&empty, 0, CAML_BA_FLOAT64)));
        expect_close("update_branch_length: partial re-evaluation = full evaluation", t_part, t_full, 1e-14);
        expect_true("  only the path to the root was recomputed (2 of 3 internal nodes)", recomputed == 2);
        static double pm[7 * 16];
        pb_stub_get_pmatrices(keep, ba1(pm, 7 * 16, CAML_BA_FLOAT64));
        expect_close("get_pmatrices: P(0.9)(A,A) of the changed branch", pm[5 * 16], 0.25 + 0.75 * exp(-4.0 * 0.9 / 3.0), 1e-13);
        finalise(keep);
    }

    /* 7. every Bigarray is measured against the context before its pointer crosses the ABI: a short one is an
     *    Invalid_argument, never an out-of-bounds access (the ABI carries no sizes) */
    {
        enum { S = 40 };
        double bl[7] = {0.0, 0.4, 0.8, 1.2, 0.1, 0.6, 0.3};
        static uint8_t tips[4][S];
        static int64_t masks[4][S];
        static double ps[S], big[7 * 16];
        value ctx = make_ctx(4, 2, S, 4, PB_FLAG_KEEP_CLV);
        value d = pb_stub_get_dims(ctx);
        expect_true("get_dims before set_tree: (4, 2, 40, 4, 0, keep_clv)",
                    Long_val(Field(d, 0)) == 4 && Long_val(Field(d, 1)) == 2 && Long_val(Field(d, 2)) == S &&
                    Long_val(Field(d, 3)) == 4 && Long_val(Field(d, 4)) == 0 && Long_val(Field(d, 5)) == PB_FLAG_KEEP_CLV);
        EXPECT_RAISES("set_pmatrices before set_tree -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_pmatrices(ctx, ba1(big, 7 * 16, CAML_BA_FLOAT64)));
        EXPECT_RAISES("set_tree with a short child_ptr -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_tree(ctx, Val_int(0), ba1(child_ptr, 7, CAML_BA_INT32), ba1(child_idx, 6, CAML_BA_INT32),
                                       ba1(tip_row, 7, CAML_BA_INT32), ba1(bl, 7, CAML_BA_FLOAT64)));
        EXPECT_RAISES("set_tree with a short child_idx -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_tree(ctx, Val_int(0), ba1(child_ptr, 8, CAML_BA_INT32), ba1(child_idx, 5, CAML_BA_INT32),
                                       ba1(tip_row, 7, CAML_BA_INT32), ba1(bl, 7, CAML_BA_FLOAT64)));
        set_tree4(ctx, bl);
        expect_true("get_dims after set_tree: 7 nodes", Long_val(Field(pb_stub_get_dims(ctx), 4)) == 7);
        EXPECT_RAISES("set_branch_lengths with 6 of 7 entries -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_branch_lengths(ctx, ba1(bl, 6, CAML_BA_FLOAT64)));
        EXPECT_RAISES("set_tips with 3 of 4 rows -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_tips(ctx, ba2(tips, 3, S, CAML_BA_UINT8, CAML_BA_C_LAYOUT)));
        EXPECT_RAISES("set_tips with too few columns -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_tips(ctx, ba2(tips, 4, S - 1, CAML_BA_UINT8, CAML_BA_C_LAYOUT)));
        EXPECT_RAISES("set_tips_packed2 with too few bytes per row -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_tips_packed2(ctx, ba2(tips, 4, S / 4 - 1, CAML_BA_UINT8, CAML_BA_C_LAYOUT)));
        EXPECT_RAISES("set_tip_masks with 3 of 4 rows -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_tip_masks(ctx, ba2(masks, 3, S, CAML_BA_INT64, CAML_BA_C_LAYOUT)));
        EXPECT_RAISES("set_root_frequencies with 3 of 4 entries -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_root_frequencies(ctx, ba1(quarter, 3, CAML_BA_FLOAT64)));
        EXPECT_RAISES("set_categories with 1 of 2 rates -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_categories(ctx, ba1(quarter, 1, CAML_BA_FLOAT64), ba1(quarter, 2, CAML_BA_FLOAT64)));
        EXPECT_RAISES("set_rate_matrix with a 3 x 3 matrix -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_rate_matrix(ctx, ba2(big, 3, 3, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT), ba1(quarter, 4, CAML_BA_FLOAT64)));
        EXPECT_RAISES("set_pmatrices with one category of two -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_set_pmatrices(ctx, ba1(big, 7 * 16, CAML_BA_FLOAT64)));
        EXPECT_RAISES("loglik with a per_site array of nsites - 1 -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_loglik(ctx, ba1(ps, S - 1, CAML_BA_FLOAT64)));
        EXPECT_RAISES("conditional_likelihood with a short carry -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_get_clv(ctx, Val_int(1), Val_int(0), ba2(big, S, 4, CAML_BA_FLOAT64, CAML_BA_C_LAYOUT), ba1(ps, S - 1, CAML_BA_FLOAT64)));
        EXPECT_RAISES("conditional_simulation with 6 of 7 node rows -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_conditional_simulation(ctx, caml_copy_int64(1), ba2(tips, 6, S, CAML_BA_UINT8, CAML_BA_C_LAYOUT)));
        EXPECT_RAISES("loglik_host with too few columns -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_loglik_host(ctx, ba2(tips, 4, S - 1, CAML_BA_UINT8, CAML_BA_C_LAYOUT), ba1(&empty, 0, CAML_BA_FLOAT64)));
        finalise(ctx);
    }

    /* 8. Felsenstein.Make(A)(Align)(E).felsenstein_single_shift ?threshold param ~site (lib/felsenstein.mli:19-35): E is a
     *    plain Site_evolution_model.S, so the binding calls E.transition_probability_matrix once per branch on the host
     *    and uploads the matrices (set_pmatrices), takes ONE column of the alignment, threshold as given, no shift after
     *    x pi.  The K80 fixture again; then felsenstein_single (no shift at all: threshold -infinity). */
    {
        double bl[7] = {0.0, 0.1, 0.1, 0.1, 0.1, 3.0, 0.1}, zero[7] = {0};
        static double p[7 * 16];
        /* (the matrices play E.transition_probability_matrix param t; here they come from the device's eigen route,
         * read back with get_pmatrices: the point is the call sequence of the binding) */
        double q[16];
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                q[j * 4 + i] = i == j ? -1.0 : ((i ^ j) == 2 ? 0.5 : 0.25);
        value src = make_ctx(4, 1, 1, 4, 0);
        set_tree4(src, bl);
        pb_stub_set_rate_matrix(src, ba2(q, 4, 4, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT), ba1(quarter, 4, CAML_BA_FLOAT64));
        pb_stub_get_pmatrices(src, ba1(p, 7 * 16, CAML_BA_FLOAT64));
        finalise(src);
        uint8_t column[4] = {0, 0, 0, 3};       /* the alignment's column `site` */
        for (int pass = 0; pass < 2; pass++) {
            value ctx = make_ctx(4, 1, 1, 4, 0);
            set_tree4(ctx, zero);
            pb_stub_set_pmatrices(ctx, ba1(p, 7 * 16, CAML_BA_FLOAT64));
            pb_stub_set_tips(ctx, ba2(column, 4, 1, CAML_BA_UINT8, CAML_BA_C_LAYOUT));
            pb_stub_set_rescaling(ctx, caml_copy_double(pass == 0 ? 0.0000001 : -INFINITY), Val_false);
            pb_stub_set_root_frequencies(ctx, ba1(quarter, 4, CAML_BA_FLOAT64));
            expect_close(pass == 0 ? "Felsenstein.Make(..).felsenstein_single_shift, K80 fixture (host matrices)"
                                   : "Felsenstein.Make(..).felsenstein_single (no shift), K80 fixture",
                         Double_val(pb_stub_loglik(ctx, ba1(&empty, 0, CAML_BA_FLOAT64))), -5.704499092002674, 1e-12);
            finalise(ctx);
        }
    }

    /* 9. multi-GPU through the binding (Phylo_b200.Sharded): one device here, so no communicator; the same device
     *    twice is NCCL's error and must come back as Failure, an empty device list as Invalid_argument */
    {
        enum { S = 9000 };
        double bl[7] = {0.0, 0.4, 0.8, 1.2, 0.1, 0.6, 0.3};
        double q[16];
        for (int k = 0; k < 16; k++) q[k] = (k % 5 == 0) ? -1.0 : 1.0 / 3.0;
        static uint8_t bytes[4][S];
        static double ps[S];
        for (int r = 0; r < 4; r++)
            for (int s = 0; s < S; s++) bytes[r][s] = (uint8_t)(s == 0 ? r : (r * 3 + s * 5 + (s >> 4)) & 3);
        int32_t dev0[2] = {0, 0};
        value sh = pb_stub_create_sharded(ba1(dev0, 1, CAML_BA_INT32), Val_int(4), Val_int(1), Val_long(S), Val_int(4), Val_int(0));
        expect_true("Sharded.create on one device", Int_val(pb_stub_sharded_ndev(sh)) == 1);
        value a[6] = {sh, Val_int(0), ba1(child_ptr, 8, CAML_BA_INT32), ba1(child_idx, 6, CAML_BA_INT32),
                      ba1(tip_row, 7, CAML_BA_INT32), ba1(bl, 7, CAML_BA_FLOAT64)};
        pb_stub_sharded_set_tree_bytecode(a, 6);
        pb_stub_sharded_set_rate_matrix(sh, ba2(q, 4, 4, CAML_BA_FLOAT64, CAML_BA_FORTRAN_LAYOUT), ba1(quarter, 4, CAML_BA_FLOAT64));
        pb_stub_sharded_set_root_frequencies(sh, ba1(quarter, 4, CAML_BA_FLOAT64));
        pb_stub_sharded_set_tips(sh, ba2(bytes, 4, S, CAML_BA_UINT8, CAML_BA_C_LAYOUT));
        const double t1 = Double_val(pb_stub_sharded_loglik(sh, ba1(ps, S, CAML_BA_FLOAT64)));
        expect_close("Sharded.loglik: site 0 (A,C,G,T under JC69)", ps[0], -6.157406549426741, 1e-13);
        double sum = 0.0;
        for (int s = 0; s < S; s++) sum += ps[s];
        expect_close("  total = sum of the per-site values", t1, sum, 1e-12);     /* (a sequential sum here, a tree sum there) */
        pb_stub_sharded_set_reduce(sh, Val_int(PB_REDUCE_ORDERED));
        const double t2 = Double_val(pb_stub_sharded_loglik_host(sh, ba2(bytes, 4, S, CAML_BA_UINT8, CAML_BA_C_LAYOUT), Val_false,
                                                                 ba1(&empty, 0, CAML_BA_FLOAT64)));
        expect_true("  Sharded.loglik_host, ordered reduction: the same total", t2 == t1);
        EXPECT_RAISES("Sharded.loglik with a short per_site -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_sharded_loglik(sh, ba1(ps, S - 1, CAML_BA_FLOAT64)));
        EXPECT_RAISES("Sharded.set_tips with too few columns -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_sharded_set_tips(sh, ba2(bytes, 4, S - 8, CAML_BA_UINT8, CAML_BA_C_LAYOUT)));
        Custom_ops_val(sh)->finalize(sh);
        Custom_ops_val(sh)->finalize(sh);       /* harmless twice */
        EXPECT_RAISES("Sharded.create with the same device twice -> Failure (PB_ERR_NCCL)", EXN_FAILURE,
                      pb_stub_create_sharded(ba1(dev0, 2, CAML_BA_INT32), Val_int(4), Val_int(1), Val_long(S), Val_int(4), Val_int(0)));
        expect_true("  and the message names NCCL", strstr(exn_msg, "NCCL") != NULL);
        EXPECT_RAISES("Sharded.create without devices -> Invalid_argument", EXN_INVALID_ARGUMENT,
                      pb_stub_create_sharded(ba1(dev0, 0, CAML_BA_INT32), Val_int(4), Val_int(1), Val_long(S), Val_int(4), Val_int(0)));
    }

    expect_true("runtime lock: every release paired with an acquire, none across a raise", lock_errors == 0 && !lock_released);
    expect_true("every context finalised", finalised == 12);
    printf(failures ? "%d FAILED\n" : "all ok\n", failures);
    return failures ? 1 : 0;
}
