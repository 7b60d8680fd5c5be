/*
 * CPU oracle: C restatement of the reference's Felsenstein pruning.
 *
 * TEST INFRASTRUCTURE ONLY.  Not part of the product; the CUDA path must never
 * call into this file.  Allowed users: tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status: pinned against the reference's golden vector -117.326089
 * (tests/expect/phylo_ctmc_substitution_mappings.expected:5) through a replay of
 * its GSL random stream (tests/test_oracle_reference_golden.py), and certified
 * to ~1e-13 by closed forms, the reference tests' RNG-free properties and a
 * 60-digit mpmath restatement (see oracle/__init__.py).
 *
 * Algorithm followed, operation for operation:
 *   Phylo_ctmc.pruning            lib/phylo_ctmc.ml:83-98
 *   SV.shift / SV.mul             lib/phylo_ctmc.ml:33-40, :76-78
 *   Missing_values / Ambiguous    lib/phylo_ctmc.ml:145-172, :280-302
 *   Felsenstein.felsenstein       lib/felsenstein.ml:22-48, :57-62, :73-78
 * One site at a time, post-order recursion, one dense n x n mat-vec per branch
 * (Matrix.apply = dgemv, lib/linear_algebra.ml:222), children folded left to
 * right.  The per-branch matrices P are inputs (what `transition_probabilities`
 * returns as [`Mat p]); building them once per branch instead of once per
 * branch per site is the CPU-favourable choice.
 *
 * Rate categories (an extension): C independent prunings per
 * site combined by log-sum-exp with weights.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 64

typedef struct {
    int n;
    const int32_t *child_ptr, *child_idx, *tip_row;
    const double *P;            /* [nnodes][n*n] COLUMN-major (element (i,j) at j*n+i, as the reference's
                                   Fortran-layout Bigarray) for the current category */
    const uint8_t *tips;        /* [ntaxa][ld] or NULL */
    const uint64_t *masks;      /* [ntaxa][ld] or NULL: state sets (bit i = state i allowed) */
    int64_t ld, site;
    double threshold;
} walk_t;

/* SV.shift, lib/phylo_ctmc.ml:33-40 (and Felsenstein.shift_normal, lib/felsenstein.ml:57-62):
 * keep when min v > threshold, else scale by 1/max and add log max to the carry. */
static void sv_shift(int n, double *v, double *carry, double threshold)
{
    double mn = v[0], mx = v[0];
    for (int i = 1; i < n; i++) {
        if (v[i] < mn) mn = v[i];
        if (v[i] > mx) mx = v[i];
    }
    if (mn > threshold) return;
    double inv = 1.0 / mx;                       /* Vector.scal_mul (1. /. mv) v */
    for (int i = 0; i < n; i++) v[i] = inv * v[i];
    *carry += log(mx);
}

/* Returns 1 and fills (v, carry) with the node's shifted vector, 0 for "None"
 * (no observation below this node). */
static int walk(const walk_t *w, int node, double *v, double *carry)
{
    const int n = w->n;
    const int c0 = w->child_ptr[node], c1 = w->child_ptr[node + 1];
    if (c0 == c1) {                              /* Leaf: indicator, carry 0 (:86-89) */
        const int64_t at = (int64_t)w->tip_row[node] * w->ld + w->site;
        if (w->masks) {
            const uint64_t m = w->masks[at];
            if (m == 0) return 0;                /* Missing / empty set */
            for (int i = 0; i < n; i++) v[i] = (double)((m >> i) & 1u);
        } else {
            const int s = w->tips[at];
            for (int i = 0; i < n; i++) v[i] = (i == s) ? 1.0 : 0.0;
        }
        *carry = 0.0;
        return 1;
    }
    int have = 0;
    double sub[MAXN], term[MAXN], sub_carry;
    for (int c = c0; c < c1; c++) {              /* List1.map over branches (:91-93) */
        const int child = w->child_idx[c];
        if (!walk(w, child, sub, &sub_carry)) continue;
        const double *Pm = w->P + (size_t)child * n * n;
        for (int i = 0; i < n; i++) term[i] = 0.0;
        for (int j = 0; j < n; j++) {            /* dgemv, axpy order over columns */
            const double xj = sub[j];
            for (int i = 0; i < n; i++) term[i] += Pm[j * n + i] * xj;
        }
        if (!have) {
            memcpy(v, term, sizeof(double) * n);
            *carry = sub_carry;
            have = 1;
        } else {                                  /* SV.mul = Vector.mul then shift (:76-78) */
            for (int i = 0; i < n; i++) v[i] = v[i] * term[i];
            *carry = *carry + sub_carry;
            sv_shift(n, v, carry, w->threshold);
        }
    }
    return have;
}

/*
 * root_mode 1: Phylo_ctmc.pruning root   — SV.mul with pi (shifted), log(sum v) + carry (:97-98)
 * root_mode 0: Felsenstein root          — sum(v * pi), log, + carry, no shift (lib/felsenstein.ml:47-48)
 * per_site_cat (nullable): [ncat][nsites] log-likelihood of every category.
 * Returns 0, or -1 on bad arguments.
 */
int oracle_pruning(int nstates, int ncat, int nnodes, int root,
                   const int32_t *child_ptr, const int32_t *child_idx, const int32_t *tip_row,
                   const double *P, const double *cat_weights,
                   const uint8_t *tips, const uint64_t *masks, int64_t ld, int64_t nsites,
                   const double *pi, double threshold, int root_mode, int nthreads,
                   double *per_site, double *per_site_cat)
{
    if (nstates < 1 || nstates > MAXN || ncat < 1 || ncat > 64 || nnodes < 1 || (!tips && !masks))
        return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    const int n = nstates;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < nsites; s++) {
        double lc[64] = {0.0};
        int none = 0;
        for (int k = 0; k < ncat; k++) {
            walk_t w = { n, child_ptr, child_idx, tip_row, P + (size_t)k * nnodes * n * n,
                         tips, masks, ld, s, threshold };
            double v[MAXN], carry;
            if (!walk(&w, root, v, &carry)) { none = 1; lc[k] = 0.0; }
            else {
                for (int i = 0; i < n; i++) v[i] = v[i] * pi[i];
                if (root_mode == 1) sv_shift(n, v, &carry, threshold);
                double sum = 0.0;
                for (int i = 0; i < n; i++) sum += v[i];
                lc[k] = log(sum) + carry;
            }
            if (per_site_cat) per_site_cat[(size_t)k * nsites + s] = lc[k];
        }
        double out;
        if (none) out = 0.0;                      /* all-missing site => 0. (:172, :302) */
        else if (ncat == 1) out = lc[0];
        else {
            double m = lc[0];
            for (int k = 1; k < ncat; k++) if (lc[k] > m) m = lc[k];
            if (isinf(m) && m < 0) out = m;
            else {
                double acc = 0.0;
                for (int k = 0; k < ncat; k++) acc += cat_weights[k] * exp(lc[k] - m);
                out = m + log(acc);
            }
        }
        per_site[s] = out;
    }
    return 0;
}

/* Felsenstein.multisite order (lib/felsenstein.ml:73-75): acc' = f(site) + acc, site 0 first. */
double oracle_sum_sequential(const double *x, int64_t n)
{
    double acc = 0.0;
    for (int64_t i = 0; i < n; i++) acc = x[i] + acc;
    return acc;
}

/* Neumaier compensated sum in long double: the "exact" total used as truth. */
double oracle_sum_exact(const double *x, int64_t n)
{
    long double s = 0.0L, c = 0.0L;
    for (int64_t i = 0; i < n; i++) {
        long double xi = x[i], t = s + xi;
        if (fabsl(s) >= fabsl(xi)) c += (s - t) + xi; else c += (xi - t) + s;
        s = t;
    }
    return (double)(s + c);
}

/* Conditional likelihoods of every node for one site and one category,
 * Phylo_ctmc.conditional_likelihoods (lib/phylo_ctmc.ml:100-120): clv [nnodes][n],
 * carry [nnodes]; leaves get their indicator with carry 0. Nodes with no
 * observation below (masks) get NaN carry. */
static int walk_store(const walk_t *w, int node, double *clv, double *carries)
{
    const int n = w->n;
    double *v = clv + (size_t)node * n;
    const int c0 = w->child_ptr[node], c1 = w->child_ptr[node + 1];
    if (c0 == c1) {
        double c;
        int ok = walk(w, node, v, &c);
        carries[node] = ok ? c : NAN;
        return ok;
    }
    int have = 0;
    double term[MAXN];
    carries[node] = NAN;
    for (int c = c0; c < c1; c++) {
        const int child = w->child_idx[c];
        if (!walk_store(w, child, clv, carries)) continue;
        const double *sub = clv + (size_t)child * n;
        const double *Pm = w->P + (size_t)child * n * n;
        for (int i = 0; i < n; i++) term[i] = 0.0;
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) term[i] += Pm[j * n + i] * sub[j];
        if (!have) {
            memcpy(v, term, sizeof(double) * n);
            carries[node] = carries[child];
            have = 1;
        } else {
            for (int i = 0; i < n; i++) v[i] = v[i] * term[i];
            carries[node] = carries[node] + carries[child];
            sv_shift(n, v, &carries[node], w->threshold);
        }
    }
    return have;
}

int oracle_conditional_likelihoods(int nstates, int nnodes, int root,
                                   const int32_t *child_ptr, const int32_t *child_idx,
                                   const int32_t *tip_row, const double *P,
                                   const uint8_t *tips, const uint64_t *masks, int64_t ld,
                                   int64_t site, double threshold, double *clv, double *carries)
{
    if (nstates < 1 || nstates > MAXN || nnodes < 1 || (!tips && !masks)) return -1;
    walk_t w = { nstates, child_ptr, child_idx, tip_row, P, tips, masks, ld, site, threshold };
    walk_store(&w, root, clv, carries);
    return 0;
}
