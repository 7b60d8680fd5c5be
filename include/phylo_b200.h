/*
 * phylo_b200 — C ABI of the B200-native phylogenetic-likelihood hot path.
 *
 * This is the drop-in boundary for ONE path of biocaml/phylogenetics (OCaml):
 * Felsenstein pruning over a rooted tree with the reference's underflow
 * rescaling, and the per-branch transition matrices P(t) = exp(Qt).
 * The reference has no FFI of its own; the entry points below
 * are what OCaml C stubs over Bigarray buffers bind (see INTEGRATION.md and
 * ocaml/phylo_b200_stubs.c).  Plain C: pointers and sizes only.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   Phylo_ctmc.pruning                     lib/phylo_ctmc.mli:55-61   (lib/phylo_ctmc.ml:83-98)
 *   Phylo_ctmc.conditional_likelihoods     lib/phylo_ctmc.mli:68-73   (lib/phylo_ctmc.ml:100-120)
 *   Phylo_ctmc.Missing_values/Ambiguous    lib/phylo_ctmc.ml:145-172, :280-302
 *   Felsenstein.Make(..).felsenstein       lib/felsenstein.mli:11-48  (lib/felsenstein.ml:22-78)
 *   Site_evolution_model.S.transition_probability_matrix
 *                                          lib/site_evolution_model.mli:7-14 (.ml:27-51)
 *   Rate_matrix.transition_probability_matrix  lib/rate_matrix.mli:73-76 (.ml:132-138)
 *   Mutsel.transition_probability_matrix   lib/mutsel.mli:59 (.ml:96-97)
 *   Matrix.expm                            lib/linear_algebra.ml:308-404
 *
 * Conventions
 *   - All matrices cross this ABI COLUMN-MAJOR, element (i,j) at [j*n + i], the
 *     layout of the reference's Lacaml/Bigarray `mat` (lib/linear_algebra.ml:151-152).
 *     P(i,j) = Prob(child state j | parent state i): rows sum to 1.
 *   - Every function returns PB_OK (0) or a negative pb_status; the message is
 *     available from pb_last_error().  Nothing aborts, exits or throws across
 *     the ABI.  An OCaml wrapper maps PB_ERR_INVALID_ARG to Invalid_argument and
 *     the rest to Failure (the reference's conventions).
 *   - One context = one caller thread.  Device work is enqueued on the context's
 *     CUDA stream; the blocking entry points synchronise that stream only.
 *     Several contexts may share a device (the library serialises its own use of the
 *     device's constant bank between them).
 *   - There is no CPU fallback: every compute entry point fails with
 *     PB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef PHYLO_B200_H
#define PHYLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_ctx pb_ctx;

typedef enum {
    PB_OK = 0,
    PB_ERR_INVALID_ARG = -1,   /* -> OCaml Invalid_argument */
    PB_ERR_CUDA = -2,          /* CUDA runtime / no device */
    PB_ERR_STATE = -3,         /* call order: something required was not set */
    PB_ERR_NOMEM = -4,
    PB_ERR_NCCL = -5,
    PB_ERR_NUMERIC = -6        /* e.g. rate matrix not reversible for the eigen path */
} pb_status;

/* pb_create flags */
#define PB_FLAG_KEEP_CLV      0x1u  /* level-scheduled path; every node's conditional-likelihood
                                       array stays resident (pb_get_clv, partial re-evaluation) */
#define PB_FLAG_LEVEL_KERNELS 0x2u  /* force the level-scheduled path even when the fused
                                       site-resident kernel applies (buffers are recycled) */

/* Root / rescaling conventions (pb_set_mode) */
#define PB_MODE_PHYLO_CTMC   0  /* SV.shift threshold 1e-6, root product with pi is shifted
                                   (lib/phylo_ctmc.ml:33,97) */
#define PB_MODE_FELSENSTEIN  1  /* shift_normal threshold 1e-7, no shift after x pi
                                   (lib/felsenstein.ml:47-48,64) */

/* ---- life cycle ------------------------------------------------------- */

/* nstates 2..64, ncat >= 1 (1 = the reference; >1 = rate-category mixture),
 * nsites >= 1 columns held by THIS context (a site shard when multi-GPU),
 * ntaxa = rows of the tip matrix.  Returns NULL on failure (pb_last_error(NULL)). */
pb_ctx *pb_create(int device, int nstates, int ncat, int64_t nsites, int ntaxa, unsigned flags);
void pb_destroy(pb_ctx *ctx);
/* ctx may be NULL: last error of a failed pb_create on this thread. */
const char *pb_last_error(const pb_ctx *ctx);
/* The pb_status that message belongs to (a binding maps a failed pb_create to the right exception with it). */
int pb_last_status(const pb_ctx *ctx);
/* dims[6] = {nstates, ncat, nsites, ntaxa, nnodes of the current tree (0: none), flags}: what a binding needs to
 * check the sizes of the buffers it is handed before it passes their pointers on (the ABI carries no sizes). */
int pb_get_dims(const pb_ctx *ctx, int64_t *dims);
/* Enqueue on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the
 * context's own non-blocking stream.  The handle 0 means the CUDA default stream. */
int pb_set_stream(pb_ctx *ctx, void *cuda_stream);
int pb_use_private_stream(pb_ctx *ctx);
int pb_set_mode(pb_ctx *ctx, int mode);
/* Explicit threshold / root-shift, overriding pb_set_mode. */
int pb_set_rescaling(pb_ctx *ctx, double threshold, int shift_at_root);

/* ---- inputs ----------------------------------------------------------- */

/* Rooted tree of arbitrary arity >= 1 (Tree.t, lib/tree.ml:3-13) flattened by the
 * caller: node v's children are child_idx[child_ptr[v] .. child_ptr[v+1]) IN THE
 * REFERENCE'S BRANCH ORDER (the left fold of lib/phylo_ctmc.ml:94-95 follows it);
 * leaves have no children and tip_row[v] = their row in the tip matrix (else -1).
 * branch_len[v] is the length of the branch above v (ignored for the root; may
 * be NULL when matrices are supplied with pb_set_pmatrices).  Branch "b" is
 * identified everywhere by the node id at its lower end. */
int pb_set_tree(pb_ctx *ctx, int nnodes, int root, const int32_t *child_ptr,
                const int32_t *child_idx, const int32_t *tip_row, const double *branch_len);
int pb_set_branch_lengths(pb_ctx *ctx, const double *branch_len /* [nnodes] */);
/* Change ONE branch length (the branch above `node`; Phylogenetic_tree.set_branch_lengths,
 * lib/phylogenetic_tree.ml:153-161, one coordinate at a time).  With PB_FLAG_KEEP_CLV the next
 * evaluation recomputes only the conditional likelihoods on the path from that branch to the
 * root: every other node's array is already resident.  pb_nodes_recomputed reports how many
 * internal nodes the last evaluation touched. */
int pb_update_branch_length(pb_ctx *ctx, int node, double t);
int pb_nodes_recomputed(const pb_ctx *ctx);

/* Observed states, one byte per (taxon, site): states[row*ld + site] in 0..nstates-1
 * (`leaf_state`, lib/phylo_ctmc.mli:58; evaluated eagerly by the caller).  Host memory. */
int pb_set_tips(pb_ctx *ctx, const uint8_t *states, int64_t ld);
/* Same, from a device pointer (copied device-to-device on the context's stream). */
int pb_set_tips_device(pb_ctx *ctx, const uint8_t *d_states, int64_t ld);
/* Nucleotide alignments 2-bit packed on the host (4 states only): packed[row*ld + site/4], site 4b+j in
 * bits 2j..2j+1 of byte b.  The binding evaluates `leaf_state` once per (leaf, site) either way; filling
 * this layout instead of one byte per site cuts the host-to-device traffic four-fold. */
int pb_set_tips_packed2(pb_ctx *ctx, const uint8_t *packed, int64_t ld);
/* State SETS per (taxon, site): bit i set = state i compatible (Ambiguous.leaf_indicator,
 * lib/phylo_ctmc.ml:234-242); 0 = missing observation (Missing_values: leaf_state = None). */
int pb_set_tip_masks(pb_ctx *ctx, const uint64_t *masks, int64_t ld);

/* f4 — alignment text straight to tip states, encoded on the device: text[row*ld + site*w .. +w) with
 * w = 1 (nucleotides: A,C,G,T either case, lib/nucleotide.ml:8-13; amino acids:
 * "ACDEFGHIKLMNPQRSTVWY", lib/amino_acid.ml:4) or w = 3 (sense codons over TCAG, universal code,
 * lib/codon.ml:51-58,73).  A character outside the alphabet or a stop codon -> PB_ERR_INVALID_ARG
 * (the reference raises Invalid_argument in of_char_exn). */
#define PB_ALPHABET_DNA        0
#define PB_ALPHABET_AMINO_ACID 1
#define PB_ALPHABET_CODON      2
int pb_set_tips_text(pb_ctx *ctx, const char *text, int64_t ld, int alphabet);
/* f4 — site-pattern compression: the context holds the DISTINCT columns and w[s] is how many
 * original sites show column s; the total becomes sum_s w[s] * logL[s] (per-site values stay per
 * pattern).  NULL removes the weights. */
int pb_set_site_weights(pb_ctx *ctx, const double *w /* [nsites] */);

int pb_set_root_frequencies(pb_ctx *ctx, const double *pi /* [nstates] */);
/* Rate categories: branch lengths are multiplied by rates[k]; site likelihood =
 * sum_k weights[k] * L_k.  (Extension: the reference has no rate heterogeneity.) */
int pb_set_categories(pb_ctx *ctx, const double *rates, const double *weights /* [ncat] */);

/* P(t) = U diag(exp(t*lambda)) Uinv  — the decomposition Site_evolution_model.Make_diag
 * stages once (lib/site_evolution_model.ml:43-47), e.g. JC69/K80 reductions (:57-73,:97-117).
 * P matrices are rebuilt ON DEVICE for every branch x category in one launch. */
int pb_set_eigensystem(pb_ctx *ctx, const double *U, const double *lambda, const double *Uinv);
/* Reversible rate matrix Q with stationary distribution pi: the library
 * symmetrises (S = D^1/2 Q D^-1/2, D = diag pi), diagonalises S on the host (one n x n
 * Jacobi, cf. Matrix.diagonalize lib/linear_algebra.ml:238-241) and installs the
 * eigensystem.  PB_ERR_NUMERIC if pi_i q_ij != pi_j q_ji. */
int pb_set_rate_matrix(pb_ctx *ctx, const double *Q, const double *pi);
/* General (possibly non-reversible) rate matrix: P(t) = expm(t*r_k*Q) for every branch x
 * category by a batched Pade-13 scaling-and-squaring kernel (Matrix.expm,
 * lib/linear_algebra.ml:308-404; callers lib/site_evolution_model.ml:31-32, lib/mutsel.ml:96-97). */
int pb_set_rate_matrix_general(pb_ctx *ctx, const double *Q);
/* Caller-built matrices ([`Mat p] of lib/phylo_ctmc.ml:4-8): P[(k*nnodes + v)*n*n ...],
 * column-major each, for category k and the branch above node v (root's is ignored). */
int pb_set_pmatrices(pb_ctx *ctx, const double *P);

/* ---- evaluation ------------------------------------------------------- */

/* Sum over sites of the log-likelihood (Felsenstein.felsenstein / sum of
 * Phylo_ctmc.pruning) and, when per_site != NULL, each site's value [nsites].
 * Blocking; results are host memory. */
int pb_loglik(pb_ctx *ctx, double *total, double *per_site);
/* One-call evaluation from HOST tip states: copies the alignment to the device in column chunks
 * on a second stream while the previous chunk is being pruned (copy/compute overlap), then
 * behaves like pb_loglik.  `states` should be pinned memory for the copies to be asynchronous.
 * The states stay resident afterwards, as after pb_set_tips. */
int pb_loglik_host(pb_ctx *ctx, const uint8_t *states, int64_t ld, double *total, double *per_site);
/* The same from a 2-bit packed host alignment (see pb_set_tips_packed2).  The alignment stays resident 2-bit
 * packed; the byte-per-site matrix other entry points read is expanded on the device when one of them needs it. */
int pb_loglik_host_packed2(pb_ctx *ctx, const uint8_t *packed, int64_t ld, double *total, double *per_site);
/* Non-blocking form of the two calls above (packed2 != 0: the 2-bit layout): copies, pruning and reduction are
 * enqueued and the call returns; the caller may queue more work behind it on the context's stream (e.g. the
 * all-reduce of the device-resident total, pb_result_device_ptrs) before it waits in pb_loglik_fetch.  `states`
 * must stay valid and unchanged until pb_loglik_fetch (or a synchronisation of the stream) returns. */
int pb_loglik_host_enqueue(pb_ctx *ctx, const uint8_t *states, int64_t ld, int packed2);
/* Page-lock / release a caller-owned host buffer (e.g. the payload of an OCaml Bigarray, which lives outside the
 * OCaml heap and does not move) so that the chunked copies above are asynchronous.  No context: the message of a
 * failure is pb_last_error(NULL). */
int pb_pin_host_memory(void *p, size_t bytes);
int pb_unpin_host_memory(void *p);
/* Non-blocking half: enqueue P(t) + pruning + root reduction on the stream. */
int pb_loglik_enqueue(pb_ctx *ctx);
/* Blocking half: wait for the stream and fetch.  per_site may be NULL. */
int pb_loglik_fetch(pb_ctx *ctx, double *total, double *per_site);
/* Device pointers of the last evaluation's results (valid until the next call):
 * total (1 double) and per-site log-likelihoods [nsites]. */
int pb_result_device_ptrs(pb_ctx *ctx, const double **d_total, const double **d_per_site);

/* Conditional-likelihood array of one node (needs PB_FLAG_KEEP_CLV and a prior
 * evaluation): v[site*nstates + i] and carry[site] for category cat, i.e. the
 * shifted_vector SV (v, carry) of lib/phylo_ctmc.ml:26 per site.  For a leaf the
 * indicator vector with carry 0. */
int pb_get_clv(pb_ctx *ctx, int node, int cat, double *v, double *carry);
/* Phylo_ctmc.conditional_simulation (lib/phylo_ctmc.mli:80-90, .ml:122-143; Missing_values
 * :199-229; Ambiguous :331-360), batched over sites on the device: one joint draw of every
 * node's state given the leaves, from the conditional likelihoods of the last evaluation (needs
 * PB_FLAG_KEEP_CLV, one rate category).  states[node*ld + site]; an observed leaf keeps its state,
 * a missing / ambiguous leaf (pb_set_tip_masks) is drawn from the prior (restricted to its set).
 * The variates come from a counter-based generator keyed by (seed, site, node): the same seed
 * gives the same draw; the stream is not GSL's. */
int pb_conditional_simulation(pb_ctx *ctx, uint64_t seed, uint8_t *states /* host [nnodes][ld] */, int64_t ld);
/* Copy out the matrices in use: P[(k*nnodes + v)*n*n ...] column-major. */
int pb_get_pmatrices(pb_ctx *ctx, double *P);

/* Tunables (all optional; defaults are the measured best).
 * 4 states, site-resident kernel: "fused_block" (128|256|512 threads), "fused_spt" (1|2|4 sites per
 *   thread), "fused_exact_log" (1: add log(max v) at every rescaling like the reference's carry;
 *   0 (default): accumulate the product of the maxima and take one log), "fused_const" (inner
 *   matrices through the constant bank), "fused_const_cats" (categories per launch), "fused_pow2" (1 (default):
 *   exact power-of-two scale factors; 0: the reference's factor 1/max v), "fused_cherry" (tabulate the nodes
 *   with two leaf children), "fused_clade_leaves" (2..8, default 7: the largest clade whose 4^k
 *   possible values are tabulated the same way; fewer when the alignment has fewer than 8 * 4^k sites unless
 *   "fused_clade_force"), "fused_triple" (0: cherries only), "fused_reduced" (1 (default): the program over the reduced tree, kernel fused4t;
 *   0: look-ups inside the full program, kernel fused4c), "fused_fold" (1 (default): the reduced program's table entries
 *   carry their scale exponent inside the vector, one gather per look-up), "fused_const_keep" (1 (default): a constant
 *   bank that already holds these matrices is not reloaded), "fused_split" (measurement hook: the resident evaluation
 *   in this many launches).
 * 4 states, level-scheduled: "level4_async" (cp.async-pipelined kernel), "level4_waves",
 *   "level4_min_blocks".
 * 20 / 61 states: "fused_dmma" (0: level-scheduled kernels), "fused_dmma_version" (3 = m8n8k4 tiles,
 *   2 = m16n8k8 tiles, 1 = CTA-synchronous), "fused_dmma_cherry" (1 (default): nodes with two leaf children are
 *   tables of N^2 entries carrying their branch matrix, version 3 with power-of-two rescaling), "fused_dmma_triple" (1:
 *   at 20 states a cherry joined by a leaf is a table of N^3 entries; default 0), "fused_dmma_slots" (cap on the matrix ring),
 *   "fused_dmma_spill" (test hook: stack in L2), "dmma_version", "dmma_per_level",
 *   "dmma_group_tiles", "dense_scalar" (DFMA level kernel).
 * "graph": see pb_graph_replays.  "host_chunks" (default 6): column chunks of pb_loglik_host*, whole rounds of the
 *   persistent blocks laid from the end of the alignment unless "host_chunk_align" is 0; "host_tail_sites": sites of
 *   the short last chunk the schedule ramps down to (0 (default): one round; -1: none); "host_side_streams" (1
 *   (default): chunks alternate between "host_lanes" streams (2..4, default 2: three and four measured equal); 0:
 *   everything on the context's stream); "host_lead_cats": rate
 *   categories pruned chunk by chunk behind a 2-bit packed copy (default: all; the others run once over all the
 *   sites); "host_trace": see pb_debug_host_trace; "fused_const_group" (1 (default): the constant bank is loaded for
 *   as many categories as it holds).  Unknown names are PB_ERR_INVALID_ARG. */
int pb_set_option(pb_ctx *ctx, const char *name, double value);

/* Device timing of the phases of an evaluation, measured with CUDA events on the
 * context's stream: ms[0] = P(t) construction, ms[1] = pruning kernels, ms[2] = root
 * reduction of the LAST evaluation (pb_get_timing waits for it). */
int pb_set_timing(pb_ctx *ctx, int on);
int pb_get_timing(pb_ctx *ctx, double *ms /* [3] */);
/* The same for the evaluation enqueued `back` evaluations before the last one (0 = last); the
 * events of the 64 most recent timed evaluations are kept, so a caller can queue many
 * evaluations without synchronising and read their phase times afterwards. */
int pb_get_timing_at(pb_ctx *ctx, int back, double *ms /* [3] */);
/* pb_set_timing(ctx, 2): the time of the DOMINANT kernel alone in the last evaluation — the site-resident pruning kernel of the
 * 4-state (fused4t) and 20 / 61-state (fused_dmma3) default paths, summed over its launches (*launches, may be null),
 * from CUDA events around each launch; 0 launches on other paths.  pb_get_timing's pruning phase also holds the
 * per-evaluation table builders and matrix staging. */
int pb_get_kernel_timing(pb_ctx *ctx, double *ms, int *launches);

/* How many kernels of this library the context has launched so far. */
int64_t pb_kernel_launches(const pb_ctx *ctx);
/* Option "graph" (pb_set_option(ctx, "graph", 1)): an optimisation loop re-evaluates one configuration with new
 * parameter VALUES (branch lengths, eigensystem, categories, root frequencies); the library then captures one
 * evaluation of the 4-state site-resident program — P(t), table build, constant-bank loads, the category launches,
 * the root reduction, the read-back — as a CUDA graph after the first plain evaluation and replays it until a call
 * changes what would be launched (tree, tips, options, rescaling, stream ...).  How many evaluations were replays: */
int64_t pb_graph_replays(const pb_ctx *ctx);
/* Which pruning path the next pb_loglik takes: "fused4", "level4", "level_dense", ... */
const char *pb_path_name(const pb_ctx *ctx);

/* ---- multi-GPU: one host thread, the GPUs of one box ------------------- */

/* Sites are independent given the tree and the model (the reference sums per-site values,
 * lib/felsenstein.ml:73-75): the alignment's site columns are cut into ndev contiguous shards
 * (bounds multiples of 4 sites), shard d lives in its own pb_ctx on devices[d], tree / branch lengths / model
 * / categories are replicated (P(t) is rebuilt on every device), and the ONLY exchange is the all-reduce of
 * one double per evaluation: ncclAllReduce(ncclDouble, count 1, sum) in place on every device's total, one
 * communicator per device from ncclCommInitAll, all driven by the calling thread (no MPI, no second process).
 * NCCL is bound at run time (libnccl.so.2); every NCCL failure — library not found, communicator
 * initialisation, the collective, a device listed twice — is reported as PB_ERR_NCCL.  ndev = 1 needs no NCCL.
 * Returns NULL on failure (message: pb_sharded_last_error(NULL)). */
typedef struct pb_sharded pb_sharded;
pb_sharded *pb_create_sharded(int ndev, const int *devices, int nstates, int ncat, int64_t nsites_total,
                              int ntaxa, unsigned flags);
void pb_sharded_destroy(pb_sharded *s);
const char *pb_sharded_last_error(const pb_sharded *s);
int pb_sharded_ndev(const pb_sharded *s);
/* Shard d's own context (per-device options, timing, pb_get_clv ...) and its columns [lo, hi). */
pb_ctx *pb_sharded_ctx(pb_sharded *s, int d);
int pb_sharded_shard_bounds(const pb_sharded *s, int d, int64_t *lo, int64_t *hi);
/* How the per-device totals are combined: PB_REDUCE_NCCL (default) or PB_REDUCE_ORDERED — the local totals
 * are read back and added in shard order on the host: bit-reproducible whatever the collective's algorithm. */
#define PB_REDUCE_NCCL    0
#define PB_REDUCE_ORDERED 1
int pb_sharded_set_reduce(pb_sharded *s, int how);
/* Replicated inputs: the same call on every shard's context (see the pb_set_* of the same name). */
int pb_sharded_set_mode(pb_sharded *s, int mode);
int pb_sharded_set_option(pb_sharded *s, const char *name, double value);
int pb_sharded_set_tree(pb_sharded *s, int nnodes, int root, const int32_t *child_ptr, const int32_t *child_idx,
                        const int32_t *tip_row, const double *branch_len);
int pb_sharded_set_branch_lengths(pb_sharded *s, const double *branch_len);
int pb_sharded_update_branch_length(pb_sharded *s, int node, double t);
int pb_sharded_set_root_frequencies(pb_sharded *s, const double *pi);
int pb_sharded_set_categories(pb_sharded *s, const double *rates, const double *weights);
int pb_sharded_set_eigensystem(pb_sharded *s, const double *U, const double *lambda, const double *Uinv);
int pb_sharded_set_rate_matrix(pb_sharded *s, const double *Q, const double *pi);
int pb_sharded_set_rate_matrix_general(pb_sharded *s, const double *Q);
int pb_sharded_set_pmatrices(pb_sharded *s, const double *P);
/* Sharded inputs: the caller passes the FULL alignment ([ntaxa][ld], ld >= nsites_total); every shard copies its
 * own columns (one strided copy per device). */
int pb_sharded_set_tips(pb_sharded *s, const uint8_t *states, int64_t ld);
int pb_sharded_set_tips_packed2(pb_sharded *s, const uint8_t *packed, int64_t ld);
int pb_sharded_set_tip_masks(pb_sharded *s, const uint64_t *masks, int64_t ld);
/* One evaluation: P(t) + pruning + root reduction enqueued on every device, the all-reduce behind them, then the
 * total (identical on every device) and, when per_site != NULL, every site's value [nsites_total] in the order of
 * the alignment.  Blocking. */
int pb_sharded_loglik(pb_sharded *s, double *total, double *per_site);
/* The same straight from a HOST alignment (pb_loglik_host / pb_loglik_host_packed2 per shard, copies overlapped
 * with the pruning on every device; packed2 != 0: the 2-bit layout). */
int pb_sharded_loglik_host(pb_sharded *s, const uint8_t *states, int64_t ld, int packed2, double *total,
                           double *per_site);

/* ---- stand-alone P(t) producers (no context) -------------------------- */

/* out[m] = U diag(exp(t[m]*lambda)) Uinv for m < count, one launch on `device`.
 * Site_evolution_model.Make_diag.transition_probability_matrix (.ml:43-47). */
int pb_transition_matrices_eigen(int device, int n, int count, const double *U,
                                 const double *lambda, const double *Uinv,
                                 const double *t, double *out);
/* out[m] = expm(t[m] * Q), batched Pade on `device`.
 * Rate_matrix.transition_probability_matrix ~tau ~rates (lib/rate_matrix.ml:132-138). */
int pb_transition_matrices_expm(int device, int n, int count, const double *Q,
                                const double *t, double *out);

/* ---- host-side helpers (no GPU; exposed for callers without LAPACK) ---- */

/* Eigen-decomposition of a reversible Q as used by pb_set_rate_matrix. */
int pb_host_reversible_eigensystem(int n, const double *Q, const double *pi,
                                   double *U, double *lambda, double *Uinv);

/* Host-only introspection of the 4-state instruction programs the library compiles for a tree — the plain program
 * and the one in which clades of up to `max_leaves` leaves are table look-ups — for tests of the host compiler
 * (tests/test_fused_program_cpu.py interprets them on the CPU against the oracle).  No device is touched.
 * counts[8] = {instructions, instructions of the tabulated program, tips, packed words per site, stack depth,
 * inner branches, leaf branches, tabulated clades}; an instruction is 4 ints {code, offA, offB, sh}
 * (pb_kernels_fused.cuh); tip_order[k] = tip-matrix row consumed k-th; inner_nodes / leaf_nodes map the compact
 * matrix indices (off / 16) to node ids; a clade is 6 ints {op0, op1, sh0, leaves, first table entry, pre}.
 * max_leaves | 0x100 plans the variant in which a table holds P . v for its consumer (pre = that inner matrix's
 * offset, else -1; the consumer carries the flag 0x200 / 0x400 for operand A / B): a kernel build option that is
 * off by default (FUSED_PREAPPLY, pb_kernels_fused.cuh). */
int pb_debug_fused_program(int nnodes, int root, const int32_t *child_ptr, const int32_t *child_idx,
                           const int32_t *tip_row, int ntaxa, int max_leaves, int32_t *counts,
                           int32_t *ops, int32_t *ops_tab, int cap_ops, int32_t *tip_order,
                           int32_t *inner_nodes, int32_t *leaf_nodes, int32_t *clades, int cap_clades);

/* The same for the program over the REDUCED tree, the default 4-state program (kernel fused4t): leaves of the reduced
 * tree are the tabulated clades (every maximal clade of at most `max_leaves` leaves) and the loose tips; a table value
 * is an operand of its parent's instruction.  counts[8] = {instructions, instructions of all clade ranges, packed-tip
 * slots, packed words per site, stack depth, inner matrices applied, loose-tip matrices, clades}; slot_rows[s] = the
 * tip-matrix row whose 2-bit state sits in slot s (32 slots per 64-bit word; -1 = padding), XORed with the state of
 * row slot_xor_rows[s] when that is >= 0 (the tips of a clade but the first are stored as differences from the first,
 * so that a table's most frequent entries are neighbours); ops address matrices by
 * compact index * 16 into mat_nodes / leaf_nodes and tables by first entry (sh = shA | shB << 8 | bitsA << 16 |
 * bitsB << 24); clade_ops address matrices by node id * 16 and tips by 2 * (ordinal within the clade); a clade is
 * 6 ints {op0, op1 (range in clade_ops), 1 if its index is difference-coded, leaves, first table entry, pre = node id * 16 of the branch matrix the
 * table carries, or -1 at the root}. */
int pb_debug_reduced_program(int nnodes, int root, const int32_t *child_ptr, const int32_t *child_idx,
                             const int32_t *tip_row, int ntaxa, int max_leaves, int32_t *counts,
                             int32_t *ops, int cap_ops, int32_t *clade_ops, int cap_clade_ops,
                             int32_t *slot_rows, int32_t *slot_xor_rows, int cap_slots, int32_t *mat_nodes, int32_t *leaf_nodes,
                             int32_t *clades, int cap_clades);

/* The dense-alphabet (20 / 61 states) program over the tree whose cherries — nodes with two leaf children — are
 * tables (kernel fused_dmma3); nstates | 0x100 also plans three-leaf clades (a cherry joined by a leaf, option
 * "fused_dmma_triple", 20 states).  counts[8] = {instructions, matrices applied (in the order the kernel's matrix ring
 * delivers them), tables, stack depth, inner branches, leaf branches, stack depth and matrices of the untabulated
 * program}; an instruction is 4 ints {code, A, B, rowA | rowB << 16}: A / B are compact indices into inner_nodes
 * (matrices), leaf_nodes (tip columns) or table numbers, by operand kind; tabs[t] = {first entry, tip row A, tip row
 * B, tip row C or -1} (entry = (state A * nstates + state B) [* nstates + state C]); tab_nodes[t] (8 ints) = {leaf A
 * node, leaf B node, cherry node, leaf C node or -1, top node (the cherry, or its parent when leaf C joins it), 1 if
 * top is the root, 0, 0}: the table holds P[top] . v (v alone at the root), v = column of P[leaf A] * column of
 * P[leaf B], or (P[cherry] . that) * column of P[leaf C] for a three-leaf clade (20 states), rescaled. */
int pb_debug_dense_reduced_program(int nnodes, int root, const int32_t *child_ptr, const int32_t *child_idx,
                                   const int32_t *tip_row, int ntaxa, int nstates, int32_t *counts, int32_t *ops,
                                   int cap_ops, int32_t *mlist, int cap_mlist, int32_t *tabs, int32_t *tab_nodes,
                                   int cap_tabs, int32_t *inner_nodes, int32_t *leaf_nodes);

/* Timeline of the last pipelined pb_loglik_host* call (option "host_trace" = 1): out[0] = number of column chunks K, then
 * per chunk {first site, sites, ms at which its host-to-device copy ended, ms at which its pruning and root reduction
 * ended}, then out[1 + 4K] = ms at which the total was on its way to the host; milliseconds since the call's first
 * enqueue, from CUDA events.  cap >= 2 + 4K doubles. */
int pb_debug_host_trace(pb_ctx *ctx, double *out, int cap);

/* Host-only: the column-chunk schedule of a pipelined host evaluation over `spad` site columns (a multiple of 1024) —
 * options "host_chunks", "host_tail_sites", and the sites one round of the persistent pruning blocks covers (0: the
 * unaligned schedule of option "host_chunk_align" = 0); host_tail_sites = -2: the schedule of the dense alphabets'
 * compute-bound pipeline, chunks of 1, 2, 4, ... rounds from the start.  bounds[0] = 0 < ... < bounds[*nbounds - 1] = spad, every bound
 * a multiple of 1024, at most 65 of them.  tests/test_abi_cpu.py checks the invariants. */
int pb_debug_host_schedule(int64_t spad, int host_chunks, int64_t host_tail_sites, int64_t round_sites, int64_t *bounds,
                           int cap, int *nbounds);

/* The device's FP64 peak, measured: kind 0 = DFMA (register-resident dependent chains, 2 flops per FMA), kind 1 = the
 * FP64 tensor-core shape mma.sync.m8n8k4 (DMMA.8x8x4, 512 flops per warp instruction); ~10 ms.  bench.py reports them
 * beside the constants its roofline fractions use. */
int pb_measure_fp64_peak(int device, int kind, double *tflops);

const char *pb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PHYLO_B200_H */
