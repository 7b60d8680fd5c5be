(** GPU (B200) evaluation of the phylogenetic likelihood behind the reference's signatures.

    Not compiled in this repository (no OCaml toolchain in the build image); the C stubs underneath are
    tested against a mock of the runtime: see INTEGRATION.md.
    [Phylo_ctmc.pruning] and [Felsenstein.Make(..).felsenstein] keep their types; the batched
    entry points are new and are what a caller with many sites should use. *)

open Phylogenetics

type ctx
(** One (tree, alignment, model) evaluator on one GPU; finalised by the GC. *)

type int32_vec = (int32, Bigarray.int32_elt, Bigarray.c_layout) Bigarray.Array1.t
type float_vec = (float, Bigarray.float64_elt, Bigarray.c_layout) Bigarray.Array1.t
type tip_matrix = (int, Bigarray.int8_unsigned_elt, Bigarray.c_layout) Bigarray.Array2.t
type float_mat = (float, Bigarray.float64_elt, Bigarray.c_layout) Bigarray.Array2.t
type mask_matrix = (int64, Bigarray.int64_elt, Bigarray.c_layout) Bigarray.Array2.t

type flat_tree = {
  root : int ;
  child_ptr : int32_vec ;   (** node v's children: child_idx.{child_ptr.{v}} .. child_idx.{child_ptr.{v+1} - 1} *)
  child_idx : int32_vec ;   (** in the order of [Tree.Node.branches] (the left fold of pruning follows it) *)
  tip_row : int32_vec ;     (** row of the tip matrix for leaves, -1 for internal nodes *)
}

val flatten : ('n, 'l, 'b) Tree.t -> flat_tree * 'l array * 'b option array
(** Pre-order numbering; also returns the leaf payloads (by tip row) and the payload of the branch
    above every node ([None] for the root). *)

val create : ?device:int -> ?ncat:int -> ?keep_clv:bool -> nstates:int -> nsites:int -> ntaxa:int -> unit -> ctx

(** {2 Drop-in replacements} *)

val pruning :
  ('n, 'l, 'b) Tree.t ->
  nstates:int ->
  transition_probabilities:('b -> Phylo_ctmc.matrix_decomposition) ->
  leaf_state:('l -> int) ->
  root_frequencies:Linear_algebra.vec ->
  float
(** Same type and value as [Phylo_ctmc.pruning] (lib/phylo_ctmc.mli:55-61).  One site per call:
    correct but dominated by transfer latency; prefer {!pruning_sites}. *)

val pruning_sites :
  ('n, 'l, 'b) Tree.t ->
  nstates:int ->
  nsites:int ->
  transition_probabilities:('b -> Phylo_ctmc.matrix_decomposition) ->
  leaf_state:('l -> int -> int) ->
  root_frequencies:Linear_algebra.vec ->
  float * float_vec
(** All sites of an alignment at once: total and per-site log-likelihoods.  [leaf_state l i] is the
    state of leaf [l] at site [i]; both closures are evaluated eagerly on the host (once per branch,
    once per (leaf, site)). *)

val pruning_sites_masks :
  ('n, 'l, 'b) Tree.t ->
  nstates:int ->
  nsites:int ->
  transition_probabilities:('b -> Phylo_ctmc.matrix_decomposition) ->
  leaf_states:('l -> int -> int64) ->
  root_frequencies:Linear_algebra.vec ->
  float * float_vec
(** [Phylo_ctmc.Missing_values.pruning] / [Phylo_ctmc.Ambiguous.pruning] (lib/phylo_ctmc.ml:145-172, 280-302) for all
    sites: [leaf_states l i] is the bit mask of the states compatible with leaf [l] at site [i] ([0L] = missing). *)

(** [Phylo_b200.Felsenstein.Make] = [Felsenstein.Make] with the reference's own functor arguments and all four of its functions
    (lib/felsenstein.mli:11-48).  [E] is a plain [Site_evolution_model.S]: [E.transition_probability_matrix param t]
    is called once per branch on the host (Padé models included) and the matrices are uploaded; pruning, rescaling
    and the sum over sites run on the device. *)
module Felsenstein : sig
module type Alignment = Phylogenetics.Felsenstein.Alignment

module Make (A : Alphabet.S) (Align : Alignment with type base := A.t)
    (E : Site_evolution_model.S with type mat := A.matrix and type vec := A.vector) : sig
  val felsenstein_single :
    ?shift:(float -> float -> A.vector -> A.vector * float) ->
    E.param -> site:int -> Phylogenetic_tree.t -> Align.t -> float
  (** No underflow prevention (the reference's default [shift]).  A caller-supplied [~shift] closure cannot cross the
      C ABI: [Invalid_argument]. *)
  val felsenstein_single_shift :
    ?threshold:float -> E.param -> site:int -> Phylogenetic_tree.t -> Align.t -> float
  val felsenstein_noshift : E.param -> Phylogenetic_tree.t -> Align.t -> float
  val felsenstein : E.param -> Phylogenetic_tree.t -> Align.t -> float
end

(** The same for models that stage a diagonalisation ([S_with_reduction]): P(t) is rebuilt on the device. *)
module Make_with_reduction (A : Alphabet.S_int) (Align : Alignment with type base := A.t)
    (E : Site_evolution_model.S_with_reduction with type mat := A.matrix and type vec := A.vector) : sig
  val felsenstein : E.param -> Phylogenetic_tree.t -> Align.t -> float
  (** Same type and value as [Felsenstein.Make(A)(Align)(E).felsenstein] (lib/felsenstein.mli:47);
      P(t) of every branch is rebuilt on the device from [E.rate_matrix_reduction]. *)
  val felsenstein_noshift : E.param -> Phylogenetic_tree.t -> Align.t -> float
end
end

(** {2 Low-level access (one call per C-ABI entry point)} *)

val set_tree : ctx -> flat_tree -> branch_lengths:float_vec -> unit
val set_branch_lengths : ctx -> float_vec -> unit
val set_tips : ctx -> tip_matrix -> unit
val set_tips_packed2 : ctx -> tip_matrix -> unit
(** Nucleotides 2-bit packed, [ntaxa] x [ceil(nsites/4)]: site [4b+j] in bits [2j..2j+1] of byte [b]. *)
val set_mode : ctx -> int -> unit
val set_rescaling : ctx -> float -> bool -> unit
val set_root_frequencies : ctx -> Linear_algebra.vec -> unit
val set_categories : ctx -> rates:float_vec -> weights:float_vec -> unit
val set_eigensystem : ctx -> Linear_algebra.mat -> Linear_algebra.vec -> Linear_algebra.mat -> unit
val set_rate_matrix : ctx -> Linear_algebra.mat -> Linear_algebra.vec -> unit
val set_rate_matrix_general : ctx -> Linear_algebra.mat -> unit
val set_pmatrices : ctx -> float_vec -> unit
val loglik : ctx -> ?per_site:float_vec -> unit -> float

(** [Phylo_ctmc.conditional_simulation] for every site at once ([pb_conditional_simulation]):
    [states.{node, site}] after a [loglik] on a context created with [~keep_clv:true]. *)
val conditional_simulation : ctx -> Gsl.Rng.t -> nnodes:int -> nsites:int -> tip_matrix

(** The shifted vectors of [Phylo_ctmc.conditional_likelihoods] (lib/phylo_ctmc.mli:68-73) at one node, for every
    site: [(v, carry)] with [v.{site, state}]; context created with [~keep_clv:true], after a [loglik]. *)
val conditional_likelihood :
  ctx -> ?cat:int -> node:int -> nsites:int -> nstates:int -> unit -> float_mat * float_vec

(** {2 Stand-alone P(t) producers} *)

val transition_matrices_eigen :
  ?device:int -> u:Linear_algebra.mat -> lambda:Linear_algebra.vec -> uinv:Linear_algebra.mat -> float array ->
  Linear_algebra.mat array
(** [Site_evolution_model.Make_diag(..).transition_probability_matrix] for many branch lengths in one launch. *)

val transition_matrices_expm : ?device:int -> Linear_algebra.mat -> float array -> Linear_algebra.mat array
(** [Matrix.expm (Matrix.scal_mul t q)] for many [t] in one launch ([Rate_matrix.transition_probability_matrix]). *)

(** {2 The rest of the C ABI} *)

val set_tip_masks : ctx -> mask_matrix -> unit
val set_tips_text : ctx -> alphabet:int -> tip_matrix -> unit
val alphabet_dna : int
val alphabet_amino_acid : int
val alphabet_codon : int
val set_site_weights : ctx -> float_vec option -> unit
val update_branch_length : ctx -> int -> float -> unit
(** One branch (by the node below it); with [~keep_clv:true] the next [loglik] recomputes only the path to the root. *)
val nodes_recomputed : ctx -> int
val set_option : ctx -> string -> float -> unit
val get_pmatrices : ctx -> ncat:int -> nnodes:int -> nstates:int -> float_vec
val path_name : ctx -> string
val kernel_launches : ctx -> int

val loglik_host : ctx -> ?per_site:float_vec -> tip_matrix -> float
val loglik_host_packed2 : ctx -> ?per_site:float_vec -> tip_matrix -> float
(** One call per evaluation from a host alignment ([pb_loglik_host], [pb_loglik_host_packed2]): the copy to the
    device runs in column chunks while the previous chunk is pruned. *)

val get_dims : ctx -> int * int * int * int * int * int
(** [(nstates, ncat, nsites, ntaxa, nnodes, flags)]: what the stubs check every Bigarray against. *)

(** {2 Multi-GPU}  One host thread drives the GPUs of one box ([pb_create_sharded]): the alignment's columns are
    sharded over [devices], tree and model replicated, the total combined by an NCCL all-reduce inside the library
    ([Failure] on any NCCL error). *)
module Sharded : sig
  type t
  val reduce_nccl : int
  val reduce_ordered : int
  val create : devices:int array -> ?ncat:int -> nstates:int -> nsites:int -> ntaxa:int -> unit -> t
  val ndev : t -> int
  val set_reduce : t -> int -> unit
  val set_rescaling : t -> float -> bool -> unit
  val set_tree : t -> flat_tree -> branch_lengths:float_vec -> unit
  val set_branch_lengths : t -> float_vec -> unit
  val set_root_frequencies : t -> Linear_algebra.vec -> unit
  val set_categories : t -> rates:float_vec -> weights:float_vec -> unit
  val set_eigensystem : t -> Linear_algebra.mat -> Linear_algebra.vec -> Linear_algebra.mat -> unit
  val set_rate_matrix : t -> Linear_algebra.mat -> Linear_algebra.vec -> unit
  val set_pmatrices : t -> float_vec -> unit
  val set_tips : t -> tip_matrix -> unit
  (** The FULL alignment; every device copies its own columns. *)
  val loglik : t -> ?per_site:float_vec -> unit -> float
  val loglik_host : t -> ?per_site:float_vec -> ?packed2:bool -> tip_matrix -> float
end

val pin : ('a, 'b, Bigarray.c_layout) Bigarray.Array2.t -> unit
val unpin : ('a, 'b, Bigarray.c_layout) Bigarray.Array2.t -> unit
(** Page-lock / release a Bigarray's payload so that those copies are asynchronous. *)
