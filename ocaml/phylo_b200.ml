(* OCaml side of the drop-in: evaluates the reference's closures eagerly, flattens the tree and
   calls the C stubs (phylo_b200_stubs.c).  Not compiled here (no OCaml toolchain); the stubs are tested against a mock runtime: see INTEGRATION.md. *)
open Core
open Phylogenetics

type ctx
type int32_vec = (int32, Bigarray.int32_elt, Bigarray.c_layout) Bigarray.Array1.t
type float_vec = (float, Bigarray.float64_elt, Bigarray.c_layout) Bigarray.Array1.t
type tip_matrix = (int, Bigarray.int8_unsigned_elt, Bigarray.c_layout) Bigarray.Array2.t
type float_mat = (float, Bigarray.float64_elt, Bigarray.c_layout) Bigarray.Array2.t
type mask_matrix = (int64, Bigarray.int64_elt, Bigarray.c_layout) Bigarray.Array2.t

type flat_tree = { root : int ; child_ptr : int32_vec ; child_idx : int32_vec ; tip_row : int32_vec }

external create_ : int -> int -> int -> int -> int -> int -> ctx = "pb_stub_create_bytecode" "pb_stub_create"
external set_mode : ctx -> int -> unit = "pb_stub_set_mode"
external set_rescaling : ctx -> float -> bool -> unit = "pb_stub_set_rescaling"
external set_tree_ : ctx -> int -> int32_vec -> int32_vec -> int32_vec -> float_vec -> unit
  = "pb_stub_set_tree_bytecode" "pb_stub_set_tree"
external set_branch_lengths : ctx -> float_vec -> unit = "pb_stub_set_branch_lengths"
external set_tips : ctx -> tip_matrix -> unit = "pb_stub_set_tips"
external set_tips_packed2 : ctx -> tip_matrix -> unit = "pb_stub_set_tips_packed2"
external set_root_frequencies : ctx -> Linear_algebra.vec -> unit = "pb_stub_set_root_frequencies"
external set_categories_ : ctx -> float_vec -> float_vec -> unit = "pb_stub_set_categories"
external set_eigensystem : ctx -> Linear_algebra.mat -> Linear_algebra.vec -> Linear_algebra.mat -> unit
  = "pb_stub_set_eigensystem"
external set_rate_matrix : ctx -> Linear_algebra.mat -> Linear_algebra.vec -> unit = "pb_stub_set_rate_matrix"
external set_rate_matrix_general : ctx -> Linear_algebra.mat -> unit = "pb_stub_set_rate_matrix_general"
external set_pmatrices : ctx -> float_vec -> unit = "pb_stub_set_pmatrices"
external loglik_ : ctx -> float_vec -> float = "pb_stub_loglik"
external conditional_simulation_ : ctx -> int64 -> tip_matrix -> unit = "pb_stub_conditional_simulation"
external get_clv_ : ctx -> int -> int -> float_mat -> float_vec -> unit = "pb_stub_get_clv"
external transition_matrices_eigen_ :
  int -> Linear_algebra.mat -> Linear_algebra.vec -> Linear_algebra.mat -> float_vec -> float_vec -> unit
  = "pb_stub_transition_matrices_eigen_bytecode" "pb_stub_transition_matrices_eigen"
external transition_matrices_expm_ : int -> Linear_algebra.mat -> float_vec -> float_vec -> unit
  = "pb_stub_transition_matrices_expm"
external set_tip_masks : ctx -> mask_matrix -> unit = "pb_stub_set_tip_masks"
external set_tips_text_ : ctx -> tip_matrix -> int -> unit = "pb_stub_set_tips_text"
external set_site_weights_ : ctx -> float_vec -> unit = "pb_stub_set_site_weights"
external update_branch_length : ctx -> int -> float -> unit = "pb_stub_update_branch_length"
external nodes_recomputed : ctx -> int = "pb_stub_nodes_recomputed"
external set_option : ctx -> string -> float -> unit = "pb_stub_set_option"
external get_pmatrices_ : ctx -> float_vec -> unit = "pb_stub_get_pmatrices"
external path_name : ctx -> string = "pb_stub_path_name"
external kernel_launches : ctx -> int = "pb_stub_kernel_launches"
external loglik_host_ : ctx -> tip_matrix -> float_vec -> float = "pb_stub_loglik_host"
external loglik_host_packed2_ : ctx -> tip_matrix -> float_vec -> float = "pb_stub_loglik_host_packed2"
external pin : ('a, 'b, Bigarray.c_layout) Bigarray.Array2.t -> unit = "pb_stub_pin"
external unpin : ('a, 'b, Bigarray.c_layout) Bigarray.Array2.t -> unit = "pb_stub_unpin"
external get_dims : ctx -> int * int * int * int * int * int = "pb_stub_get_dims"

(* multi-GPU: one host thread, the alignment's columns sharded over the devices, NCCL all-reduce of the scalar *)
type sharded
external create_sharded_ : int32_vec -> int -> int -> int -> int -> int -> sharded
  = "pb_stub_create_sharded_bytecode" "pb_stub_create_sharded"
external sharded_ndev : sharded -> int = "pb_stub_sharded_ndev"
external sharded_set_reduce : sharded -> int -> unit = "pb_stub_sharded_set_reduce"
external sharded_set_rescaling : sharded -> float -> bool -> unit = "pb_stub_sharded_set_rescaling"
external sharded_set_tree_ : sharded -> int -> int32_vec -> int32_vec -> int32_vec -> float_vec -> unit
  = "pb_stub_sharded_set_tree_bytecode" "pb_stub_sharded_set_tree"
external sharded_set_branch_lengths : sharded -> float_vec -> unit = "pb_stub_sharded_set_branch_lengths"
external sharded_set_root_frequencies : sharded -> Linear_algebra.vec -> unit = "pb_stub_sharded_set_root_frequencies"
external sharded_set_categories_ : sharded -> float_vec -> float_vec -> unit = "pb_stub_sharded_set_categories"
external sharded_set_eigensystem : sharded -> Linear_algebra.mat -> Linear_algebra.vec -> Linear_algebra.mat -> unit
  = "pb_stub_sharded_set_eigensystem"
external sharded_set_rate_matrix : sharded -> Linear_algebra.mat -> Linear_algebra.vec -> unit = "pb_stub_sharded_set_rate_matrix"
external sharded_set_pmatrices : sharded -> float_vec -> unit = "pb_stub_sharded_set_pmatrices"
external sharded_set_tips : sharded -> tip_matrix -> unit = "pb_stub_sharded_set_tips"
external sharded_loglik_ : sharded -> float_vec -> float = "pb_stub_sharded_loglik"
external sharded_loglik_host_ : sharded -> tip_matrix -> bool -> float_vec -> float = "pb_stub_sharded_loglik_host"

let keep_clv_flag = 1
let mode_phylo_ctmc = 0

let create ?(device = 0) ?(ncat = 1) ?(keep_clv = false) ~nstates ~nsites ~ntaxa () =
  create_ device nstates ncat nsites ntaxa (if keep_clv then keep_clv_flag else 0)

let set_tree ctx t ~branch_lengths = set_tree_ ctx t.root t.child_ptr t.child_idx t.tip_row branch_lengths
let set_categories ctx ~rates ~weights = set_categories_ ctx rates weights

(* Phylo_ctmc.conditional_simulation for every site: states.{node, site}; needs ~keep_clv:true and a
   prior [loglik].  The seed is drawn from the caller's Gsl.Rng.t so that runs stay reproducible. *)
let conditional_simulation ctx rng ~nnodes ~nsites =
  let states = Bigarray.(Array2.create int8_unsigned c_layout nnodes nsites) in
  let seed = Int64.of_int (Gsl.Rng.uniform_int rng (1 lsl 30)) in
  conditional_simulation_ ctx seed states ;
  states

(* The shifted vectors (v, carry) of one node for every site (lib/phylo_ctmc.ml:26): v.{site, state}, carry.{site};
   needs ~keep_clv:true and a prior [loglik]. *)
let conditional_likelihood ctx ?(cat = 0) ~node ~nsites ~nstates () =
  let v = Bigarray.(Array2.create float64 c_layout nsites nstates) in
  let carry = Bigarray.(Array1.create float64 c_layout nsites) in
  get_clv_ ctx node cat v carry ;
  v, carry

(* P(t) for many t in one launch; the k-th matrix is returned as a Lacaml matrix (column-major as it comes back). *)
let unpack_matrices ~n ~count (out : float_vec) =
  Array.init count ~f:(fun k ->
      let m = Lacaml.D.Mat.create n n in
      for j = 1 to n do for i = 1 to n do m.{i, j} <- out.{k * n * n + (j - 1) * n + (i - 1)} done done ;
      m)

let float_vec_of_array a = Bigarray.(Array1.of_array float64 c_layout a)

(* Site_evolution_model.Make_diag(..).transition_probability_matrix (lib/site_evolution_model.ml:43-47) for every t *)
let transition_matrices_eigen ?(device = 0) ~u ~lambda ~uinv ts =
  let n = Lacaml.D.Vec.dim lambda and count = Array.length ts in
  let out = Bigarray.(Array1.create float64 c_layout (count * n * n)) in
  transition_matrices_eigen_ device u lambda uinv (float_vec_of_array ts) out ;
  unpack_matrices ~n ~count out

(* Matrix.expm (t * q) for every t (lib/linear_algebra.ml:308-404; lib/rate_matrix.ml:132-138) *)
let transition_matrices_expm ?(device = 0) q ts =
  let n = Lacaml.D.Mat.dim1 q and count = Array.length ts in
  let out = Bigarray.(Array1.create float64 c_layout (count * n * n)) in
  transition_matrices_expm_ device q (float_vec_of_array ts) out ;
  unpack_matrices ~n ~count out

let empty_vec = Bigarray.(Array1.create float64 c_layout 0)
let loglik ctx ?(per_site = empty_vec) () = loglik_ ctx per_site

(* One call per evaluation from a host alignment: the copy runs in column chunks while the previous chunk is pruned.
   [pin] the tip matrix once if it is evaluated repeatedly. *)
let loglik_host ctx ?(per_site = empty_vec) tips = loglik_host_ ctx tips per_site
let loglik_host_packed2 ctx ?(per_site = empty_vec) tips = loglik_host_packed2_ ctx tips per_site

let set_site_weights ctx = function
  | Some w -> set_site_weights_ ctx w
  | None -> set_site_weights_ ctx empty_vec

let alphabet_dna = 0 and alphabet_amino_acid = 1 and alphabet_codon = 2
let set_tips_text ctx ~alphabet text = set_tips_text_ ctx text alphabet

let get_pmatrices ctx ~ncat ~nnodes ~nstates =
  let p = Bigarray.(Array1.create float64 c_layout (ncat * nnodes * nstates * nstates)) in
  get_pmatrices_ ctx p ;
  p

(* Pre-order numbering of a Tree.t (lib/tree.ml:3-13). *)
let flatten (t : _ Tree.t) =
  let nodes = Queue.create () in   (* (children ids in branch order, leaf payload option, branch payload option) *)
  let rec visit t bdata =
    let id = Queue.length nodes in
    Queue.enqueue nodes (ref [], None, bdata) ;
    (match t with
     | Tree.Leaf l -> Queue.set nodes id (ref [], Some l, bdata)
     | Node n ->
       let kids = List1.map n.branches ~f:(fun (Tree.Branch b) -> visit b.tip (Some b.data)) |> List1.to_list in
       let (r, _, _) = Queue.get nodes id in
       r := kids) ;
    id
  in
  let root = visit t None in
  let n = Queue.length nodes in
  let child_ptr = Bigarray.(Array1.create int32 c_layout (n + 1)) in
  let child_idx = Bigarray.(Array1.create int32 c_layout (max 1 (n - 1))) in
  let tip_row = Bigarray.(Array1.create int32 c_layout n) in
  let leaves = Queue.create () and branches = Array.create ~len:n None in
  let k = ref 0 in
  child_ptr.{0} <- 0l ;
  Queue.iteri nodes ~f:(fun v (kids, leaf, bdata) ->
      branches.(v) <- bdata ;
      (match leaf with
       | Some l -> tip_row.{v} <- Int32.of_int_exn (Queue.length leaves) ; Queue.enqueue leaves l
       | None -> tip_row.{v} <- (-1l)) ;
      List.iter !kids ~f:(fun c -> child_idx.{!k} <- Int32.of_int_exn c ; incr k) ;
      child_ptr.{v + 1} <- Int32.of_int_exn !k) ;
  { root ; child_ptr ; child_idx ; tip_row }, Queue.to_array leaves, branches

let pruning_sites t ~nstates ~nsites ~transition_probabilities ~leaf_state ~root_frequencies =
  let flat, leaves, branches = flatten t in
  let nnodes = Bigarray.Array1.dim flat.tip_row and ntaxa = Array.length leaves in
  let ctx = create ~nstates ~nsites ~ntaxa () in
  let zero = Bigarray.(Array1.create float64 c_layout nnodes) in
  Bigarray.Array1.fill zero 0. ;
  set_tree ctx flat ~branch_lengths:zero ;
  (* transition_probabilities is called once per branch; decompositions are multiplied out on the host
     (Phylo_ctmc.matrix_decomposition_reduce, lib/phylo_ctmc.ml:13-24) *)
  let p = Bigarray.(Array1.create float64 c_layout (nnodes * nstates * nstates)) in
  Bigarray.Array1.fill p 0. ;
  Array.iteri branches ~f:(fun v -> function
      | None -> ()
      | Some b ->
        let m = Phylo_ctmc.matrix_decomposition_reduce ~dim:nstates (transition_probabilities b) in
        for j = 0 to nstates - 1 do
          for i = 0 to nstates - 1 do
            p.{v * nstates * nstates + j * nstates + i} <- Linear_algebra.Matrix.get m i j   (* column-major *)
          done
        done) ;
  set_pmatrices ctx p ;
  set_mode ctx mode_phylo_ctmc ;
  set_root_frequencies ctx root_frequencies ;
  let per_site = Bigarray.(Array1.create float64 c_layout nsites) in
  (* leaf_state is evaluated once per (leaf, site) either way; nucleotides are written 2-bit packed, 4 sites per
     byte, which is what crosses PCIe; the evaluation is ONE call that overlaps the copy with the pruning
     (pb_loglik_host_packed2 / pb_loglik_host) *)
  let tips = if nstates = 4 then begin
    let nbytes = (nsites + 3) / 4 in
    let tips = Bigarray.(Array2.create int8_unsigned c_layout ntaxa nbytes) in
    Bigarray.Array2.fill tips 0 ;
    Array.iteri leaves ~f:(fun row l ->
        for s = 0 to nsites - 1 do
          tips.{row, s / 4} <- tips.{row, s / 4} lor (leaf_state l s lsl (2 * (s land 3)))
        done) ;
    `Packed tips
  end
  else begin
    let tips = Bigarray.(Array2.create int8_unsigned c_layout ntaxa nsites) in
    Array.iteri leaves ~f:(fun row l -> for s = 0 to nsites - 1 do tips.{row, s} <- leaf_state l s done) ;
    `Bytes tips
  end in
  let total = match tips with
    | `Packed t -> loglik_host_packed2 ctx ~per_site t
    | `Bytes t -> loglik_host ctx ~per_site t
  in
  total, per_site

(* Phylo_ctmc.Missing_values.pruning / Ambiguous.pruning (lib/phylo_ctmc.ml:145-172, 280-302) for every site:
   [leaf_states l i] is the set of states compatible with the observation at leaf [l], site [i], as a bit mask
   (bit k = state k; 0 = missing).  [Missing_values]: [Some s -> 1 lsl s | None -> 0]; [Ambiguous]: the mask of
   the states for which the reference's [leaf_state l s] is true. *)
let pruning_sites_masks t ~nstates ~nsites ~transition_probabilities ~leaf_states ~root_frequencies =
  let flat, leaves, branches = flatten t in
  let nnodes = Bigarray.Array1.dim flat.tip_row and ntaxa = Array.length leaves in
  let ctx = create ~nstates ~nsites ~ntaxa () in
  let zero = Bigarray.(Array1.create float64 c_layout nnodes) in
  Bigarray.Array1.fill zero 0. ;
  set_tree ctx flat ~branch_lengths:zero ;
  let p = Bigarray.(Array1.create float64 c_layout (nnodes * nstates * nstates)) in
  Bigarray.Array1.fill p 0. ;
  Array.iteri branches ~f:(fun v -> function
      | None -> ()
      | Some b ->
        let m = Phylo_ctmc.matrix_decomposition_reduce ~dim:nstates (transition_probabilities b) in
        for j = 0 to nstates - 1 do
          for i = 0 to nstates - 1 do
            p.{v * nstates * nstates + j * nstates + i} <- Linear_algebra.Matrix.get m i j
          done
        done) ;
  set_pmatrices ctx p ;
  let masks = Bigarray.(Array2.create int64 c_layout ntaxa nsites) in
  Array.iteri leaves ~f:(fun row l -> for s = 0 to nsites - 1 do masks.{row, s} <- leaf_states l s done) ;
  set_tip_masks ctx masks ;
  set_mode ctx mode_phylo_ctmc ;
  set_root_frequencies ctx root_frequencies ;
  let per_site = Bigarray.(Array1.create float64 c_layout nsites) in
  let total = loglik ctx ~per_site () in
  total, per_site

let pruning t ~nstates ~transition_probabilities ~leaf_state ~root_frequencies =
  fst (pruning_sites t ~nstates ~nsites:1 ~transition_probabilities
         ~leaf_state:(fun l _ -> leaf_state l) ~root_frequencies)

module Sharded = struct
  type t = sharded
  let reduce_nccl = 0 and reduce_ordered = 1
  let create ~devices ?(ncat = 1) ~nstates ~nsites ~ntaxa () =
    let d = Bigarray.(Array1.of_array int32 c_layout (Array.map devices ~f:Int32.of_int_exn)) in
    create_sharded_ d nstates ncat nsites ntaxa 0
  let ndev = sharded_ndev
  let set_reduce = sharded_set_reduce
  let set_rescaling = sharded_set_rescaling
  let set_tree s t ~branch_lengths = sharded_set_tree_ s t.root t.child_ptr t.child_idx t.tip_row branch_lengths
  let set_branch_lengths = sharded_set_branch_lengths
  let set_root_frequencies = sharded_set_root_frequencies
  let set_categories s ~rates ~weights = sharded_set_categories_ s rates weights
  let set_eigensystem = sharded_set_eigensystem
  let set_rate_matrix = sharded_set_rate_matrix
  let set_pmatrices = sharded_set_pmatrices
  let set_tips = sharded_set_tips
  let loglik s ?(per_site = empty_vec) () = sharded_loglik_ s per_site
  let loglik_host s ?(per_site = empty_vec) ?(packed2 = false) tips = sharded_loglik_host_ s tips packed2 per_site
end

(* Felsenstein.Make (lib/felsenstein.mli:11-48) with the reference's own functor arguments: E is a plain
   Site_evolution_model.S, so P(t) of every branch is E.transition_probability_matrix param t, called ONCE PER BRANCH on
   the host (the reference calls it once per branch per site) and uploaded ([`Mat]-style, pb_set_pmatrices); the
   pruning, the rescaling and the sum over sites run on the device. *)
module Felsenstein = struct
module type Alignment = Phylogenetics.Felsenstein.Alignment

module Make (A : Alphabet.S) (Align : Alignment with type base := A.t)
    (E : Site_evolution_model.S with type mat := A.matrix and type vec := A.vector) =
struct
  let symbols = Array.of_list A.all          (* A.to_int is the state index (lib/alphabet.ml) *)

  (* sites [first, first + nsites) of the alignment; threshold = -infinity: no rescaling at all *)
  let evaluate ~threshold ~first ~nsites param tree align =
    let spec_eMt = E.transition_probability_matrix param in       (* as lib/felsenstein.ml:24: specialised once *)
    let t = Phylogenetic_tree.to_tree tree in
    let flat, leaves, branches = flatten t in
    let nnodes = Bigarray.Array1.dim flat.tip_row and ntaxa = Array.length leaves in
    let n = A.card in
    let ctx = create ~nstates:n ~nsites ~ntaxa () in
    let zero = Bigarray.(Array1.create float64 c_layout nnodes) in
    Bigarray.Array1.fill zero 0. ;
    set_tree ctx flat ~branch_lengths:zero ;
    let p = Bigarray.(Array1.create float64 c_layout (nnodes * n * n)) in
    Bigarray.Array1.fill p 0. ;
    Array.iteri branches ~f:(fun v -> function
        | None -> ()
        | Some l ->
          let m = spec_eMt l in
          Array.iter symbols ~f:(fun x ->
              Array.iter symbols ~f:(fun y ->
                  p.{v * n * n + A.to_int y * n + A.to_int x} <- A.Matrix.get m x y))) ;     (* column-major *)
    set_pmatrices ctx p ;
    let tips = Bigarray.(Array2.create int8_unsigned c_layout ntaxa nsites) in
    Array.iteri leaves ~f:(fun row (_, index) ->
        for s = 0 to nsites - 1 do tips.{row, s} <- A.to_int (Align.get_base align ~seq:index ~pos:(first + s)) done) ;
    set_tips ctx tips ;
    set_rescaling ctx threshold false ;     (* shift_normal, no shift after x pi: lib/felsenstein.ml:47-48,57-64 *)
    let pi = Lacaml.D.Vec.create n in
    let stat = E.stationary_distribution param in
    Array.iter symbols ~f:(fun x -> pi.{A.to_int x + 1} <- A.Vector.get stat x) ;
    set_root_frequencies ctx pi ;
    loglik ctx ()

  (* lib/felsenstein.mli:19-26.  [?shift] is an arbitrary OCaml closure applied at every node: it cannot cross the C
     ABI, so a caller-supplied one is refused; without it the reference does not rescale at all, and neither do we. *)
  let felsenstein_single ?shift param ~site tree align =
    (match shift with
     | Some _ -> invalid_arg "Phylo_b200.Felsenstein.Make.felsenstein_single: a custom ~shift closure cannot run on the device; use felsenstein_single_shift ?threshold"
     | None -> ()) ;
    evaluate ~threshold:Float.neg_infinity ~first:site ~nsites:1 param tree align

  (* lib/felsenstein.mli:28-35 *)
  let felsenstein_single_shift ?(threshold = 0.0000001) param ~site tree align =
    evaluate ~threshold ~first:site ~nsites:1 param tree align

  let felsenstein_noshift param tree align =
    evaluate ~threshold:Float.neg_infinity ~first:0 ~nsites:(Align.length align) param tree align
  let felsenstein param tree align =
    evaluate ~threshold:0.0000001 ~first:0 ~nsites:(Align.length align) param tree align
end

(* The same for models that stage a diagonalisation (Site_evolution_model.S_with_reduction: JC69, K80, Make_diag):
   P(t) of every branch is rebuilt ON THE DEVICE from the eigensystem (kernel pmat_eigen). *)
module Make_with_reduction (A : Alphabet.S_int) (Align : Alignment with type base := A.t)
    (E : Site_evolution_model.S_with_reduction with type mat := A.matrix and type vec := A.vector) =
struct
  let evaluate ~threshold param tree align =
    let t = Phylogenetic_tree.to_tree tree in
    let flat, leaves, branches = flatten t in
    let nnodes = Bigarray.Array1.dim flat.tip_row and ntaxa = Array.length leaves in
    let nsites = Align.length align in
    let ctx = create ~nstates:A.card ~nsites ~ntaxa () in
    let bl = Bigarray.(Array1.create float64 c_layout nnodes) in
    Array.iteri branches ~f:(fun v b -> bl.{v} <- Option.value b ~default:0.) ;
    set_tree ctx flat ~branch_lengths:bl ;
    let u, lambda, uinv = E.rate_matrix_reduction param in
    set_eigensystem ctx (u :> Linear_algebra.mat) (lambda :> Linear_algebra.vec) (uinv :> Linear_algebra.mat) ;
    let tips = Bigarray.(Array2.create int8_unsigned c_layout ntaxa nsites) in
    Array.iteri leaves ~f:(fun row (_, index) ->
        for s = 0 to nsites - 1 do tips.{row, s} <- A.to_int (Align.get_base align ~seq:index ~pos:s) done) ;
    set_tips ctx tips ;
    set_rescaling ctx threshold false ;     (* shift_normal, no shift after x pi: lib/felsenstein.ml:47-48,57-64 *)
    set_root_frequencies ctx (E.stationary_distribution param :> Linear_algebra.vec) ;
    loglik ctx ()

  let felsenstein param tree align = evaluate ~threshold:0.0000001 param tree align
  let felsenstein_noshift param tree align = evaluate ~threshold:Float.neg_infinity param tree align
end
end
