"""Oracle restatement of Phylo_ctmc / Felsenstein pruning, one site at a time.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pure-Python recursion over
numpy vectors: use it on small cases; oracle/pruning_oracle.c is the same
algorithm in C for large ones.

Follows lib/phylo_ctmc.ml:4-120,145-197,232-332 and lib/felsenstein.ml:17-78.
"""
import math

import numpy as np


# ------------------------------------------------------------------ Tree.t
class Leaf:
    """Tree.Leaf, lib/tree.ml:3-13."""
    __slots__ = ("data",)

    def __init__(self, data):
        self.data = data


class Branch:
    __slots__ = ("data", "tip")

    def __init__(self, data, tip):
        self.data = data
        self.tip = tip


class Node:
    __slots__ = ("data", "branches")

    def __init__(self, data, branches):
        if len(branches) < 1:
            raise ValueError("List1: a node needs at least one branch")
        self.data = data
        self.branches = list(branches)


def binary_node(data, left, right):
    return Node(data, [left, right])


# ------------------------------------------------- matrix_decomposition, SV
def matrix_decomposition_reduce(dim, decomp):
    """lib/phylo_ctmc.ml:13-24 — multiply the factor list out, rightmost first."""
    if len(decomp) == 0:
        return np.eye(dim)

    def dense(op):
        kind, x = op
        if kind == "Mat":
            return np.asarray(x)
        if kind == "Transpose":
            return np.asarray(x).T
        return np.diag(x)
    if len(decomp) == 1:
        return dense(decomp[0])
    return dense(decomp[0]) @ matrix_decomposition_reduce(dim, decomp[1:])


def sv_shift(v, carry, threshold=1e-6):
    """SV.shift, lib/phylo_ctmc.ml:33-40."""
    if np.min(v) > threshold:
        return v, carry
    mv = np.max(v)
    # An all-zero vector gives 1/0 = inf and 0*inf = NaN, as in the reference.
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.float64(1.0) / mv * v, carry + float(np.log(mv))


def _apply_op(op, v):
    """SV.decomp_vec_mul_aux, lib/phylo_ctmc.ml:42-45."""
    kind, x = op
    if kind == "Mat":
        return matvec(x, v)
    if kind == "Transpose":
        return matvec(np.asarray(x).T, v)
    return np.asarray(x) * v


def matvec(m, v):
    """Matrix.apply = dgemv, lib/linear_algebra.ml:222.

    Sequential accumulation over j (reference-BLAS axpy order); the summation
    order inside an optimised BLAS is implementation-defined.
    """
    m = np.asarray(m)
    out = np.zeros(m.shape[0])
    for j in range(m.shape[1]):
        out = out + m[:, j] * v[j]
    return out


def sv_decomp_vec_mul(decomp, sv):
    """SV.decomp_vec_mul, lib/phylo_ctmc.ml:47-48 (fold_right: last factor first)."""
    v, carry = sv
    for op in reversed(decomp):
        v = _apply_op(op, v)
    return v, carry


def sv_mul(a, b, threshold=1e-6):
    """SV.mul, lib/phylo_ctmc.ml:76-78."""
    return sv_shift(a[0] * b[0], a[1] + b[1], threshold)


def sv_scal_vec_mul(alpha, sv):
    """SV.scal_vec_mul, lib/phylo_ctmc.ml:53-54."""
    return alpha * sv[0], sv[1]


def sv_add(a, b):
    """SV.add, lib/phylo_ctmc.ml:56-62."""
    (v1, c1), (v2, c2) = (a, b) if a[1] > b[1] else (b, a)
    return v1 + math.exp(c2 - c1) * v2, c1


def indicator(i, n):
    """lib/phylo_ctmc.ml:81."""
    v = np.zeros(n)
    v[i] = 1.0
    return v


def _as_decomp(tp):
    """Accept either a matrix_decomposition list or a bare matrix."""
    if isinstance(tp, list):
        return tp
    return [("Mat", tp)]


def _reduce_left(items, f):
    acc = items[0]
    for x in items[1:]:
        acc = f(acc, x)
    return acc


# ------------------------------------------------------------------ pruning
def pruning(t, nstates, transition_probabilities, leaf_state, root_frequencies,
            threshold=1e-6):
    """Phylo_ctmc.pruning, lib/phylo_ctmc.ml:83-98.

    `threshold` is SV.shift's default (1e-6); the reference does not expose it
    through `pruning`, the argument exists only so tests can probe the rule.
    """
    def tree(t):
        if isinstance(t, Leaf):
            return indicator(leaf_state(t.data), nstates), 0.0
        terms = [sv_decomp_vec_mul(_as_decomp(transition_probabilities(b.data)), tree(b.tip))
                 for b in t.branches]
        return _reduce_left(terms, lambda x, y: sv_mul(x, y, threshold))
    v, carry = sv_mul(tree(t), (np.asarray(root_frequencies, dtype=np.float64), 0.0), threshold)
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.log(np.sum(v))) + carry


def conditional_likelihoods(t, nstates, leaf_state, transition_probabilities, threshold=1e-6):
    """Phylo_ctmc.conditional_likelihoods, lib/phylo_ctmc.ml:100-120.

    Returns a tree of the same shape: Leaf(state) / Node((v, carry), branches)
    with Branch((branch_data, P), subtree).
    """
    def tree(t):
        if isinstance(t, Leaf):
            s = leaf_state(t.data)
            return Leaf(s), (indicator(s, nstates), 0.0)
        res = [branch(b) for b in t.branches]
        cl = _reduce_left([c for _, c in res], lambda x, y: sv_mul(x, y, threshold))
        return Node(cl, [b for b, _ in res]), cl

    def branch(b):
        mat = transition_probabilities(b.data)
        tip, tip_cl = tree(b.tip)
        cl = (matvec(mat, tip_cl[0]), tip_cl[1])          # SV.mat_vec_mul :50-51
        return Branch((b.data, mat), tip), cl
    return tree(t)[0]


def missing_values_pruning(t, nstates, transition_probabilities, leaf_state, root_frequencies):
    """Phylo_ctmc.Missing_values.pruning, lib/phylo_ctmc.ml:145-172.

    leaf_state returns None for a missing observation; missing children are
    skipped (no product, no shift); an all-missing site has log-likelihood 0.
    """
    def maybe_mul(x, y):
        if x is None:
            return y
        if y is None:
            return x
        return sv_mul(x, y)

    def tree(t):
        if isinstance(t, Leaf):
            s = leaf_state(t.data)
            return None if s is None else (indicator(s, nstates), 0.0)
        acc = None
        for b in t.branches:
            sub = tree(b.tip)
            term = None if sub is None else sv_decomp_vec_mul(
                _as_decomp(transition_probabilities(b.data)), sub)
            acc = maybe_mul(acc, term)
        return acc
    res = tree(t)
    if res is None:
        return 0.0
    v, carry = sv_mul(res, (np.asarray(root_frequencies, dtype=np.float64), 0.0))
    return float(np.log(np.sum(v))) + carry


def ambiguous_pruning(t, nstates, transition_probabilities, leaf_state, root_frequencies):
    """Phylo_ctmc.Ambiguous.pruning, lib/phylo_ctmc.ml:280-307.

    leaf_state l is a predicate state -> bool; the tip vector is its 0/1
    indicator (:234-242); an empty set behaves like a missing observation.
    """
    def tree(t):
        if isinstance(t, Leaf):
            pred = leaf_state(t.data)
            v = np.array([1.0 if pred(i) else 0.0 for i in range(nstates)])
            return (v, 0.0) if v.any() else None
        terms = []
        for b in t.branches:
            sub = tree(b.tip)
            if sub is not None:
                terms.append(sv_decomp_vec_mul(_as_decomp(transition_probabilities(b.data)), sub))
        if not terms:
            return None
        return _reduce_left(terms, sv_mul)
    res = tree(t)
    if res is None:
        return 0.0
    v, carry = sv_mul(res, (np.asarray(root_frequencies, dtype=np.float64), 0.0))
    return float(np.log(np.sum(v))) + carry


def ambiguous_conditional_likelihoods(t, nstates, leaf_state, transition_probabilities):
    """Phylo_ctmc.Ambiguous.conditional_likelihoods, lib/phylo_ctmc.ml:304-332."""
    def node(t):
        if isinstance(t, Leaf):
            pred = leaf_state(t.data)
            states = [i for i in range(nstates) if pred(i)]
            if not states:
                return None
            v = np.zeros(nstates)
            v[states] = 1.0
            return Leaf(states), (v, 0.0)
        res = [r for r in (branch(b) for b in t.branches) if r is not None]
        if not res:
            return None
        cl = _reduce_left([c for _, c in res], sv_mul)
        return Node(cl, [b for b, _ in res]), cl

    def branch(b):
        mat = transition_probabilities(b.data)
        sub = node(b.tip)
        if sub is None:
            return None
        tip, tip_cl = sub
        return Branch((b.data, mat), tip), (matvec(mat, tip_cl[0]), tip_cl[1])
    res = node(t)
    return Leaf(list(range(nstates))) if res is None else res[0]


# -------------------------------------------------------------- Felsenstein
class PhyloLeaf:
    """Phylogenetic_tree.Leaf {index}, lib/phylogenetic_tree.ml:6-10."""
    __slots__ = ("index",)

    def __init__(self, index):
        self.index = index


class PhyloNode:
    """Phylogenetic_tree.Node {left = (len, t); right = (len, t)}."""
    __slots__ = ("left", "right")

    def __init__(self, left, right):
        self.left = left
        self.right = right


def of_preorder(s):
    """Phylogenetic_tree.of_preorder, lib/phylogenetic_tree.ml:72-97.

    Tokens with a '.' are branch lengths (two per node, then left and right
    subtrees); integers are leaves named T<i>.
    """
    toks = s.split(";")
    pos = 0

    def full():
        nonlocal pos
        if pos >= len(toks):
            raise ValueError("Empty token list.")
        tok = toks[pos]
        pos += 1
        if "." in tok:
            f1 = float(tok)
            if pos >= len(toks) or "." not in toks[pos]:
                raise ValueError("Expected second float")
            f2 = float(toks[pos])
            pos += 1
            l = full()
            r = full()
            return PhyloNode((f1, l), (f2, r))
        return PhyloLeaf("T%d" % int(tok))
    return full()


def to_tree(t):
    """Phylogenetic_tree.to_tree, lib/phylogenetic_tree.ml:137-144."""
    if isinstance(t, PhyloLeaf):
        return Leaf((None, t.index))
    return binary_node(None, Branch(t.left[0], to_tree(t.left[1])),
                       Branch(t.right[0], to_tree(t.right[1])))


def shift_normal(thre, acc1, acc2, v):
    """Felsenstein.shift_normal, lib/felsenstein.ml:57-62."""
    if np.min(v) > thre:
        return v, acc1 + acc2
    mv = np.max(v)
    return (1.0 / mv) * v, acc1 + acc2 + math.log(mv)


def felsenstein_single(tpm, stationary, get_base, nstates, tree, site, shift=None):
    """Felsenstein.felsenstein_single, lib/felsenstein.ml:22-48.

    tpm: branch length -> P(t); get_base(index, site) -> state int.
    The default `shift` never rescales (:22).  Note: no shift after x pi (:47-48).
    """
    if shift is None:
        def shift(a, b, v):
            return v, 0.0

    def aux(t):
        if isinstance(t, PhyloLeaf):
            return shift(0.0, 0.0, indicator(get_base(t.index, site), nstates))
        (l1, t1), (l2, t2) = t.left, t.right
        (vl, sl), (vr, sr) = aux(t1), aux(t2)
        return shift(sl, sr, matvec(tpm(l1), vl) * matvec(tpm(l2), vr))
    res_vec, res_shift = aux(tree)
    return math.log(float(np.sum(res_vec * stationary))) + res_shift


def felsenstein(tpm, stationary, get_base, nstates, tree, length, threshold=0.0000001):
    """Felsenstein.felsenstein = multisite (felsenstein_single_shift), :64-78.

    Sequential sum, site 0 first: acc' = f(site) + acc (:73-75).
    """
    acc = 0.0
    for x in range(length):
        acc = felsenstein_single(tpm, stationary, get_base, nstates, tree, x,
                                 shift=lambda a, b, v: shift_normal(threshold, a, b, v)) + acc
    return acc


def felsenstein_noshift(tpm, stationary, get_base, nstates, tree, length):
    """Felsenstein.felsenstein_noshift, lib/felsenstein.ml:77."""
    acc = 0.0
    for x in range(length):
        acc = felsenstein_single(tpm, stationary, get_base, nstates, tree, x) + acc
    return acc


# ------------------------------------------------ +Gamma mixture (extension)
def pruning_mixture(t, nstates, tp_for_category, leaf_state, root_frequencies, weights):
    """Rate-category mixture on top of `pruning` (NOT in the reference).

    site lik = sum_k w_k exp(pruning with category k's matrices); evaluated as a
    log-sum-exp of the C independent pruning results.
    """
    ls = [pruning(t, nstates, tp_for_category(k), leaf_state, root_frequencies)
          for k in range(len(weights))]
    m = max(ls)
    if m == -math.inf:
        return m
    return m + math.log(sum(w * math.exp(l - m) for w, l in zip(weights, ls)))


# ------------------------------------------------------------------ conditional simulation
def conditional_simulation(rng, root, child_ptr, child_idx, P, clv, carry, root_frequencies, leaf_sets):
    """Phylo_ctmc.conditional_simulation (lib/phylo_ctmc.ml:122-143) with the Missing_values (:199-229)
    and Ambiguous (:331-360) leaf rules, on a flattened tree, for one site: top down, a node's state is
    drawn with weights prior_i * cl_i (prior = root frequencies, then the parent's row of P); a node whose
    carry is NaN (nothing observed below it) and a missing leaf are drawn from the prior, an ambiguous leaf
    from the prior restricted to its set.  `rng` is a numpy Generator (the reference uses
    Gsl.Randist.discrete: same distribution, another stream).  leaf_sets[v] = allowed states of leaf v
    (None = missing).  Returns the state of every node."""
    n = len(root_frequencies)
    states = np.full(len(child_ptr) - 1, -1, dtype=np.int64)

    def draw(w):
        w = np.asarray(w, float)
        return int(rng.choice(len(w), p=w / w.sum()))

    stack = [(root, np.asarray(root_frequencies, float))]
    while stack:
        v, prior = stack.pop()
        kids = child_idx[child_ptr[v]:child_ptr[v + 1]]
        if len(kids) == 0:
            allowed = leaf_sets[v]
            if allowed is None:
                states[v] = draw(prior)
            elif len(allowed) == 1:
                states[v] = allowed[0]
            else:
                w = np.zeros(n)
                w[list(allowed)] = prior[list(allowed)]
                states[v] = draw(w)
            continue
        states[v] = draw(prior if np.isnan(carry[v]) else prior * clv[v])
        for ch in kids:
            stack.append((ch, P[ch][states[v]]))
    return states


def conditional_simulation_exact(root, child_ptr, child_idx, P, root_frequencies, leaf_sets):
    """The distribution conditional_simulation draws from, by enumeration (small trees): the joint
    probability of every assignment of the unobserved nodes given the leaves,
    proportional to pi[root] * prod over branches of P[child][state(parent), state(child)].
    Returns (free_nodes, table) with table[assignment tuple] = probability."""
    import itertools
    n = len(root_frequencies)
    nn = len(child_ptr) - 1
    parent = np.full(nn, -1)
    for v in range(nn):
        for ch in child_idx[child_ptr[v]:child_ptr[v + 1]]:
            parent[ch] = v
    choices = []
    for v in range(nn):
        if child_ptr[v] == child_ptr[v + 1]:
            s = leaf_sets[v]
            choices.append(list(range(n)) if s is None else list(s))
        else:
            choices.append(list(range(n)))
    free = [v for v in range(nn) if len(choices[v]) > 1]
    table = {}
    for assign in itertools.product(*choices):
        p = root_frequencies[assign[root]]
        for v in range(nn):
            if v != root:
                p *= P[v][assign[parent[v]], assign[v]]
        key = tuple(assign[v] for v in free)
        table[key] = table.get(key, 0.0) + p
    z = sum(table.values())
    return free, {k: p / z for k, p in table.items()}
