"""Host-side tree containers mirroring the reference's input types, and the
flattening that crosses the C ABI.

  Tree.t  (k-ary, node / leaf / branch payloads)      lib/tree.ml:3-13
  Phylogenetic_tree.t (binary, string leaf indices)   lib/phylogenetic_tree.ml:6-10

OCaml closures (`leaf_state`, `transition_probabilities`) cannot cross a C ABI,
so the wrapper walks the tree once, numbers the nodes in pre-order and emits the
CSR arrays pb_set_tree takes.
"""
import random as _random

import numpy as np


class Leaf:
    __slots__ = ("data",)

    def __init__(self, data):
        self.data = data


class Branch:
    __slots__ = ("data", "tip")

    def __init__(self, data, tip):
        self.data = data
        self.tip = tip


class Node:
    """Tree.Node {data; branches : branch List1.t} — at least one branch."""
    __slots__ = ("data", "branches")

    def __init__(self, data, branches):
        branches = list(branches)
        if not branches:
            raise ValueError("Tree.node: empty branch list")
        self.data = data
        self.branches = branches


def leaf(data):
    return Leaf(data)


def node(data, branches):
    return Node(data, branches)


def binary_node(data, left, right):
    return Node(data, [left, right])


def branch(data, tip):
    return Branch(data, tip)


def leaves(t):
    """Tree.leaves: leaf payloads, left to right."""
    out, stack = [], [t]
    while stack:
        x = stack.pop()
        if isinstance(x, Leaf):
            out.append(x.data)
        else:
            stack.extend(b.tip for b in reversed(x.branches))
    return out


class FlatTree:
    """CSR description of a rooted tree, nodes numbered in pre-order.

    child_ptr[nnodes+1], child_idx[nnodes-1] (reference branch order), tip_row[nnodes]
    (-1 for internal nodes), branch_data[v] = payload of the branch above node v
    (None for the root), leaf_data[row] / node_data[v] = payloads.
    """

    def __init__(self, nnodes, root, child_ptr, child_idx, tip_row, branch_data, leaf_data, node_data):
        self.nnodes = nnodes
        self.root = root
        self.child_ptr = np.asarray(child_ptr, dtype=np.int32)
        self.child_idx = np.asarray(child_idx, dtype=np.int32)
        self.tip_row = np.asarray(tip_row, dtype=np.int32)
        self.branch_data = branch_data
        self.leaf_data = leaf_data
        self.node_data = node_data

    @property
    def nleaves(self):
        return len(self.leaf_data)

    @property
    def ninternal(self):
        return self.nnodes - self.nleaves

    def is_leaf(self, v):
        return self.child_ptr[v] == self.child_ptr[v + 1]

    def children(self, v):
        return self.child_idx[self.child_ptr[v]:self.child_ptr[v + 1]]

    def branch_lengths(self, length=float):
        """[nnodes] array: length(branch_data[v]) for v != root, 0 for the root."""
        out = np.zeros(self.nnodes)
        for v in range(self.nnodes):
            if v != self.root:
                out[v] = length(self.branch_data[v])
        return out


def flatten(t):
    """Number the nodes of a Tree.t in pre-order and emit the CSR arrays (no recursion)."""
    if isinstance(t, Leaf):
        raise ValueError("the root must be an internal node")
    ids = {}
    order = []
    branch_data = []
    stack = [(t, None)]
    while stack:
        x, bdata = stack.pop()
        ids[id(x)] = len(order)
        order.append(x)
        branch_data.append(bdata)
        if isinstance(x, Node):
            for b in reversed(x.branches):
                stack.append((b.tip, b.data))
    nnodes = len(order)
    child_ptr = [0]
    child_idx = []
    tip_row = []
    leaf_data = []
    node_data = []
    for x in order:
        if isinstance(x, Leaf):
            tip_row.append(len(leaf_data))
            leaf_data.append(x.data)
            node_data.append(None)
        else:
            tip_row.append(-1)
            node_data.append(x.data)
            child_idx.extend(ids[id(b.tip)] for b in x.branches)
        child_ptr.append(len(child_idx))
    return FlatTree(nnodes, 0, child_ptr, child_idx, tip_row, branch_data, leaf_data, node_data)


def unflatten(flat, node=lambda v, data: data, leaf=lambda row, data: data, branch=lambda v, data: data):
    """Rebuild a Tree.t from a FlatTree, mapping payloads (used to return
    conditional-likelihood trees of the same shape)."""
    built = [None] * flat.nnodes
    for v in range(flat.nnodes - 1, -1, -1):           # pre-order ids: children have larger ids
        if flat.is_leaf(v):
            row = int(flat.tip_row[v])
            built[v] = Leaf(leaf(row, flat.leaf_data[row]))
        else:
            built[v] = Node(node(v, flat.node_data[v]),
                            [Branch(branch(int(c), flat.branch_data[int(c)]), built[int(c)])
                             for c in flat.children(v)])
    return built[flat.root]


# ------------------------------------------------------- Phylogenetic_tree.t
class PhyloLeaf:
    """Phylogenetic_tree.Leaf {index; meta}."""
    __slots__ = ("index", "meta")

    def __init__(self, index, meta=None):
        self.index = index
        self.meta = meta


class PhyloNode:
    """Phylogenetic_tree.Node {left = (length, t); right = (length, t); meta}."""
    __slots__ = ("left", "right", "meta")

    def __init__(self, left, right, meta=None):
        self.left = left
        self.right = right
        self.meta = meta


def of_preorder(s):
    """Phylogenetic_tree.of_preorder (lib/phylogenetic_tree.ml:72-97): ';'-separated
    tokens; a token containing '.' opens a node (two branch lengths, then both
    subtrees), an integer i is the leaf "T<i>".  Raises ValueError (invalid_arg)."""
    toks = s.split(";")
    pos = [0]

    def parse():
        if pos[0] >= len(toks):
            raise ValueError("Empty token list.")
        tok = toks[pos[0]]
        pos[0] += 1
        if "." in tok:
            f1 = float(tok)
            if pos[0] >= len(toks) or "." not in toks[pos[0]]:
                raise ValueError("Expected second float")
            f2 = float(toks[pos[0]])
            pos[0] += 1
            left = parse()
            right = parse()
            return PhyloNode((f1, left), (f2, right))
        return PhyloLeaf("T%d" % int(tok))
    return parse()


def make_random(n, rng=None):
    """Phylogenetic_tree.make_random (lib/phylogenetic_tree.ml:99-110): repeatedly join
    two random subtrees, branch lengths U(0, 0.5) + 1e-5.  (Own RNG stream: the
    reference uses OCaml's global Random.)"""
    rng = rng or _random.Random(0)
    forest = [PhyloLeaf("T%d" % i) for i in range(n)]
    if not forest:
        raise RuntimeError("tree list should not be empty")
    while len(forest) > 1:
        rng.shuffle(forest)
        a, b = forest[0], forest[1]
        joined = PhyloNode((rng.random() * 0.5 + 0.00001, a), (rng.random() * 0.5 + 0.00001, b))
        forest = [joined] + forest[2:]
    return forest[0]


def to_tree(t):
    """Phylogenetic_tree.to_tree (lib/phylogenetic_tree.ml:137-144): leaves carry
    (meta, index), branches carry their length."""
    done = {}
    stack = [(t, False)]
    while stack:
        x, expanded = stack.pop()
        if isinstance(x, PhyloLeaf):
            done[id(x)] = Leaf((x.meta, x.index))
        elif expanded:
            done[id(x)] = binary_node(x.meta, Branch(x.left[0], done[id(x.left[1])]),
                                      Branch(x.right[0], done[id(x.right[1])]))
        else:
            stack.append((x, True))
            stack.append((x.right[1], False))
            stack.append((x.left[1], False))
    return done[id(t)]


def get_branch_lengths(t):
    """Phylogenetic_tree.get_branch_lengths (lib/phylogenetic_tree.ml:163-165):
    l1 :: l2 :: lengths(left) @ lengths(right)."""
    if isinstance(t, PhyloLeaf):
        return []
    return [t.left[0], t.right[0]] + get_branch_lengths(t.left[1]) + get_branch_lengths(t.right[1])


def nb_branches(t):
    if isinstance(t, PhyloLeaf):
        return 0
    return 2 + nb_branches(t.left[1]) + nb_branches(t.right[1])


def set_branch_lengths(t, lengths):
    """Phylogenetic_tree.set_branch_lengths (lib/phylogenetic_tree.ml:153-161)."""
    lengths = list(lengths)
    if isinstance(t, PhyloLeaf):
        if lengths:
            raise RuntimeError("Branch list does not match tree")
        return t
    if len(lengths) < 2:
        raise RuntimeError("Branch list does not match tree")
    l1, l2, tl = lengths[0], lengths[1], lengths[2:]
    nl, nr = nb_branches(t.left[1]), nb_branches(t.right[1])
    return PhyloNode((l1, set_branch_lengths(t.left[1], tl[:nl])),
                     (l2, set_branch_lengths(t.right[1], tl[nl:nl + nr])), t.meta)
