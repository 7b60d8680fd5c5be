"""The host compiler of the 4-state site-resident kernel, checked on the CPU.

`pb_set_tree` compiles a tree into an accumulator/stack program (pb_api.cu: compile_tree) and, per evaluation
mode, into a variant in which small clades are table look-ups (build_table_program).  `pb_debug_fused_program`
returns both without touching a device; this test INTERPRETS them in numpy — tips consumed from packed 2-bit
windows exactly as the kernel does, no rescaling (small trees: nothing underflows) — and compares the per-site
log-likelihoods with the oracle.  That covers the tip order and word boundaries, the stack discipline, the
compact matrix indices, and the clade finder (every table entry is produced by interpreting the clade's own
instruction range on a synthetic window, which is what build_clade_tables does on the device).
"""
import ctypes

import numpy as np
import pytest

import cases
from oracle import c_oracle as C
from phylogenetics_b200 import _lib, synthetic, tree as T

TIP_TIP, TIP, APPLY, MUL_TIP, PUSH, POP_MUL, APPLY_MUL_TIP, ROOT, LOADW, ACC_POP, CHERRY = range(11)
PUSH_AFTER, PRE_A, PRE_B = 0x100, 0x200, 0x400


def compile_program(flat, ntaxa, max_leaves, preapply=False):
    nn = flat.nnodes
    cap = 8 * nn + 64
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    child_ptr, child_idx, tip_row = i32(flat.child_ptr), i32(flat.child_idx), i32(flat.tip_row)
    counts = np.zeros(8, np.int32)
    ops, ops_tab = np.zeros((cap, 4), np.int32), np.zeros((cap, 4), np.int32)
    order, inner, leaf = (np.zeros(nn, np.int32) for _ in range(3))
    clades = np.zeros((nn, 6), np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = _lib.lib.pb_debug_fused_program(nn, int(flat.root), p(child_ptr), p(child_idx), p(tip_row), ntaxa,
                                         max_leaves | (0x100 if preapply else 0),
                                         p(counts), p(ops), p(ops_tab), cap, p(order), p(inner), p(leaf), p(clades), nn)
    assert rc == 0, _lib.lib.pb_last_error(None)
    nops, nops_tab, ntips, nwords, stack, ninner, nleaf, nclades = (int(x) for x in counts)
    return dict(ops=ops[:nops], ops_tab=ops_tab[:nops_tab], order=order[:ntips], nwords=nwords, stack=stack,
                inner=inner[:ninner], leaf=leaf[:nleaf], clades=clades[:nclades])


def windows_of(prog, tips):
    """[nwords][nsites] uint64: tip k of the program's order in bits 2(k % 32).. of word k // 32 (pack_tips4)."""
    nsites = tips.shape[1]
    w = np.zeros((max(prog["nwords"], 1), nsites), dtype=np.uint64)
    for k, row in enumerate(prog["order"]):
        w[k // 32] |= tips[row].astype(np.uint64) << np.uint64(2 * (k % 32))
    return w


def run(ops, prog, P, pi, windows, table=None, max_stack=None):
    """Interpret `ops` for all sites at once.  P: [nnodes][4][4] row-stochastic (i = parent state).  Returns the
    per-site likelihood at FOP_ROOT, or (no ROOT in the range) the accumulator [nsites][4]."""
    nsites = windows.shape[1]
    word, cur = 0, windows[0]
    acc, stack, depth = np.ones((nsites, 4)), [], 0
    inner = lambda off: P[prog["inner"][off // 16]]
    col = lambda off, sh: P[prog["leaf"][off // 16]][:, ((cur >> np.uint64(sh)) & np.uint64(3)).astype(int)].T   # [nsites][4]
    for code_w, offA, offB, sh in ops:
        code, shA, shB = code_w & 0xff, sh & 0xff, sh >> 8
        if code == LOADW:
            word += 1
            cur = windows[word]
        elif code == TIP_TIP:
            acc = col(offA, shA) * col(offB, shB)
        elif code == TIP:
            acc = col(offA, shA)
        elif code == APPLY:
            acc = acc @ inner(offA).T
        elif code == MUL_TIP:
            acc = acc * col(offA, shA)
        elif code == POP_MUL:
            acc = stack.pop() * acc
        elif code == APPLY_MUL_TIP:
            acc = (acc if code_w & PRE_A else acc @ inner(offA).T) * col(offB, shB)
        elif code == ACC_POP:
            x = stack.pop()
            acc = (acc if code_w & PRE_A else acc @ inner(offA).T) * (x if code_w & PRE_B else x @ inner(offB).T)
        elif code == CHERRY:
            acc = table(offA, ((cur >> np.uint64(shA)) & np.uint64(offB)).astype(int))
        elif code == ROOT:
            assert not stack
            return acc @ pi
        else:
            assert code == PUSH, code
        if code_w & PUSH_AFTER:
            stack.append(acc)
            depth = max(depth, len(stack))
            if max_stack is not None:
                assert depth <= max_stack
    return acc


def clade_table(prog, P, pi):
    """entry -> vector, by interpreting each clade's instruction range on the synthetic window idx << sh0."""
    tabs = {}
    for op0, op1, sh0, k, ent0, pre in prog["clades"]:
        idx = np.arange(4 ** k, dtype=np.uint64)
        rng_ops = prog["ops"][op0:op1 + 1].copy()
        rng_ops[-1, 0] &= ~PUSH_AFTER                        # the last instruction's push stays in the program
        assert not (rng_ops[:, 0] & 0xff == LOADW).any()     # a clade lives in one window word
        v = run(rng_ops, prog, P, pi, (idx << np.uint64(sh0))[None, :])
        tabs[int(ent0)] = v if pre < 0 else v @ P[prog["inner"][pre // 16]].T      # the consumer's branch matrix
    return lambda ent0, i: tabs[int(ent0)][i]


def random_kary_tree(rng, nleaves):
    """Arities 1..4 mixed (lib/tree.ml allows any arity >= 1)."""
    items = [T.leaf(i) for i in range(nleaves)]
    while len(items) > 1:
        k = min(len(items), int(rng.integers(2, 5)))
        picks = sorted(rng.choice(len(items), size=k, replace=False), reverse=True)
        kids = [items.pop(int(i)) for i in picks]
        node = T.node(None, [T.branch(float(rng.uniform(0.02, 0.6)), kid) for kid in kids])
        if rng.random() < 0.15:
            node = T.node(None, [T.branch(float(rng.uniform(0.02, 0.6)), node)])      # a unary node
        items.append(node)
    return items[0]


def check(flat, P, pi, tips, max_leaves, preapply=False):
    prog = compile_program(flat, tips.shape[0], max_leaves, preapply)
    assert sorted(prog["order"]) == list(range(tips.shape[0]))          # every tip consumed exactly once
    assert prog["nwords"] == (tips.shape[0] + 31) // 32
    win = windows_of(prog, tips)
    expect = C.pruning(4, flat.root, flat.child_ptr, flat.child_idx, flat.tip_row, P[None], pi, tips=tips)
    plain = np.log(run(prog["ops"], prog, P, pi, win, max_stack=prog["stack"]))
    np.testing.assert_allclose(plain, expect, rtol=1e-11)
    tab = np.log(run(prog["ops_tab"], prog, P, pi, win, table=clade_table(prog, P, pi), max_stack=prog["stack"]))
    np.testing.assert_allclose(tab, expect, rtol=1e-11)
    ks = [int(k) for k in prog["clades"][:, 3]]
    assert all(2 <= k <= max(max_leaves, 2) for k in ks) if max_leaves >= 2 else not ks
    # look-ups replace instructions: the tabulated program is shorter by the clades' instructions minus one each
    saved = sum(int(op1 - op0) for op0, op1, _, _, _, _ in prog["clades"])
    assert len(prog["ops_tab"]) == len(prog["ops"]) - saved
    # pre-applied matrices: one flag in the program per table that carries its consumer's matrix, none otherwise
    flags = sum(int(bool(w & PRE_A)) + int(bool(w & PRE_B)) for w in prog["ops_tab"][:, 0])
    assert flags == int((prog["clades"][:, 5] >= 0).sum()) and (preapply or flags == 0)
    return prog


@pytest.mark.parametrize("ntaxa,max_leaves", [(5, 5), (33, 2), (64, 3), (97, 4), (128, 5), (70, 6), (40, 0)])
def test_binary_trees(ntaxa, max_leaves):
    case = cases.make_case("prog", "gtr", ntaxa, 300, ncat=1, seed=ntaxa)
    P = case.host_P()[0]
    prog = check(case.flat, P, case.pi, case.tips, max_leaves)
    check(case.flat, P, case.pi, case.tips, max_leaves, preapply=True)
    if max_leaves >= 2 and ntaxa >= 33:
        assert len(prog["clades"]) >= ntaxa // 8          # a birth-death tree has about n/3 cherries


@pytest.mark.parametrize("seed", range(6))
def test_kary_and_unary_trees(seed):
    rng = np.random.default_rng(100 + seed)
    nleaves = int(rng.integers(6, 75))
    flat = T.flatten(random_kary_tree(rng, nleaves))
    Q, pi = synthetic.random_gtr(rng, 10.0)
    bl = flat.branch_lengths(float)
    P = synthetic.host_transition_matrices(Q, pi, bl, np.ones(1))
    tips = synthetic.simulate_alignment(rng, flat, P, pi, 200, np.ones(1))
    k = int(rng.integers(2, 7))
    check(flat, P[0], pi, tips, k)
    check(flat, P[0], pi, tips, k, preapply=True)


def test_caterpillar_crosses_word_boundaries():
    """A 100-leaf caterpillar consumes its tips in one long chain: clades stop at the 32-tip word boundaries."""
    t = T.leaf(0)
    for i in range(1, 100):
        t = T.binary_node(None, T.branch(0.05 + 0.001 * i, t), T.branch(0.07, T.leaf(i)))
    flat = T.flatten(t)
    rng = np.random.default_rng(9)
    Q, pi = synthetic.random_gtr(rng, 10.0)
    P = synthetic.host_transition_matrices(Q, pi, flat.branch_lengths(float), np.ones(1))
    tips = synthetic.simulate_alignment(rng, flat, P, pi, 150, np.ones(1))
    prog = check(flat, P[0], pi, tips, 6)
    assert len(prog["clades"]) == 1 and int(prog["clades"][0][3]) == 6      # only the innermost six leaves form one
    assert (prog["ops"][:, 0] & 0xff == LOADW).sum() == 3


def test_bad_arguments_are_reported():
    z = np.zeros(8, np.int32)
    p = z.ctypes.data_as(ctypes.c_void_p)
    assert _lib.lib.pb_debug_fused_program(1, 0, p, p, p, 1, 2, p, p, p, 8, p, p, p, p, 8) == _lib.PB_ERR_INVALID_ARG
    assert b"bad arguments" in _lib.lib.pb_last_error(None)


def test_random_shapes_property():
    """hypothesis: random arities, sizes, clade limits; the interpreted programs equal the oracle every time."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(st.integers(0, 10 ** 6), st.integers(3, 140), st.integers(0, 6), st.booleans(), st.booleans())
    def prop(seed, nleaves, max_leaves, kary, preapply):
        rng = np.random.default_rng(seed)
        if kary:
            t = random_kary_tree(rng, nleaves)
        else:
            t = synthetic.random_binary_tree(rng, nleaves)
        flat = T.flatten(t)
        Q, pi = synthetic.random_gtr(rng, 10.0)
        P = synthetic.host_transition_matrices(Q, pi, flat.branch_lengths(float), np.ones(1))
        tips = synthetic.simulate_alignment(rng, flat, P, pi, 64, np.ones(1))
        check(flat, P[0], pi, tips, max_leaves, preapply)
    prop()


def test_preapplied_matrices_at_the_bench_shape():
    """The c2 bench tree with clades of up to 5 leaves: 32 look-ups, and with their consumers' matrices folded into
    the tables 32 of the 76 remaining matrix-vector products disappear."""
    case = cases.make_case("c2", "gtr", 128, 64, ncat=4, seed=100)
    prog = check(case.flat, case.host_P()[0], case.pi, case.tips, 5, preapply=True)
    codes = prog["ops_tab"][:, 0]
    matvecs = sum(2 - bool(w & PRE_A) - bool(w & PRE_B) for w in codes if w & 0xff == ACC_POP) + \
        sum(1 - bool(w & PRE_A) for w in codes if w & 0xff == APPLY_MUL_TIP)
    assert (codes & 0xff == CHERRY).sum() == 32 and matvecs == 76 - 32
