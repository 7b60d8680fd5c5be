"""The planner of the default 4-state kernel (fused4t), checked on the CPU.

`build_reduced_program` (pb_api.cu) turns every maximal clade of at most k leaves into a table and compiles the
program over the REDUCED tree, whose leaves are those tables and the loose tips; a table value is an operand of its
parent's instruction.  `pb_debug_reduced_program` returns the plan without touching a device; this test INTERPRETS it
in numpy — fields taken from packed windows exactly as the kernel takes them, tables filled by interpreting each
clade's own instruction range (what build_clade_tables does on the device) — and compares the per-site
log-likelihoods with the oracle.  Covers: clade selection, operand kinds of every instruction, field layout and
word boundaries (no field straddles a word), stack discipline on the reduced tree, compact matrix indices, the
branch matrix folded into each table.
"""
import ctypes

import numpy as np
import pytest

import cases
from oracle import c_oracle as C
from phylogenetics_b200 import _lib, synthetic, tree as T
from test_fused_program_cpu import random_kary_tree

(TIP_TIP, TIP, APPLY, MUL_TIP, PUSH, POP_MUL, APPLY_MUL_TIP, ROOT, LOADW, ACC_POP, CHERRY,
 TAB, TAB_TAB, TAB_TIP, APPLY_MUL_TAB, MUL_TAB) = range(16)
PUSH_AFTER = 0x100


def plan(flat, ntaxa, max_leaves):
    nn = flat.nnodes
    cap = 8 * nn + 64
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    child_ptr, child_idx, tip_row = i32(flat.child_ptr), i32(flat.child_idx), i32(flat.tip_row)
    counts = np.zeros(8, np.int32)
    ops, cops = np.zeros((cap, 4), np.int32), np.zeros((cap, 4), np.int32)
    slots = np.zeros(32 * (nn + 2), np.int32)
    xslots = np.zeros_like(slots)
    mats, leaves = np.zeros(nn, np.int32), np.zeros(nn, np.int32)
    clades = np.zeros((nn, 6), np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = _lib.lib.pb_debug_reduced_program(nn, int(flat.root), p(child_ptr), p(child_idx), p(tip_row), ntaxa, max_leaves,
                                           p(counts), p(ops), cap, p(cops), cap, p(slots), p(xslots), len(slots), p(mats), p(leaves),
                                           p(clades), nn)
    assert rc == 0, _lib.lib.pb_last_error(None)
    nops, ncops, nslots, nwords, stack, nmat, nleaf, nclades = (int(x) for x in counts)
    return dict(ops=ops[:nops], cops=cops[:ncops], slots=slots[:nslots], xslots=xslots[:nslots], nwords=nwords, stack=stack,
                mats=mats[:nmat], leaves=leaves[:nleaf], clades=clades[:nclades])


def windows_of(pl, tips):
    """[nwords][nsites] uint64: slot s of the plan in bits 2(s % 32).. of word s // 32 (pack_tips4); padding slots are 0."""
    w = np.zeros((pl["nwords"], tips.shape[1]), dtype=np.uint64)
    for s, (row, xrow) in enumerate(zip(pl["slots"], pl["xslots"])):
        if row >= 0:
            st = tips[row] ^ tips[xrow] if xrow >= 0 else tips[row]     # difference coding inside a clade's field
            w[s // 32] |= st.astype(np.uint64) << np.uint64(2 * (s % 32))
    return w


def clade_tables(pl, P):
    """first entry -> [4^k][4]: the clade's own instruction range on the synthetic window idx, then the branch matrix."""
    tabs = {}
    for op0, op1, xor_coded, k, ent0, pre in pl["clades"]:
        cur = np.arange(4 ** k, dtype=np.uint64)
        if xor_coded:               # entry index = tip 0, then the other tips XOR tip 0
            rep = (cur & np.uint64(3)) * np.uint64(0x5555555555555555)
            cur = cur ^ (rep & ~np.uint64(3) & np.uint64(4 ** k - 1))
        col = lambda off, sh: P[off // 16][:, ((cur >> np.uint64(sh)) & np.uint64(3)).astype(int)].T
        acc, stack = np.ones((len(cur), 4)), []
        for code_w, offA, offB, sh in pl["cops"][op0:op1 + 1]:
            code, shA, shB = code_w & 0xff, sh & 0xff, (sh >> 8) & 0xff
            if code == TIP_TIP:
                acc = col(offA, shA) * col(offB, shB)
            elif code == TIP:
                acc = col(offA, shA)
            elif code == APPLY:
                acc = acc @ P[offA // 16].T
            elif code == MUL_TIP:
                acc = acc * col(offA, shA)
            elif code == POP_MUL:
                acc = stack.pop() * acc
            elif code == APPLY_MUL_TIP:
                acc = (acc @ P[offA // 16].T) * col(offB, shB)
            elif code == ACC_POP:
                acc = (acc @ P[offA // 16].T) * (stack.pop() @ P[offB // 16].T)
            else:
                assert code == PUSH, code
            if code_w & PUSH_AFTER:
                stack.append(acc)
        assert not stack
        tabs[int(ent0)] = acc if pre < 0 else acc @ P[pre // 16].T
    return tabs


def run(pl, P, pi, windows, tabs):
    word, cur = 0, windows[0]
    acc, stack, depth = np.ones((windows.shape[1], 4)), [], 0
    mat = lambda off: P[pl["mats"][off // 16]]
    col = lambda off, sh: P[pl["leaves"][off // 16]][:, ((cur >> np.uint64(sh)) & np.uint64(3)).astype(int)].T
    tab = lambda ent0, sh, bits: tabs[int(ent0)][((cur >> np.uint64(sh)) & np.uint64((1 << bits) - 1)).astype(int)]
    for code_w, offA, offB, sh in pl["ops"]:
        code = code_w & 0xff
        shA, shB, bitsA, bitsB = sh & 0xff, (sh >> 8) & 0xff, (sh >> 16) & 0xff, (sh >> 24) & 0xff
        assert shA + bitsA <= 64 and shB + bitsB <= 64
        if code == LOADW:
            word += 1
            cur = windows[word]
        elif code == TAB:
            acc = tab(offA, shA, bitsA)
        elif code == TAB_TAB:
            acc = tab(offA, shA, bitsA) * tab(offB, shB, bitsB)
        elif code == TAB_TIP:
            acc = tab(offA, shA, bitsA) * col(offB, shB)
        elif code == APPLY_MUL_TAB:
            acc = (acc @ mat(offA).T) * tab(offB, shB, bitsB)
        elif code == MUL_TAB:
            acc = acc * tab(offA, shA, bitsA)
        elif code == TIP_TIP:
            acc = col(offA, shA) * col(offB, shB)
        elif code == TIP:
            acc = col(offA, shA)
        elif code == APPLY:
            acc = acc @ mat(offA).T
        elif code == MUL_TIP:
            acc = acc * col(offA, shA)
        elif code == POP_MUL:
            acc = stack.pop() * acc
        elif code == APPLY_MUL_TIP:
            acc = (acc @ mat(offA).T) * col(offB, shB)
        elif code == ACC_POP:
            acc = (acc @ mat(offA).T) * (stack.pop() @ mat(offB).T)
        elif code == ROOT:
            assert not stack
            return acc @ pi
        else:
            assert code == PUSH, code
        if code_w & PUSH_AFTER:
            stack.append(acc)
            depth = max(depth, len(stack))
            assert depth <= pl["stack"]
    raise AssertionError("no ROOT instruction")


def check(flat, P, pi, tips, max_leaves):
    pl = plan(flat, tips.shape[0], max_leaves)
    rows = [int(r) for r in pl["slots"] if r >= 0]
    assert sorted(rows) == list(range(tips.shape[0]))                    # every tip sits in exactly one slot
    assert len(pl["slots"]) == 32 * pl["nwords"]
    expect = C.pruning(4, flat.root, flat.child_ptr, flat.child_idx, flat.tip_row, P[None], pi, tips=tips)
    got = np.log(run(pl, P, pi, windows_of(pl, tips), clade_tables(pl, P)))
    np.testing.assert_allclose(got, expect, rtol=1e-11)
    ks = [int(k) for k in pl["clades"][:, 3]]
    assert all(1 <= k <= max(max_leaves, 1) for k in ks) if max_leaves >= 2 else not ks
    codes = pl["ops"][:, 0] & 0xff
    assert not (codes == CHERRY).any()
    if max_leaves < 2:
        assert not np.isin(codes, [TAB, TAB_TAB, TAB_TIP, APPLY_MUL_TAB, MUL_TAB]).any()
    return pl


@pytest.mark.parametrize("ntaxa,max_leaves", [(5, 5), (33, 2), (64, 3), (97, 4), (128, 5), (70, 6), (40, 0), (3, 6)])
def test_binary_trees(ntaxa, max_leaves):
    case = cases.make_case("rprog", "gtr", ntaxa, 300, ncat=1, seed=ntaxa)
    pl = check(case.flat, case.host_P()[0], case.pi, case.tips, max_leaves)
    if max_leaves >= 2 and ntaxa >= 33:
        assert len(pl["clades"]) >= ntaxa // 8


@pytest.mark.parametrize("seed", range(6))
def test_kary_and_unary_trees(seed):
    rng = np.random.default_rng(100 + seed)
    nleaves = int(rng.integers(6, 75))
    flat = T.flatten(random_kary_tree(rng, nleaves))
    Q, pi = synthetic.random_gtr(rng, 10.0)
    P = synthetic.host_transition_matrices(Q, pi, flat.branch_lengths(float), np.ones(1))
    tips = synthetic.simulate_alignment(rng, flat, P, pi, 200, np.ones(1))
    check(flat, P[0], pi, tips, int(rng.integers(2, 7)))


def test_caterpillar():
    """A 100-leaf caterpillar: one clade at the bottom, then a chain of APPLY_MUL_TIP; no stack at all."""
    t = T.leaf(0)
    for i in range(1, 100):
        t = T.binary_node(None, T.branch(0.05 + 0.001 * i, t), T.branch(0.07, T.leaf(i)))
    flat = T.flatten(t)
    rng = np.random.default_rng(9)
    Q, pi = synthetic.random_gtr(rng, 10.0)
    P = synthetic.host_transition_matrices(Q, pi, flat.branch_lengths(float), np.ones(1))
    tips = synthetic.simulate_alignment(rng, flat, P, pi, 150, np.ones(1))
    pl = check(flat, P[0], pi, tips, 6)
    assert len(pl["clades"]) == 1 and int(pl["clades"][0][3]) == 6
    assert not (pl["ops"][:, 0] & PUSH_AFTER).any()


def test_bench_shape_pushes_and_matrices():
    """c2 (128 taxa), clades of up to 5 leaves: the reduced program pushes only where both children are computed
    and applies far fewer matrices than the 126 inner branches of the tree."""
    case = cases.make_case("c2", "gtr", 128, 64, ncat=4, seed=100)
    pl = check(case.flat, case.host_P()[0], case.pi, case.tips, 5)
    codes = pl["ops"][:, 0]
    pushes = int(((codes & PUSH_AFTER) != 0).sum())
    assert pushes == int(((codes & 0xff) == ACC_POP).sum()) + int(((codes & 0xff) == POP_MUL).sum())
    assert pushes <= 16 and len(pl["mats"]) <= 64 and pl["stack"] <= 3
    assert len(pl["leaves"]) + sum(int(k) for k in pl["clades"][:, 3]) == 128


def test_random_shapes_property():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(st.integers(0, 10 ** 6), st.integers(3, 140), st.integers(0, 6), st.booleans())
    def prop(seed, nleaves, max_leaves, kary):
        rng = np.random.default_rng(seed)
        t = random_kary_tree(rng, nleaves) if kary else synthetic.random_binary_tree(rng, nleaves)
        flat = T.flatten(t)
        Q, pi = synthetic.random_gtr(rng, 10.0)
        P = synthetic.host_transition_matrices(Q, pi, flat.branch_lengths(float), np.ones(1))
        tips = synthetic.simulate_alignment(rng, flat, P, pi, 64, np.ones(1))
        check(flat, P[0], pi, tips, max_leaves)
    prop()
