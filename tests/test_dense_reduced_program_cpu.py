"""The planner of the dense-alphabet kernel's reduced program (fused_dmma3 with tabulated cherries), on the CPU.

`build_dense_reduced_program` (pb_api.cu) makes every node with two leaf children a table of nstates^2 entries that
carries the matrix of the branch above it, and compiles the program over the tree whose leaves are those tables and
the loose tips.  `pb_debug_dense_reduced_program` returns the plan; this test interprets it in numpy against the
oracle and checks what the kernel's matrix ring relies on: the list of matrices is exactly the sequence of
matrix-vector products the instructions perform, in order (a mismatch would dead-lock the ring)."""
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


def plan(flat, ntaxa, nstates, triple=False):
    nn = flat.nnodes
    cap = 4 * nn + 16
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    child_ptr, child_idx, tip_row = i32(flat.child_ptr), i32(flat.child_idx), i32(flat.tip_row)
    counts = np.zeros(8, np.int32)
    ops = np.zeros((cap, 4), np.int32)
    mlist = np.zeros(cap, np.int32)
    tabs, tab_nodes = np.zeros((nn, 4), np.int32), np.zeros((nn, 8), np.int32)
    inner, leaf = np.zeros(nn, np.int32), np.zeros(nn, np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = _lib.lib.pb_debug_dense_reduced_program(nn, int(flat.root), p(child_ptr), p(child_idx), p(tip_row), ntaxa, nstates | (0x100 if triple else 0),
                                                 p(counts), p(ops), cap, p(mlist), cap, p(tabs), p(tab_nodes), nn, p(inner), p(leaf))
    assert rc == 0, _lib.lib.pb_last_error(None)
    nops, nmat, ntab, stack, ninner, nleaf, stack_full, nmat_full = (int(x) for x in counts)
    return dict(ops=ops[:nops], mlist=mlist[:nmat], tabs=tabs[:ntab], tab_nodes=tab_nodes[:ntab], stack=stack,
                inner=inner[:ninner], leaf=leaf[:nleaf], stack_full=stack_full, nmat_full=nmat_full)


def run(pl, P, pi, tips, n):
    """Interpret the plan for all sites; returns the per-site likelihood and the matrices applied, in order."""
    nsites = tips.shape[1]
    tables, nent = [], 0
    for (ent0, rowA, rowB, rowC), (la, lb, cherry, lc, top, is_root, _, _) in zip(pl["tabs"], pl["tab_nodes"]):
        assert ent0 == nent
        cols = P[la][:, :, None] * P[lb][:, None, :]                       # [state][a][b]: clv of the cherry
        if lc >= 0:                                                       # a leaf joins the cherry: [state][a][b][c]
            assert n ** 3 <= 8000 and rowC >= 0 and top != cherry
            inner = np.einsum("ij,jab->iab", P[cherry], cols)
            cols = inner[:, :, :, None] * P[lc][:, None, None, :]
        else:
            assert rowC < 0 and top == cherry
        flat_cols = cols.reshape(n, -1)                                   # entry = (a * n + b) [* n + c]
        t = (flat_cols if is_root else P[top] @ flat_cols).T.copy()       # [entry][state]
        nent += t.shape[0]
        tables.append((t, rowA, rowB, rowC))

    def tab(k):
        t, rowA, rowB, rowC = tables[k]
        idx = tips[rowA].astype(int) * n + tips[rowB].astype(int)
        if rowC >= 0:
            idx = idx * n + tips[rowC].astype(int)
        return t[idx]                                                     # [nsites][n]
    col = lambda leaf_idx, row: P[pl["leaf"][leaf_idx]][:, tips[row].astype(int)].T
    applied = []

    def apply(idx, x):
        applied.append(int(idx))
        return x @ P[pl["inner"][idx]].T
    acc, stack, depth = np.ones((nsites, n)), [], 0
    for code_w, A, B, sh in pl["ops"]:
        code, rowA, rowB = code_w & 0xff, sh & 0xffff, (sh >> 16) & 0xffff
        if code == TAB:
            acc = tab(A)
        elif code == TAB_TAB:
            acc = tab(A) * tab(B)
        elif code == TAB_TIP:
            acc = tab(A) * col(B, rowB)
        elif code == APPLY_MUL_TAB:
            acc = apply(A, acc) * tab(B)
        elif code == MUL_TAB:
            acc = acc * tab(A)
        elif code == TIP_TIP:
            acc = col(A, rowA) * col(B, rowB)
        elif code == TIP:
            acc = col(A, rowA)
        elif code == APPLY:
            acc = apply(A, acc)
        elif code == MUL_TIP:
            acc = acc * col(A, rowA)
        elif code == POP_MUL:
            acc = stack.pop() * acc
        elif code == APPLY_MUL_TIP:
            acc = apply(A, acc) * col(B, rowB)
        elif code == ACC_POP:
            y = apply(A, acc)                       # the kernel applies the accumulator's matrix first, then the popped one's
            acc = y * apply(B, stack.pop())
        elif code == ROOT:
            assert not stack
            return acc @ pi, applied
        else:
            assert code == PUSH, code
        if code_w & PUSH_AFTER:
            stack.append(acc)
            depth = max(depth, len(stack))
            assert depth <= pl["stack"]
    raise AssertionError("no ROOT instruction")


def check(flat, P, pi, tips, n):
    """Both plans: cherries only (the default) and with three-leaf clades (option fused_dmma_triple, 20 states)."""
    pl = check_plan(flat, P, pi, tips, n, False)
    assert int((pl["tab_nodes"][:, 3] >= 0).sum()) == 0
    pl3 = check_plan(flat, P, pi, tips, n, True)
    assert n ** 3 <= 8000 or int((pl3["tab_nodes"][:, 3] >= 0).sum()) == 0
    return pl


def check_plan(flat, P, pi, tips, n, triple):
    pl = plan(flat, tips.shape[0], n, triple)
    lik, applied = run(pl, P, pi, tips, n)
    assert applied == [int(m) for m in pl["mlist"]]          # the ring delivers exactly these matrices, in this order
    expect = C.pruning(n, flat.root, flat.child_ptr, flat.child_idx, flat.tip_row, P[None], pi, tips=tips)
    np.testing.assert_allclose(np.log(lik), expect, rtol=1e-10)
    ncherry = sum(1 for v in range(flat.nnodes)
                  if flat.child_ptr[v + 1] - flat.child_ptr[v] == 2 and
                  all(flat.child_ptr[c] == flat.child_ptr[c + 1] for c in flat.child_idx[flat.child_ptr[v]:flat.child_ptr[v + 1]]))
    assert len(pl["tabs"]) == ncherry                         # every cherry is in exactly one table (alone or with a third leaf)
    triples = int((pl["tab_nodes"][:, 3] >= 0).sum())
    assert triples == 0 or n ** 3 <= 8000                      # three-leaf clades only where N^3 entries are affordable (20 states)
    # matrix-vector products saved: one per table that is not the root (its own branch), one more per three-leaf clade (the cherry's)
    saved = sum(1 for row in pl["tab_nodes"] if not row[5]) + triples
    assert len(pl["mlist"]) == pl["nmat_full"] - saved
    assert pl["stack"] <= pl["stack_full"]
    return pl


@pytest.mark.parametrize("model,ntaxa", [("wag", 64), ("wag", 7), ("wag", 2), ("mg94", 32), ("mg94", 3), ("mutsel", 20)])
def test_binary_trees(model, ntaxa):
    case = cases.make_case("dred", model, ntaxa, 60, seed=ntaxa)
    pl = check(case.flat, case.host_P()[0], case.pi, case.tips, case.n)
    if ntaxa >= 20:
        assert len(pl["tabs"]) >= ntaxa // 5              # a birth-death tree has about n/3 cherries


@pytest.mark.parametrize("seed", range(4))
def test_kary_and_unary_trees(seed):
    rng = np.random.default_rng(400 + seed)
    flat = T.flatten(random_kary_tree(rng, int(rng.integers(5, 50))))
    w = cases.wag()
    pi = w.freqs / w.freqs.sum()
    from phylogenetics_b200 import models
    Q = models.Rate_matrix.scaled_rate_matrix(pi, models.Rate_matrix.make(20, lambda i, j: w.rate_matrix[i, j] * pi[j]))
    P = synthetic.host_transition_matrices(Q, pi, flat.branch_lengths(float), np.ones(1))
    tips = synthetic.simulate_alignment(rng, flat, P, pi, 40, np.ones(1))
    check(flat, P[0], pi, tips, 20)
