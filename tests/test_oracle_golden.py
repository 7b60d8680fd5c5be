"""The oracle against the high-precision fixtures (tests/golden/known_answers.json, written by
oracle/make_golden.py with 60-digit mpmath) and against closed forms: this is what certifies
the oracle at the 1e-10 level, since the reference's own tests stop at 1e-5."""
import math

import numpy as np
import pytest

import golden_io
from oracle import c_oracle as C
from oracle import linalg_ref as la
from oracle import models_ref as M
from oracle import phylo_ctmc_ref as R

DOC = golden_io.load()


def pade_P(c):
    f = c["flat"]
    P = np.zeros((c["ncat"], f.nnodes, c["n"], c["n"]))
    for k, r in enumerate(c["rates"]):
        for v in range(f.nnodes):
            if v != f.root:
                P[k, v] = la.expm((c["bl"][v] * r) * c["Q"])
    return P


@pytest.mark.parametrize("entry", DOC["cases"], ids=[c["name"] for c in DOC["cases"]])
def test_c_oracle_matches_mpmath(entry):
    c = golden_io.case_arrays(entry)
    f = c["flat"]
    ps = C.pruning(c["n"], f.root, f.child_ptr, f.child_idx, f.tip_row, pade_P(c), c["pi"], tips=c["tips"],
                   weights=c["weights"])
    np.testing.assert_allclose(ps, c["site_loglik"], rtol=2e-13, atol=0)
    assert abs(C.sum_exact(ps) - c["total"]) <= 1e-13 * abs(c["total"])
    assert abs(C.sum_sequential(ps) - c["total"]) <= 1e-13 * abs(c["total"])


@pytest.mark.parametrize("entry", DOC["cases"][:3], ids=[c["name"] for c in DOC["cases"][:3]])
def test_python_oracle_matches_mpmath(entry):
    c = golden_io.case_arrays(entry)
    f = c["flat"]
    P = pade_P(c)
    tree = C.tree_from_csr(f.root, f.child_ptr, f.child_idx, f.tip_row)
    for s in range(c["tips"].shape[1]):
        got = R.pruning_mixture(tree, c["n"], lambda k: (lambda b: [("Mat", P[k, b])]),
                                lambda row: c["tips"][row, s], c["pi"], c["weights"])
        assert abs(got - c["site_loglik"][s]) <= 2e-13 * abs(c["site_loglik"][s])


def test_closed_form_jc69_pair():
    """Two leaves under JC69: ln(1/4 (1/4 +- ... e^{-4(t1+t2)/3}))."""
    pi = np.full(4, 0.25)
    red = M.jc69_reduction()
    t = R.binary_node(None, R.Branch(0.3, R.Leaf("a")), R.Branch(0.7, R.Leaf("b")))
    tp = lambda l: [("Mat", M.make_diag_tpm(red, l))]
    same = R.pruning(t, 4, tp, lambda l: 2, pi)
    diff = R.pruning(t, 4, tp, lambda l: 0 if l == "a" else 3, pi)
    assert abs(same - float(DOC["closed_forms"]["jc69_pair_same"])) < 1e-14
    assert abs(diff - float(DOC["closed_forms"]["jc69_pair_diff"])) < 1e-14


def test_reference_fixed_fixture_k80():
    """lib/mCMC.ml:22,36-38 and tests/test_MCMC.ml:11-12: the only RNG-free (tree, alignment,
    model) triples in the reference; values from the 60-digit restatement."""
    seqs = {"T0": 0, "T1": 0, "T2": 0, "T3": 3}
    red = M.k80_reduction(2.0)
    pi = np.full(4, 0.25)
    for x, pre in (("3.0", "0.1;0.1;0.1;0.1;0;1;3.0;0.1;2;3"), ("2.5", "0.1;0.1;0.1;0.1;0;1;2.5;0.1;2;3")):
        tree = R.of_preorder(pre)
        got = R.felsenstein(lambda l: M.make_diag_tpm(red, l), pi, lambda idx, s: seqs[idx], 4, tree, 1)
        assert abs(got - float(DOC["mcmc_fixture"][x])) < 1e-13
    # numbers derived independently during the survey
    assert abs(float(DOC["mcmc_fixture"]["3.0"]) - (-5.704499092002674)) < 1e-12
    assert abs(float(DOC["mcmc_fixture"]["2.5"]) - (-5.708825294632623)) < 1e-12


def test_rescaling_is_neutral_and_fires():
    """Long caterpillar: the shift rule must fire and must not change the value beyond rounding."""
    rng = np.random.default_rng(3)
    n, T = 4, 120
    Q = M.gtr(np.array([0.1, 0.2, 0.3, 0.4]), M.make_symetric(4, lambda i, j: 1.0 + 0.3 * i + 0.1 * j))
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    # caterpillar in CSR: internal nodes 0..T-2, node i has children (i+1 or last leaf, leaf)
    child_ptr, child_idx, tip_row = [0], [], []
    nnodes = 2 * T - 1
    for v in range(T - 1):
        inner = v + 1 if v < T - 2 else nnodes - 1
        child_idx += [inner, T - 1 + v]
        child_ptr.append(len(child_idx))
        tip_row.append(-1)
    for v in range(T - 1, nnodes):
        child_ptr.append(len(child_idx))
        tip_row.append(v - (T - 1))
    bl = rng.uniform(0.5, 2.0, nnodes)
    P = np.array([la.expm(b * Q) for b in bl])[None]
    tips = rng.integers(0, 4, size=(T, 7), dtype=np.uint8)
    with_shift = C.pruning(n, 0, child_ptr, child_idx, tip_row, P, pi, tips=tips)
    no_shift = C.pruning(n, 0, child_ptr, child_idx, tip_row, P, pi, tips=tips, threshold=-math.inf, root_mode=0)
    assert with_shift.min() < -150
    np.testing.assert_allclose(with_shift, no_shift, rtol=1e-13)
    felsenstein_mode = C.pruning(n, 0, child_ptr, child_idx, tip_row, P, pi, tips=tips, threshold=1e-7, root_mode=0)
    np.testing.assert_allclose(with_shift, felsenstein_mode, rtol=1e-13)


def test_impossible_pattern_is_not_finite():
    """All-zero vector: the reference computes 1/0 (lib/phylo_ctmc.ml:36-39); we keep non-finite output."""
    P = np.zeros((1, 3, 2, 2))
    P[0, 1] = P[0, 2] = np.eye(2)          # zero-length branches: states can never differ
    out = C.pruning(2, 0, [0, 2, 2, 2], [1, 2], [-1, 0, 1], P, np.array([0.5, 0.5]),
                    tips=np.array([[0], [1]], dtype=np.uint8))
    assert not np.isfinite(out[0])
