"""The reference's own tests for this path, replayed through the host mirror of its API on the GPU
(names and argument meaning as in lib/phylo_ctmc.mli, lib/felsenstein.mli), with the oracle as
the expected value."""
import math
import random

import numpy as np
import pytest

from oracle import models_ref as M
from oracle import phylo_ctmc_ref as R
from phylogenetics_b200 import _lib, phylo_ctmc, tree as T
from phylogenetics_b200.felsenstein import Make, StringAlignment
from phylogenetics_b200.models import Site_evolution_model

pytestmark = pytest.mark.gpu


def check_likelihood(a, b):
    """Test_utils.check_likelihood, tests/test_utils.ml:14-21: relative 1e-5."""
    assert abs(a - b) / abs(a) <= 0.00001


def simulate_strings(tree, tpm, nsites, rng):
    """Sequence_generation-style simulation along a Phylogenetic_tree (JC69/K80: uniform root)."""
    names = []

    def walk(t, states):
        if isinstance(t, T.PhyloLeaf):
            names.append((t.index, "".join("ACGT"[s] for s in states)))
            return
        for l, sub in (t.left, t.right):
            P = tpm(l)
            walk(sub, [rng.choices(range(4), weights=P[s])[0] for s in states])
    walk(tree, [rng.randrange(4) for _ in range(nsites)])
    names.sort(key=lambda x: int(x[0][1:]))
    return [s for _, s in names]


def to_oracle_phylo(t):
    if isinstance(t, T.PhyloLeaf):
        return R.PhyloLeaf(t.index)
    return R.PhyloNode((t.left[0], to_oracle_phylo(t.left[1])), (t.right[0], to_oracle_phylo(t.right[1])))


def test_pruning_felsenstein_vs_phylo_ctmc():
    """tests/test_phylo_ctmc.ml:4-26,96 — JC69, tree_size 100, seq_size 10."""
    rng = random.Random(7)
    M_ = Site_evolution_model.JC69
    Align = StringAlignment("ACGT")
    F = Make(4, Align, M_)
    tree = T.make_random(100, rng)
    red = M.jc69_reduction()
    align = Align.of_string_list(simulate_strings(tree, lambda l: M.make_diag_tpm(red, l), 10, rng))
    felsenstein_result = F.felsenstein(None, tree, align)
    ctmc_tree = T.to_tree(tree)
    tpm = M_.transition_probability_matrix()
    root_frequencies = M_.stationary_distribution()
    ctmc_result = sum(
        phylo_ctmc.pruning(ctmc_tree, 4, lambda l: [("Mat", tpm(l))],
                           lambda leaf: Align.get_base(align, seq=leaf[1], pos=i), root_frequencies)
        for i in range(Align.length(align)))
    check_likelihood(felsenstein_result, ctmc_result)
    # and both equal the oracle's restatement of the reference
    get_base = lambda idx, s: Align.get_base(align, seq=idx, pos=s)
    expect = R.felsenstein(lambda l: M.make_diag_tpm(red, l), np.full(4, 0.25), get_base, 4, to_oracle_phylo(tree), 10)
    assert abs(felsenstein_result - expect) <= 1e-10 * abs(expect)
    assert abs(ctmc_result - expect) <= 1e-10 * abs(expect)
    assert abs(F.felsenstein_noshift(None, tree, align) - expect) <= 1e-10 * abs(expect)
    assert abs(F.felsenstein_single_shift(None, 3, tree, align)
               - R.felsenstein_single(lambda l: M.make_diag_tpm(red, l), np.full(4, 0.25), get_base, 4,
                                      to_oracle_phylo(tree), 3,
                                      shift=lambda a, b, v: R.shift_normal(1e-7, a, b, v))) < 1e-12


@pytest.mark.parametrize("model,param", [("K80", 2.0), ("K80", 0.5), ("K80_numerical", 4.0), ("JC69_numerical", None)])
def test_felsenstein_models_of_the_bppml_suite(model, param):
    """tests/test_felsenstein.ml:42-59 runs these models against bppml (absent here): same models,
    tree sizes {10, 250} x seq sizes {1, 100}, expected value from the oracle instead."""
    rng = random.Random(11)
    Mod = getattr(Site_evolution_model, model)
    Align = StringAlignment("ACGT")
    F = Make(4, Align, Mod)
    Q = Mod.rate_matrix(param)
    from oracle import linalg_ref as la
    for treesize, seqsize in [(10, 1), (10, 100), (250, 1), (250, 100)]:
        tree = T.make_random(treesize, rng)
        align = Align.of_string_list(simulate_strings(tree, lambda l: la.expm(l * Q), seqsize, rng))
        got = F.felsenstein(param, tree, align)
        expect = R.felsenstein(lambda l: la.expm(l * Q), np.full(4, 0.25),
                               lambda idx, s: Align.get_base(align, seq=idx, pos=s), 4, to_oracle_phylo(tree), seqsize)
        check_likelihood(got, expect)
        assert abs(got - expect) <= 1e-10 * abs(expect)


def test_mcmc_fixture_through_the_mirror():
    """lib/mCMC.ml:22,32-38: Felsenstein(K80) 2.0 on "0.1;0.1;0.1;0.1;0;1;3.0;0.1;2;3" with A,A,A,T."""
    Align = StringAlignment("ACGT")
    F = Make(4, Align, Site_evolution_model.K80)
    got = F.felsenstein(2.0, T.of_preorder("0.1;0.1;0.1;0.1;0;1;3.0;0.1;2;3"), Align.of_string_list(["A", "A", "A", "T"]))
    assert abs(got - (-5.704499092002674)) < 1e-12


def test_conditional_likelihoods_tree():
    """Phylo_ctmc.conditional_likelihoods (lib/phylo_ctmc.ml:100-120): same shape, SV per node, (b, P) per branch."""
    t = T.binary_node(None,
                      T.branch(0.4, T.binary_node(None, T.branch(0.8, T.leaf("A1")), T.branch(1.2, T.leaf("A2")))),
                      T.branch(0.1, T.binary_node(None, T.branch(0.6, T.leaf("B")), T.branch(0.3, T.leaf("C")))))
    # the fixed 4-leaf tree of tests/expect/phylo_ctmc_conditional_simulation.ml:7-17
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    Q = M.gtr(pi, M.make_symetric(4, lambda a, b: float(a + 7 + b + 7) / 20.0))
    from oracle import linalg_ref as la
    states = {"A1": 0, "A2": 1, "B": 2, "C": 2}
    got = phylo_ctmc.conditional_likelihoods(t, 4, lambda l: states[l], lambda bl: la.expm(bl * Q))
    ot = R.binary_node(None,
                       R.Branch(0.4, R.binary_node(None, R.Branch(0.8, R.Leaf("A1")), R.Branch(1.2, R.Leaf("A2")))),
                       R.Branch(0.1, R.binary_node(None, R.Branch(0.6, R.Leaf("B")), R.Branch(0.3, R.Leaf("C")))))
    exp = R.conditional_likelihoods(ot, 4, lambda l: states[l], lambda bl: la.expm(bl * Q))

    def walk(a, b):
        if isinstance(b, R.Leaf):
            assert isinstance(a, T.Leaf) and a.data == b.data
            return
        np.testing.assert_allclose(a.data[0], b.data[0], rtol=1e-12)
        assert abs(a.data[1] - b.data[1]) < 1e-12
        for x, y in zip(a.branches, b.branches):
            assert x.data[0] == y.data[0]
            np.testing.assert_allclose(x.data[1], y.data[1], rtol=1e-13)
            walk(x.tip, y.tip)
    walk(got, exp)


def test_matrix_decomposition_closure_and_errors():
    """A `Mat; `Diag; `Mat decomposition (lib/phylo_ctmc.ml:4-8) gives the same value as the dense product."""
    t = T.binary_node(None, T.branch(0.3, T.leaf(0)), T.branch(0.7, T.leaf(1)))
    red = M.jc69_reduction()
    pi = np.full(4, 0.25)
    dec = lambda l: [("Mat", red[0]), ("Diag", np.exp(l * red[1])), ("Mat", red[2])]
    a = phylo_ctmc.pruning(t, 4, dec, lambda leaf: 2, pi)
    assert abs(a - (-2.189931069177981)) < 1e-14          # closed form, SURVEY.md 8c
    b = phylo_ctmc.pruning(t, 4, dec, lambda leaf: 0 if leaf == 0 else 3, pi)
    assert abs(b - (-3.078566665552957)) < 1e-14
    with pytest.raises(_lib.PhyloB200InvalidArgument):
        phylo_ctmc.pruning(t, 4, lambda l: np.eye(3), lambda leaf: 0, pi)
    with pytest.raises(_lib.PhyloB200InvalidArgument):
        phylo_ctmc.pruning(t, 4, dec, lambda leaf: 5, pi)


def test_missing_values_and_ambiguous_mirrors():
    """Phylo_ctmc.Missing_values.pruning and Phylo_ctmc.Ambiguous.pruning through the host mirror (names and
    closures as in lib/phylo_ctmc.mli:91-160), against the oracle's restatement of lib/phylo_ctmc.ml:145-172, 280-302,
    on the 4-leaf tree of tests/expect/phylo_ctmc_conditional_simulation.ml:7-17 and a random 40-leaf tree."""
    from oracle import linalg_ref as la
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    Q = M.gtr(pi, M.make_symetric(4, lambda a, b: float(a + 7 + b + 7) / 20.0))
    tp = lambda bl: [("Mat", la.expm(bl * Q))]
    rng = random.Random(5)

    def both(build):
        return build(T.leaf, T.branch, T.binary_node), build(R.Leaf, R.Branch, R.binary_node)

    def four(leaf, branch, node):
        return node(None, branch(0.4, node(None, branch(0.8, leaf("A1")), branch(1.2, leaf("A2")))),
                    branch(0.1, node(None, branch(0.6, leaf("B")), branch(0.3, leaf("C")))))

    def random_shape(n):
        shape = list(range(n))
        while len(shape) > 1:
            i, j = sorted(rng.sample(range(len(shape)), 2))
            b = shape.pop(j)
            a = shape.pop(i)
            shape.append((a, rng.uniform(0.01, 0.8), b, rng.uniform(0.01, 0.8)))
        return shape[0]

    shape = random_shape(40)

    def big(leaf, branch, node):
        def rec(s):
            if not isinstance(s, tuple):
                return leaf(s)
            return node(None, branch(s[1], rec(s[0])), branch(s[3], rec(s[2])))
        return rec(shape)

    for build, names in ((four, ["A1", "A2", "B", "C"]), (big, list(range(40)))):
        t, ot = both(build)
        for trial in range(6):
            obs = {l: (None if rng.random() < 0.3 else rng.randrange(4)) for l in names}
            if trial == 0:
                obs = {l: None for l in names}                    # nothing observed: log-likelihood 0
            got = phylo_ctmc.Missing_values.pruning(t, 4, tp, lambda l: obs[l], pi)
            exp = R.missing_values_pruning(ot, 4, tp, lambda l: obs[l], pi)
            assert (got == 0.0 and exp == 0.0) if trial == 0 else abs(got - exp) <= 1e-10 * abs(exp)
            sets = {l: {i for i in range(4) if rng.random() < 0.4} for l in names}
            got = phylo_ctmc.Ambiguous.pruning(t, 4, tp, lambda l, i: i in sets[l], pi)
            exp = R.ambiguous_pruning(ot, 4, tp, lambda l: (lambda i: i in sets[l]), pi)
            assert abs(got - exp) <= 1e-10 * max(abs(exp), 1e-300)
    # the batched forms: one call for all sites
    t, ot = both(big)
    nsites = 33
    obs = [{l: (None if rng.random() < 0.25 else rng.randrange(4)) for l in range(40)} for _ in range(nsites)]
    got = phylo_ctmc.Missing_values.pruning_sites(t, 4, tp, lambda l, s: obs[s][l], nsites, pi)
    exp = np.array([R.missing_values_pruning(ot, 4, tp, lambda l: obs[s][l], pi) for s in range(nsites)])
    np.testing.assert_allclose(got, exp, rtol=1e-10)
