"""f3: ancestral-state sampling on the device (pb_conditional_simulation) against the exact distribution
Phylo_ctmc.conditional_simulation draws from (lib/phylo_ctmc.ml:122-143; Missing_values :199-229;
Ambiguous :331-360), computed by enumeration in the oracle, on the 4-leaf tree of the reference's own
test (tests/expect/phylo_ctmc_conditional_simulation.ml:7-17).  The reference's printed frequencies depend
on GSL's stream and a random profile; what is compared here is the distribution."""
import collections

import numpy as np
import pytest

import cases
from oracle import phylo_ctmc_ref as R
from phylogenetics_b200 import _lib, synthetic, tree as T
from phylogenetics_b200.engine import LikelihoodContext

pytestmark = pytest.mark.gpu


def small_tree():
    return T.binary_node(None,
                         T.branch(0.4, T.binary_node(None, T.branch(0.8, T.leaf(0)), T.branch(1.2, T.leaf(1)))),
                         T.branch(0.1, T.binary_node(None, T.branch(0.6, T.leaf(2)), T.branch(0.3, T.leaf(3)))))


def gtr_like_the_reference_test(rng):
    """Exchangeabilities (a + 7 + b + 7) / 20 as in the reference test (:31-35), random profile."""
    pi = rng.dirichlet(np.ones(4))
    Q = np.zeros((4, 4))
    for a in range(4):
        for b in range(4):
            if a != b:
                Q[a, b] = (a + 7 + b + 7) / 20.0 * pi[b]
        Q[a, a] = -Q[a].sum()
    return Q, pi


def run(flat, bl, Q, pi, nsites, tips=None, masks=None, seed=1):
    P = synthetic.host_transition_matrices(Q, pi, bl)
    with LikelihoodContext(len(pi), nsites, flat.nleaves, keep_clv=True) as ctx:
        ctx.set_tree(flat, bl)
        ctx.set_pmatrices(P)
        if masks is not None:
            ctx.set_tip_masks(masks)
        else:
            ctx.set_tips(tips)
        ctx.set_root_frequencies(pi)
        ctx.loglik()
        return P[0], ctx.conditional_simulation(seed), ctx.conditional_simulation(seed), ctx.conditional_simulation(seed + 1)


def check_distribution(states, free, table, nrep):
    """Joint frequencies of the free nodes over nrep independent draws vs the exact table."""
    counts = collections.Counter(map(tuple, states[free].T.tolist()))
    assert set(counts) <= set(k for k, p in table.items() if p > 0)
    for key, p in table.items():
        sd = np.sqrt(p * (1 - p) / nrep)
        assert abs(counts.get(key, 0) / nrep - p) <= 5 * sd + 2e-5, (key, counts.get(key, 0) / nrep, p)


@pytest.mark.parametrize("column", [(2, 2, 2, 2), (0, 1, 2, 3), (3, 3, 0, 0)])
def test_joint_distribution_of_inner_nodes(column):
    rng = np.random.default_rng(12)
    flat = T.flatten(small_tree())
    bl = flat.branch_lengths(float)
    Q, pi = gtr_like_the_reference_test(rng)
    nrep = 200_000
    tips = np.repeat(np.array(column, dtype=np.uint8)[:, None], nrep, axis=1)
    P, s1, s1again, s2 = run(flat, bl, Q, pi, nrep, tips=tips)
    assert np.array_equal(s1, s1again) and not np.array_equal(s1, s2)         # keyed by the seed
    leaf_sets = {v: [int(column[flat.tip_row[v]])] for v in range(flat.nnodes) if flat.is_leaf(v)}
    for v, st in leaf_sets.items():
        assert np.all(s1[v] == st[0])                                         # observed leaves keep their state
    free, table = R.conditional_simulation_exact(flat.root, flat.child_ptr, flat.child_idx, P, pi, leaf_sets)
    assert len(free) == 3
    check_distribution(s1, free, table, nrep)
    check_distribution(s2, free, table, nrep)


def test_missing_and_ambiguous_leaves():
    """Leaf None -> drawn from the prior; an ambiguous leaf -> prior restricted to its set; a cherry with
    both leaves missing -> its parent has no conditional likelihood and is drawn from the prior."""
    rng = np.random.default_rng(13)
    flat = T.flatten(small_tree())
    bl = flat.branch_lengths(float)
    Q, pi = gtr_like_the_reference_test(rng)
    nrep = 200_000
    for sets in ([[2], None, [1], [1]], [[0], [1, 3], [2], None], [None, None, [3], [0, 1, 2]]):
        row_mask = [0 if s is None else sum(1 << i for i in s) for s in sets]
        masks = np.repeat(np.array(row_mask, dtype=np.uint64)[:, None], nrep, axis=1)
        P, s1, _, _ = run(flat, bl, Q, pi, nrep, masks=masks, seed=5)
        leaf_sets = {v: sets[flat.tip_row[v]] for v in range(flat.nnodes) if flat.is_leaf(v)}
        free, table = R.conditional_simulation_exact(flat.root, flat.child_ptr, flat.child_idx, P, pi, leaf_sets)
        check_distribution(s1, free, table, nrep)


def test_device_draws_match_the_oracle_sampler_in_distribution():
    """The oracle's restatement of the reference sampler (numpy stream) and the device sampler agree on
    the marginals of a 10-taxon WAG problem (too many states to enumerate the joint)."""
    case = cases.make_case("sim20", "wag", 10, 1, seed=3)
    flat = case.flat
    nrep = 40_000
    tips = np.repeat(case.tips[:, :1], nrep, axis=1)
    P, s1, _, _ = run(flat, case.bl, case.Q, case.pi, nrep, tips=tips, seed=9)
    with LikelihoodContext(20, 1, flat.nleaves, keep_clv=True) as ctx:
        ctx.set_tree(flat, case.bl)
        ctx.set_pmatrices(P[None])
        ctx.set_tips(case.tips[:, :1])
        ctx.set_root_frequencies(case.pi)
        ctx.loglik()
        clv = np.array([ctx.get_clv(v)[0][0] for v in range(flat.nnodes)])
        carry = np.array([ctx.get_clv(v)[1][0] for v in range(flat.nnodes)])
    leaf_sets = {v: [int(case.tips[flat.tip_row[v], 0])] for v in range(flat.nnodes) if flat.is_leaf(v)}
    rng = np.random.default_rng(1)
    ndraw = 4000
    ref = np.array([R.conditional_simulation(rng, flat.root, flat.child_ptr, flat.child_idx, P, clv, carry, case.pi, leaf_sets)
                    for _ in range(ndraw)]).T
    # root marginal has a closed form: pi_i * cl_root[i], normalised
    w = case.pi * clv[flat.root]
    w /= w.sum()
    froot = np.bincount(s1[flat.root], minlength=20) / nrep
    assert np.all(np.abs(froot - w) <= 5 * np.sqrt(w * (1 - w) / nrep) + 1e-4)
    for v in range(flat.nnodes):
        fd = np.bincount(s1[v], minlength=20) / nrep
        fr = np.bincount(ref[v], minlength=20) / ndraw
        assert np.all(np.abs(fd - fr) <= 5 * np.sqrt(fr * (1 - fr) / ndraw + 1e-6) + 1e-3)


def test_preconditions():
    case = cases.make_case("simerr", "gtr", 6, 50, ncat=2)
    with LikelihoodContext(4, 50, 6, ncat=2, keep_clv=True) as ctx:
        ctx.set_tree(case.flat, case.bl)
        ctx.set_categories(case.rates, case.weights)
        ctx.set_rate_matrix(case.Q, case.pi)
        ctx.set_tips(case.tips)
        ctx.set_root_frequencies(case.pi)
        with pytest.raises(_lib.PhyloB200Error):
            ctx.conditional_simulation(1)                  # nothing evaluated
        ctx.loglik()
        with pytest.raises(_lib.PhyloB200Error):
            ctx.conditional_simulation(1)                  # rate categories: not defined by the reference
    case = cases.make_case("simerr", "gtr", 6, 50)
    with LikelihoodContext(4, 50, 6) as ctx:
        ctx.set_tree(case.flat, case.bl)
        ctx.set_rate_matrix(case.Q, case.pi)
        ctx.set_tips(case.tips)
        ctx.set_root_frequencies(case.pi)
        ctx.loglik()
        with pytest.raises(_lib.PhyloB200Error):
            ctx.conditional_simulation(1)                  # conditional likelihoods were not kept


def test_reference_api_mirror():
    """phylo_ctmc.conditional_simulation_sites with the reference test's closures (leaf_state, transition_probabilities)."""
    from oracle import linalg_ref as la
    from phylogenetics_b200 import phylo_ctmc as PC
    rng = np.random.default_rng(12)
    Q, pi = gtr_like_the_reference_test(rng)
    column = (2, 0, 2, 1)
    nrep = 100_000
    flat, states = PC.conditional_simulation_sites(
        7, small_tree(), 4, lambda leaf, _site: column[leaf], lambda bl: [("Mat", la.expm(bl * Q))], nrep, pi)
    P = np.array([la.expm(b * Q) if v != flat.root else np.eye(4) for v, b in enumerate(flat.branch_lengths(float))])
    leaf_sets = {v: [column[flat.tip_row[v]]] for v in range(flat.nnodes) if flat.is_leaf(v)}
    free, table = R.conditional_simulation_exact(flat.root, flat.child_ptr, flat.child_idx, P, pi, leaf_sets)
    check_distribution(states, free, table, nrep)
