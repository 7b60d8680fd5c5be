"""Property test: random trees of any arity (Tree.t allows >= 1 branch per node, lib/tree.ml:3-8), random branch
lengths from very short to very long, random alignments (uniform states: many impossible-looking columns, deep
rescaling), random category counts and both modes, 4 / 20 / 61 states — every default CUDA path against the oracle."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import cases
from oracle import c_oracle as C
from phylogenetics_b200 import _lib, models, synthetic, tree as T
from phylogenetics_b200.engine import LikelihoodContext

pytestmark = pytest.mark.gpu


def random_kary_tree(rng, ntaxa, max_arity):
    """Random joins of 1..max_arity subtrees (unary nodes included) until one tree is left."""
    forest = [T.leaf(i) for i in range(ntaxa)]
    while len(forest) > 1:
        k = int(min(len(forest), rng.integers(1 if len(forest) > 2 else 2, max_arity + 1)))
        idx = rng.choice(len(forest), size=k, replace=False)
        joined = T.node(None, [T.branch(float(10.0 ** rng.uniform(-4, 0.7)), forest[i]) for i in idx])
        forest = [x for q, x in enumerate(forest) if q not in set(idx.tolist())] + [joined]
    t = forest[0]
    return t if not isinstance(t, T.Leaf) else T.node(None, [T.branch(0.1, t)])


def model(kind, rng):
    if kind == 4:
        return synthetic.random_gtr(rng, 3.0)
    if kind == 20:
        w = cases.wag()
        pi = rng.dirichlet(np.full(20, 2.0))
        return models.Rate_matrix.scaled_rate_matrix(pi, models.Rate_matrix.make(20, lambda i, j: w.rate_matrix[i, j] * pi[j])), pi
    np_ = models.Nucleotide_process.random_gtr(rng, 10.0)
    p = dict(nucleotide_rates=models.Nucleotide_process.rate_matrix(np_),
             nucleotide_stat_dist=models.Nucleotide_process.stationary_distribution(np_), omega=float(rng.uniform(0.1, 2.0)))
    return models.MG94.rate_matrix(p), models.MG94.stationary_distribution(p)


@settings(max_examples=80, deadline=None, derandomize=True)
@given(seed=st.integers(0, 10**6), ntaxa=st.integers(2, 60), nsites=st.integers(1, 700), max_arity=st.integers(2, 5),
       nstates=st.sampled_from([4, 4, 20, 61]), ncat=st.integers(1, 3), felsenstein=st.booleans())
def test_random_problem_matches_oracle(seed, ntaxa, nsites, max_arity, nstates, ncat, felsenstein):
    rng = np.random.default_rng(seed)
    flat = T.flatten(random_kary_tree(rng, ntaxa, max_arity))
    bl = flat.branch_lengths(float)
    Q, pi = model(nstates, rng)
    rates, weights = models.gamma_category_rates(float(rng.uniform(0.2, 2.0)), ncat)
    P = synthetic.host_transition_matrices(Q, pi, bl, rates)
    tips = rng.integers(0, nstates, size=(flat.nleaves, nsites), dtype=np.uint8)
    mode, thr, root_mode = (_lib.PB_MODE_FELSENSTEIN, 1e-7, 0) if felsenstein else (_lib.PB_MODE_PHYLO_CTMC, 1e-6, 1)
    expect = C.pruning(nstates, flat.root, flat.child_ptr, flat.child_idx, flat.tip_row, P, pi, tips=tips,
                       weights=weights, threshold=thr, root_mode=root_mode)
    with LikelihoodContext(nstates, nsites, flat.nleaves, ncat=ncat) as ctx:
        ctx.set_tree(flat, bl)
        ctx.set_mode(mode)
        ctx.set_categories(rates, weights)
        ctx.set_pmatrices(P)
        ctx.set_tips(tips)
        ctx.set_root_frequencies(pi)
        total, per_site = ctx.loglik(per_site=True)
    assert np.all(np.isfinite(per_site))
    rel = np.abs(per_site - expect) / np.abs(expect)
    assert rel.max() <= 1e-9, (ctx.path_name, rel.max())
    assert abs(total - float(np.sum(expect))) <= 1e-10 * abs(float(np.sum(expect)))
