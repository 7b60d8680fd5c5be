"""f1: Phylo_ctmc.Missing_values.pruning and Phylo_ctmc.Ambiguous.pruning on the GPU
through tip state-set masks (pb_set_tip_masks), against the oracle's restatement of
lib/phylo_ctmc.ml:145-172 and :280-302."""
import numpy as np
import pytest

import cases
from oracle import c_oracle as C
from phylogenetics_b200.engine import LikelihoodContext

pytestmark = pytest.mark.gpu


def make_masks(case, rng, p_missing, p_ambiguous):
    n = case.n
    masks = (np.uint64(1) << case.tips.astype(np.uint64))
    amb = rng.random(masks.shape) < p_ambiguous
    extra = rng.integers(1, 1 << min(n, 20), size=masks.shape, dtype=np.uint64)
    masks = np.where(amb, masks | extra, masks)
    masks = np.where(rng.random(masks.shape) < p_missing, np.uint64(0), masks)
    return np.ascontiguousarray(masks, dtype=np.uint64)


@pytest.mark.parametrize("model,ntaxa,nsites,ncat", [("gtr", 30, 1500, 4), ("k80", 12, 257, 1), ("wag", 16, 300, 1),
                                                     ("mg94", 8, 130, 1)])
def test_missing_and_ambiguous_tips(model, ntaxa, nsites, ncat):
    rng = np.random.default_rng(17)
    case = cases.make_case("mask" + model, model, ntaxa, nsites, ncat=ncat, seed=17)
    P = case.host_P()
    f = case.flat
    masks = make_masks(case, rng, 0.25, 0.2)
    masks[:, 3] = 0                                    # an all-missing column: log-likelihood 0.
    masks[: ntaxa - 1, 5] = 0                          # a single observed leaf
    expect = C.pruning(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi, masks=masks,
                       weights=case.weights)
    assert expect[3] == 0.0
    for keep in (False, True):
        with LikelihoodContext(case.n, nsites, f.nleaves, ncat=ncat, keep_clv=keep) as ctx:
            ctx.set_tree(f, case.bl)
            ctx.set_categories(case.rates, case.weights)
            ctx.set_pmatrices(P)
            ctx.set_tip_masks(masks)
            ctx.set_root_frequencies(case.pi)
            # 4 states without kept CLVs: the site-resident reduced program (fused4t) with the alignment's state sets as
            # its alphabet; everything else: the scalar level kernel
            assert ctx.path_name == ("fused4" if case.n == 4 and not keep else "level_dense")
            total, per_site = ctx.loglik(per_site=True)
            assert per_site[3] == 0.0
            nz = expect != 0.0
            assert (np.abs(per_site[nz] - expect[nz]) / np.abs(expect[nz])).max() <= 1e-9
            assert abs(total - expect.sum()) <= 1e-10 * abs(expect.sum())
            if keep:
                clv, car = C.conditional_likelihoods(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P[0], 7,
                                                     masks=masks)
                for v in range(f.nnodes):
                    gv, gc = ctx.get_clv(v)
                    if np.isnan(car[v]):
                        assert np.isnan(gc[7])
                    else:
                        np.testing.assert_allclose(gv[7], clv[v], rtol=1e-11, atol=1e-300)
                        np.testing.assert_allclose(gc[7], car[v], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("ntaxa,nsites,ncat,p_missing,p_ambiguous,tree_kind",
                         [(128, 70001, 4, 0.10, 0.0, "bd"), (97, 40000, 2, 0.3, 0.05, "bd"), (33, 9000, 1, 0.0, 0.5, "random"),
                          (64, 20000, 4, 0.9, 0.0, "bd"), (5, 3000, 1, 0.2, 0.2, "bd"), (300, 12000, 1, 0.05, 0.01, "bd")])
def test_state_sets_in_the_site_resident_kernel(ntaxa, nsites, ncat, p_missing, p_ambiguous, tree_kind):
    """fused4t with MASKS: clades tabulated over the alignment's distinct state sets (mixed-radix indices), None
    carried through tables, stack and loose tips; against the oracle, against the level kernel, for several clade
    limits, and after going back to plain tips."""
    rng = np.random.default_rng(5)
    case = cases.make_case("maskf", "gtr", ntaxa, nsites, ncat=ncat, seed=19, tree_kind=tree_kind)
    P = case.host_P()
    f = case.flat
    masks = make_masks(case, rng, p_missing, p_ambiguous)
    masks[:, 11] = 0                                   # no observation at all
    masks[1:, 12] = 0                                  # a single observed leaf
    masks[:, 13] = 15                                  # every leaf fully ambiguous: likelihood 1
    expect = C.pruning(4, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi, masks=masks, weights=case.weights)
    assert expect[11] == 0.0 and abs(expect[13]) < 1e-12
    nz = expect != 0.0
    nz[13] = False                                     # log 1 up to rounding: compared absolutely below

    def agrees(per_site, total):
        assert abs(per_site[13]) < 1e-12
        assert per_site[11] == 0.0
        assert (np.abs(per_site[nz] - expect[nz]) / np.abs(expect[nz])).max() <= 1e-9
        assert abs(total - expect.sum()) <= 1e-10 * abs(expect.sum())
    with LikelihoodContext(4, nsites, f.nleaves, ncat=ncat) as ctx:
        ctx.set_tree(f, case.bl)
        ctx.set_categories(case.rates, case.weights)
        ctx.set_pmatrices(P)
        ctx.set_tip_masks(masks)
        ctx.set_root_frequencies(case.pi)
        assert ctx.path_name == "fused4"
        total, per_site = ctx.loglik(per_site=True)
        agrees(per_site, total)
        ctx.set_option("fused_clade_force", 1)
        for leaves in (8, 5, 3, 2):
            ctx.set_option("fused_clade_leaves", leaves)
            t2, p2 = ctx.loglik(per_site=True)
            agrees(p2, t2)
            np.testing.assert_allclose(p2[nz], per_site[nz], rtol=1e-13)
        ctx.set_option("fused_spt", 1)
        t2, p2 = ctx.loglik(per_site=True)
        agrees(p2, t2)
        ctx.set_option("fused_masks", 0)               # the level kernel on the same sets
        assert ctx.path_name == "level_dense"
        t3, p3 = ctx.loglik(per_site=True)
        agrees(p3, t3)
        np.testing.assert_allclose(p3[nz], per_site[nz], rtol=1e-12)
        ctx.set_option("fused_masks", 1)
        ctx.set_tips(case.tips)                        # back to plain states: the plan and the packed tips follow
        assert ctx.path_name == "fused4"
        t4, p4 = ctx.loglik(per_site=True)
        plain = C.pruning(4, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi, tips=case.tips, weights=case.weights)
        assert (np.abs(p4 - plain) / np.abs(plain)).max() <= 1e-9
        ctx.set_tip_masks(masks)                       # ... and to sets again
        t5, p5 = ctx.loglik(per_site=True)
        assert t5 == total and np.array_equal(p5, per_site)


def test_single_state_masks_equal_plain_tips():
    """A one-bit mask per tip is the ordinary pruning (tests/test_phylo_ctmc.ml:58-93 in spirit)."""
    case = cases.make_case("mask1", "gtr", 20, 999, ncat=2, seed=3)
    P = case.host_P()
    f = case.flat
    masks = np.ascontiguousarray(np.uint64(1) << case.tips.astype(np.uint64))
    out = []
    for use_masks in (False, True):
        with LikelihoodContext(4, case.nsites, f.nleaves, ncat=2) as ctx:
            ctx.set_tree(f, case.bl)
            ctx.set_categories(case.rates, case.weights)
            ctx.set_pmatrices(P)
            ctx.set_tip_masks(masks) if use_masks else ctx.set_tips(case.tips)
            ctx.set_root_frequencies(case.pi)
            out.append(ctx.loglik(per_site=True)[1])
    np.testing.assert_allclose(out[0], out[1], rtol=1e-13)
