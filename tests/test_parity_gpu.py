"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Tolerances are the north star's: 1e-9 relative per site, 1e-10
relative on the total (BASELINE.json)."""
import math
import zlib

import numpy as np
import pytest

import cases
from oracle import c_oracle as C
from phylogenetics_b200 import _lib
from phylogenetics_b200.engine import LikelihoodContext

pytestmark = pytest.mark.gpu

PER_SITE_RTOL = 1e-9
TOTAL_RTOL = 1e-10


def oracle_per_site(case, P, mode=_lib.PB_MODE_PHYLO_CTMC):
    thr, root_mode = (1e-6, 1) if mode == _lib.PB_MODE_PHYLO_CTMC else (1e-7, 0)
    f = case.flat
    return C.pruning(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi, tips=case.tips,
                     weights=case.weights, threshold=thr, root_mode=root_mode)


def check(per_site, total, expect):
    assert np.all(np.isfinite(per_site))
    rel = np.abs(per_site - expect) / np.abs(expect)
    assert rel.max() <= PER_SITE_RTOL, "per-site rel err %.3e" % rel.max()
    truth = math.fsum(expect)
    assert abs(total - truth) <= TOTAL_RTOL * abs(truth), "total rel err %.3e" % (abs(total - truth) / abs(truth))


def run_ctx(case, source, level=False, keep=False, mode=_lib.PB_MODE_PHYLO_CTMC, P=None):
    ctx = LikelihoodContext(case.n, case.nsites, case.flat.nleaves, ncat=case.ncat, level_kernels=level, keep_clv=keep)
    ctx.set_tree(case.flat, case.bl)
    ctx.set_mode(mode)
    ctx.set_categories(case.rates, case.weights)
    if source == "eigen":
        ctx.set_rate_matrix(case.Q, case.pi)
    elif source == "expm":
        ctx.set_rate_matrix_general(case.Q)
    else:
        ctx.set_pmatrices(P)
    ctx.set_tips(case.tips)
    ctx.set_root_frequencies(case.pi)
    return ctx


CONFIGS = [
    # name, model, ntaxa, nsites, ncat
    ("c1_jc69_16x1000", "jc69", 16, 1000, 1),           # BASELINE configs[0]
    ("k80_33x777", "k80", 33, 777, 1),                  # ragged site count
    ("c2s_gtr_g4_128x4096", "gtr", 128, 4096, 4),       # configs[1] shape, fewer sites
    ("c3s_wag_64x1500", "wag", 64, 1500, 1),            # configs[2] shape
    ("c4s_mg94_32x700", "mg94", 32, 700, 1),            # configs[3] shape
    ("c4s_mutsel_32x300", "mutsel", 32, 300, 1),
]


@pytest.mark.parametrize("name,model,ntaxa,nsites,ncat", CONFIGS)
@pytest.mark.parametrize("level", [False, True])
def test_same_matrices_parity(name, model, ntaxa, nsites, ncat, level):
    """Identical P matrices fed to both sides ([`Mat p] route): pruning + rescaling + root only."""
    case = cases.make_case(name, model, ntaxa, nsites, ncat=ncat, seed=zlib.crc32(name.encode()) % 1000)
    P = case.host_P()
    with run_ctx(case, "direct", level=level, P=P) as ctx:
        total, per_site = ctx.loglik(per_site=True)
        assert ctx.kernel_launches > 0
    check(per_site, total, oracle_per_site(case, P))


@pytest.mark.parametrize("name,model,ntaxa,nsites,ncat", CONFIGS)
def test_eigen_pmatrices_parity(name, model, ntaxa, nsites, ncat):
    """P(t) rebuilt on the device from the eigensystem (K1) vs the oracle's Pade expm restatement."""
    from oracle import linalg_ref as la
    case = cases.make_case(name, model, ntaxa, nsites, ncat=ncat, seed=zlib.crc32(name.encode()) % 1000)
    f = case.flat
    Pref = np.zeros((case.ncat, f.nnodes, case.n, case.n))
    for k, r in enumerate(case.rates):
        for v in range(f.nnodes):
            if v != f.root:
                Pref[k, v] = la.expm((case.bl[v] * r) * case.Q)
    with run_ctx(case, "eigen") as ctx:
        Pdev = ctx.get_pmatrices()
        total, per_site = ctx.loglik(per_site=True)
    mask = np.ones(f.nnodes, bool)
    mask[f.root] = False
    assert np.abs(Pdev[:, mask] - Pref[:, mask]).max() < 5e-14
    check(per_site, total, oracle_per_site(case, Pref))


def test_felsenstein_mode():
    case = cases.make_case("fels", "k80", 40, 500, tree_kind="random")
    P = case.host_P()
    for level in (False, True):
        with run_ctx(case, "direct", level=level, mode=_lib.PB_MODE_FELSENSTEIN, P=P) as ctx:
            total, per_site = ctx.loglik(per_site=True)
        check(per_site, total, oracle_per_site(case, P, _lib.PB_MODE_FELSENSTEIN))


def test_conditional_likelihoods_match_oracle():
    case = cases.make_case("clv", "gtr", 24, 64)
    P = case.host_P()
    f = case.flat
    with run_ctx(case, "direct", keep=True, P=P) as ctx:
        ctx.loglik()
        for site in (0, 17, 63):
            clv, car = C.conditional_likelihoods(4, f.root, f.child_ptr, f.child_idx, f.tip_row, P[0], site,
                                                 tips=case.tips)
            for v in range(f.nnodes):
                gv, gc = ctx.get_clv(v)
                np.testing.assert_allclose(gv[site], clv[v], rtol=1e-12, atol=0)
                np.testing.assert_allclose(gc[site], car[v], rtol=1e-12, atol=1e-15)


def test_kary_and_unary_nodes():
    """Tree.t has arbitrary arity >= 1 (lib/tree.ml:3-8): left fold order over the children."""
    from phylogenetics_b200 import tree as T, synthetic
    rng = np.random.default_rng(5)
    t = T.node(None, [
        T.branch(0.11, T.leaf(0)),
        T.branch(0.2, T.node(None, [T.branch(0.05, T.leaf(1)), T.branch(0.3, T.leaf(2)), T.branch(0.1, T.leaf(3)),
                                    T.branch(0.4, T.node(None, [T.branch(0.2, T.leaf(4)), T.branch(0.15, T.leaf(5))]))])),
        T.branch(0.07, T.node(None, [T.branch(0.3, T.node(None, [T.branch(0.1, T.leaf(6)), T.branch(0.2, T.leaf(7))]))])),
        T.branch(0.25, T.node(None, [T.branch(0.2, T.node(None, [T.branch(0.3, T.leaf(8)), T.branch(0.1, T.leaf(9))])),
                                     T.branch(0.1, T.node(None, [T.branch(0.2, T.leaf(10)), T.branch(0.6, T.leaf(11))])),
                                     T.branch(0.3, T.leaf(12))])),
    ])
    flat = T.flatten(t)
    bl = flat.branch_lengths(float)
    Q, pi = synthetic.random_gtr(rng, 10.0)
    P = synthetic.host_transition_matrices(Q, pi, bl)
    tips = synthetic.simulate_alignment(rng, flat, P, pi, 333)
    case = cases.Case("kary", flat, bl, Q, pi, tips)
    for level in (False, True):
        with run_ctx(case, "direct", level=level, P=P) as ctx:
            total, per_site = ctx.loglik(per_site=True)
        check(per_site, total, oracle_per_site(case, P))


def test_deep_rescaling_long_branches():
    """Caterpillar tree with long branches: rescaling fires at almost every node."""
    from phylogenetics_b200 import tree as T, synthetic
    rng = np.random.default_rng(11)
    t = T.leaf(0)
    for i in range(1, 200):
        t = T.binary_node(None, T.branch(float(rng.uniform(0.5, 3.0)), t), T.branch(float(rng.uniform(0.5, 3.0)), T.leaf(i)))
    flat = T.flatten(t)
    bl = flat.branch_lengths(float)
    Q, pi = synthetic.random_gtr(rng, 10.0)
    P = synthetic.host_transition_matrices(Q, pi, bl)
    tips = synthetic.simulate_alignment(rng, flat, P, pi, 257)
    case = cases.Case("caterpillar", flat, bl, Q, pi, tips)
    expect = oracle_per_site(case, P)
    assert expect.min() < -250            # far below what unscaled doubles of 4^-200 could carry safely
    for level in (False, True):
        with run_ctx(case, "direct", level=level, P=P) as ctx:
            total, per_site = ctx.loglik(per_site=True)
        check(per_site, total, expect)


def test_single_site_and_repeat_evaluation():
    case = cases.make_case("one", "jc69", 5, 1)
    P = case.host_P()
    with run_ctx(case, "eigen") as ctx:
        a = ctx.loglik()
        b = ctx.loglik()
        ctx.set_branch_lengths(case.bl * 1.5)
        c = ctx.loglik()
    assert a == b and a != c
    check(np.array([a]), a, oracle_per_site(case, P))


def test_errors_are_reported_not_thrown():
    with LikelihoodContext(4, 10, 3) as ctx:
        with pytest.raises(_lib.PhyloB200Error) as e:
            ctx.loglik()
        assert e.value.code == _lib.PB_ERR_STATE
        with pytest.raises(_lib.PhyloB200InvalidArgument):
            ctx.set_mode(7)
    with pytest.raises(_lib.PhyloB200InvalidArgument):
        LikelihoodContext(1, 10, 3)


@pytest.mark.parametrize("pow2", [1, 0])
@pytest.mark.parametrize("block,spt,exact", [(128, 1, 0), (256, 1, 1), (512, 1, 0), (128, 2, 1), (256, 2, 0), (128, 4, 0), (256, 4, 0), (128, 4, 1)])
def test_fused_variants(block, spt, exact, pow2):
    """Every compiled shape of the fused kernel in its three scaling modes — power-of-two factors (default),
    the reference's factor 1/max with the mantissa/exponent accumulator, and the reference's literal log
    carry — on a tree deep enough to use the stack and >32 tips (several packed words)."""
    case = cases.make_case("fusedv", "gtr", 97, 3001, ncat=2, seed=3)
    P = case.host_P()
    with run_ctx(case, "direct", P=P) as ctx:
        ctx.set_option("fused_const", 0)
        ctx.set_option("fused_block", block)
        ctx.set_option("fused_spt", spt)
        ctx.set_option("fused_pow2", pow2)
        ctx.set_option("fused_exact_log", exact)
        assert ctx.path_name == "fused4"
        total, per_site = ctx.loglik(per_site=True)
    check(per_site, total, oracle_per_site(case, P))


def test_out_of_range_tip_state_is_an_error():
    case = cases.make_case("bad", "jc69", 8, 100)
    tips = case.tips.copy()
    tips[3, 57] = 9
    for level in (False, True):
        with run_ctx(case, "eigen", level=level) as ctx:
            ctx.set_tips(tips)
            with pytest.raises(_lib.PhyloB200InvalidArgument):
                ctx.loglik()


@pytest.mark.parametrize("model,ntaxa,nsites", [("wag", 40, 2100), ("mg94", 24, 900)])
@pytest.mark.parametrize("opts", [{}, {"dmma_version": 1}, {"dmma_per_level": 0}, {"dmma_per_level": 0, "dmma_group_tiles": 8}, {"dense_scalar": 1}])
def test_dense_kernel_modes(model, ntaxa, nsites, opts):
    """DMMA kernel with per-level launches (default), in its persistent tree-walking mode, with a
    different group size, and the scalar DFMA kernel: all against the oracle, with and without CLVs kept."""
    case = cases.make_case("dm" + model, model, ntaxa, nsites, seed=5)
    P = case.host_P()
    expect = oracle_per_site(case, P)
    for keep in (False, True):
        with run_ctx(case, "direct", level=True, keep=keep, P=P) as ctx:
            for k, v in opts.items():
                ctx.set_option(k, v)
            assert ctx.path_name == ("level_dense" if "dense_scalar" in opts else "level_dmma")
            total, per_site = ctx.loglik(per_site=True)
            check(per_site, total, expect)
            if keep:
                f = case.flat
                clv, car = C.conditional_likelihoods(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P[0], 5,
                                                     tips=case.tips)
                for v in range(f.nnodes):
                    gv, gc = ctx.get_clv(v)
                    np.testing.assert_allclose(gv[5], clv[v], rtol=1e-11, atol=1e-300)
                    np.testing.assert_allclose(gc[5], car[v], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("model,ntaxa,nsites,ncat,tree_kind", [
    ("wag", 64, 1500, 1, "bd"), ("wag", 40, 2100, 3, "random"), ("wag", 150, 333, 2, "bd"), ("wag", 3, 65, 1, "bd"),
    ("mg94", 32, 700, 1, "bd"), ("mg94", 24, 900, 2, "random"), ("mutsel", 70, 129, 1, "bd"), ("mg94", 2, 64, 1, "bd"),
])
@pytest.mark.parametrize("version", [3, 2, 1])
def test_fused_dmma(model, ntaxa, nsites, ncat, tree_kind, version):
    """The site-resident tensor-core kernel (default path for 20 / 61 states): birth-death and random
    trees, several categories, ragged site counts, trees deep enough that the stack spills out of
    shared memory (61 states keeps at most 3 levels on chip), Felsenstein mode, and the fall-back switch."""
    case = cases.make_case("fd" + model, model, ntaxa, nsites, ncat=ncat, seed=7, tree_kind=tree_kind)
    P = case.host_P()
    expect = oracle_per_site(case, P)
    with run_ctx(case, "direct", P=P) as ctx:
        ctx.set_option("fused_dmma_version", version)
        assert ctx.path_name == "fused_dmma"
        total, per_site = ctx.loglik(per_site=True)
        check(per_site, total, expect)
        total2, per_site2 = ctx.loglik(per_site=True)            # bit-reproducible
        assert total2 == total and np.array_equal(per_site, per_site2)
        if version == 3:                                        # default: nodes with two leaf children are tables
            ctx.set_option("fused_dmma_cherry", 0)              # ... and without them (scalar table build vs DMMA order:
            total5, per_site5 = ctx.loglik(per_site=True)       # equal to rounding, not to the bit)
            check(per_site5, total5, expect)
            np.testing.assert_allclose(per_site5, per_site, rtol=1e-12)
            ctx.set_option("fused_dmma_cherry", 1)
            ctx.set_option("fused_dmma_builder_waves", 3)       # the table builder cut into more, shorter blocks: same tables
            total7, per_site7 = ctx.loglik(per_site=True)
            assert total7 == total and np.array_equal(per_site7, per_site)
            ctx.set_option("fused_dmma_builder_waves", 0)
            ctx.set_option("fused_dmma_triple", 1)              # three-leaf clades as tables too (20 states; off by default)
            total6, per_site6 = ctx.loglik(per_site=True)
            check(per_site6, total6, expect)
            np.testing.assert_allclose(per_site6, per_site, rtol=1e-12)
            ctx.set_option("fused_dmma_triple", 0)
        ctx.set_option("fused_pow2", 0)                         # the reference's scale factor 1/max instead of 2^-k
        total4, per_site4 = ctx.loglik(per_site=True)
        check(per_site4, total4, expect)
        ctx.set_option("fused_dmma", 0)
        assert ctx.path_name == "level_dmma"
        total3, per_site3 = ctx.loglik(per_site=True)
        check(per_site3, total3, expect)
    with run_ctx(case, "direct", mode=_lib.PB_MODE_FELSENSTEIN, P=P) as ctx:
        ctx.set_option("fused_dmma_version", version)
        total, per_site = ctx.loglik(per_site=True)
    check(per_site, total, oracle_per_site(case, P, _lib.PB_MODE_FELSENSTEIN))


def test_fused_dmma_balanced_tree_spills_the_stack():
    """A perfectly balanced 64-taxon tree needs 5 stack levels: more than shared memory holds for 61 states."""
    from phylogenetics_b200 import tree as T, synthetic, models
    rng = np.random.default_rng(21)
    nodes = [T.leaf(i) for i in range(64)]
    while len(nodes) > 1:
        nodes = [T.binary_node(None, T.branch(float(rng.uniform(0.02, 0.4)), nodes[i]),
                               T.branch(float(rng.uniform(0.02, 0.4)), nodes[i + 1])) for i in range(0, len(nodes), 2)]
    flat = T.flatten(nodes[0])
    bl = flat.branch_lengths(float)
    np_ = models.Nucleotide_process.random_gtr(rng, 10.0)
    p = dict(nucleotide_rates=models.Nucleotide_process.rate_matrix(np_),
             nucleotide_stat_dist=models.Nucleotide_process.stationary_distribution(np_), omega=0.4)
    Q, pi = models.MG94.rate_matrix(p), models.MG94.stationary_distribution(p)
    P = synthetic.host_transition_matrices(Q, pi, bl)
    tips = synthetic.simulate_alignment(rng, flat, P, pi, 200)
    case = cases.Case("balanced", flat, bl, Q, pi, tips)
    expect = oracle_per_site(case, P)
    for version, stagger, spill in ((3, 0, 0), (3, 0, 1), (2, 0, 0), (2, 2000, 0), (2, 0, 1), (1, 0, 0)):
        with run_ctx(case, "direct", P=P) as ctx:
            ctx.set_option("fused_dmma_version", version)
            ctx.set_option("fused_dmma_stagger", stagger)
            ctx.set_option("fused_dmma_spill", spill)
            assert ctx.path_name == "fused_dmma"
            total, per_site = ctx.loglik(per_site=True)
        check(per_site, total, expect)


@pytest.mark.parametrize("model,ntaxa,nsites,ncat", [("wag", 150, 5000, 2), ("mg94", 32, 20000, 1)])
def test_fused_dmma_many_passes_are_reproducible(model, ntaxa, nsites, ncat):
    """Several rounds and categories per CTA: the ring hand-over between decoupled warps (a slot must not be
    refilled before every warp's fragment loads have returned) is exercised under load; repeated
    evaluations must agree bit for bit with each other and with the CTA-synchronous kernel."""
    case = cases.make_case("fd" + model, model, ntaxa, nsites, ncat=ncat, seed=7)
    P = case.host_P()
    with run_ctx(case, "direct", P=P) as ctx:
        ctx.set_option("fused_dmma_version", 1)
        ref = ctx.loglik(per_site=True)[1]
    check(ref, math.fsum(ref), oracle_per_site(case, P))
    for version in (3, 2):
        first = None
        for spill in (0, 1):
            with run_ctx(case, "direct", P=P) as ctx:
                ctx.set_option("fused_dmma_version", version)
                ctx.set_option("fused_dmma_spill", spill)
                for _ in range(6):
                    got = ctx.loglik(per_site=True)[1]
                    if version == 2:
                        assert np.array_equal(got, ref)       # same tile shape, same summation order as version 1
                    else:
                        first = got if first is None else first
                        assert np.array_equal(got, first)
                        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0)


@pytest.mark.parametrize("model,ntaxa,nsites", [("wag", 700, 40), ("mg94", 300, 24), ("wag", 2500, 9)])
def test_dense_alphabet_big_trees(model, ntaxa, nsites):
    """Trees whose per-warp tip tiles, program or stack outgrow shared memory: the planner must step down
    (fewer ring slots, spilled stack levels, the CTA-synchronous kernel, finally the level kernels) and stay exact."""
    case = cases.make_case("big" + model, model, ntaxa, nsites, seed=9)
    P = case.host_P()
    expect = oracle_per_site(case, P)
    with run_ctx(case, "direct", P=P) as ctx:
        assert ctx.path_name in ("fused_dmma", "level_dmma")
        total, per_site = ctx.loglik(per_site=True)
    check(per_site, total, expect)


def test_fused_dmma_kary_tree():
    """Multifurcating and unary nodes through the tensor-core program (TIP / MUL_TIP / APPLY / POP_MUL ops)."""
    from phylogenetics_b200 import tree as T, synthetic
    rng = np.random.default_rng(5)
    t = T.node(None, [
        T.branch(0.11, T.leaf(0)),
        T.branch(0.2, T.node(None, [T.branch(0.05, T.leaf(1)), T.branch(0.3, T.leaf(2)), T.branch(0.1, T.leaf(3)),
                                    T.branch(0.4, T.node(None, [T.branch(0.2, T.leaf(4)), T.branch(0.15, T.leaf(5))]))])),
        T.branch(0.07, T.node(None, [T.branch(0.3, T.node(None, [T.branch(0.1, T.leaf(6)), T.branch(0.2, T.leaf(7))]))])),
        T.branch(0.25, T.node(None, [T.branch(0.2, T.node(None, [T.branch(0.3, T.leaf(8)), T.branch(0.1, T.leaf(9))])),
                                     T.branch(0.1, T.node(None, [T.branch(0.2, T.leaf(10)), T.branch(0.6, T.leaf(11))])),
                                     T.branch(0.3, T.leaf(12))])),
    ])
    flat = T.flatten(t)
    bl = flat.branch_lengths(float)
    w = cases.wag()
    pi = w.freqs / w.freqs.sum()
    from phylogenetics_b200 import models
    Q = models.Rate_matrix.scaled_rate_matrix(pi, models.Rate_matrix.make(20, lambda i, j: w.rate_matrix[i, j] * pi[j]))
    P = synthetic.host_transition_matrices(Q, pi, bl)
    tips = synthetic.simulate_alignment(rng, flat, P, pi, 333)
    case = cases.Case("kary20", flat, bl, Q, pi, tips)
    for version in (3, 2, 1):
        with run_ctx(case, "direct", P=P) as ctx:
            ctx.set_option("fused_dmma_version", version)
            assert ctx.path_name == "fused_dmma"
            total, per_site = ctx.loglik(per_site=True)
        check(per_site, total, oracle_per_site(case, P))


@pytest.mark.parametrize("minb", [3, 4, 5, 6])
def test_level4_occupancy_variants(minb):
    case = cases.make_case("lv4", "gtr", 50, 5000, ncat=3, seed=6)
    P = case.host_P()
    with run_ctx(case, "direct", level=True, P=P) as ctx:
        ctx.set_option("level4_async", 0)
        ctx.set_option("level4_min_blocks", minb)
        total, per_site = ctx.loglik(per_site=True)
    check(per_site, total, oracle_per_site(case, P))


@pytest.mark.parametrize("nsites,chunks", [(20000, 8), (9001, 3), (50000, 1), (1000, 8)])
def test_pipelined_host_evaluation(nsites, chunks):
    """pb_loglik_host: chunked copy overlapped with pruning gives the same per-site values as the
    two-step pb_set_tips + pb_loglik, on ragged sizes and when it falls back (few sites)."""
    case = cases.make_case("host", "gtr", 40, nsites, ncat=4, seed=8)
    P = case.host_P()
    expect = oracle_per_site(case, P)
    with run_ctx(case, "direct", P=P) as ctx:
        ctx.set_option("host_chunks", chunks)
        total, per_site = ctx.loglik_host(case.tips, per_site=True)
        check(per_site, total, expect)
        total2, per_site2 = ctx.loglik(per_site=True)            # tips stay resident afterwards
        assert total2 == total and np.array_equal(per_site, per_site2)
        bad = case.tips.copy()
        bad[7, nsites - 1] = 4
        with pytest.raises(_lib.PhyloB200InvalidArgument):
            ctx.loglik_host(bad)
    # a level-path context falls back to copy-then-evaluate
    with run_ctx(case, "direct", level=True, P=P) as ctx:
        total3, per_site3 = ctx.loglik_host(case.tips, per_site=True)
        check(per_site3, total3, expect)


@pytest.mark.parametrize("model,ntaxa,nsites,ncat", [("gtr", 60, 3000, 4), ("wag", 20, 700, 1)])
def test_partial_reevaluation_after_one_branch_change(model, ntaxa, nsites, ncat):
    """f2: with the CLVs resident, changing one branch length recomputes only the path
    from that branch to the root; the result equals a full evaluation and the oracle."""
    case = cases.make_case("partial" + model, model, ntaxa, nsites, ncat=ncat, seed=31)
    f = case.flat
    rng = np.random.default_rng(2)
    with run_ctx(case, "eigen", keep=True) as ctx, run_ctx(case, "eigen") as full:
        ctx.loglik()
        assert ctx.nodes_recomputed == f.ninternal
        bl = case.bl.copy()
        for _ in range(4):
            v = int(rng.integers(1, f.nnodes))
            bl[v] = float(rng.uniform(0.01, 0.6))
            ctx.update_branch_length(v, bl[v])
            total, per_site = ctx.loglik(per_site=True)
            depth = 0
            u = v
            parent = {int(c): p for p in range(f.nnodes) for c in f.children(p)}
            while u in parent:
                u = parent[u]
                depth += 1
            assert ctx.nodes_recomputed == depth < f.ninternal
            full.set_branch_lengths(bl)
            t2, ps2 = full.loglik(per_site=True)
            np.testing.assert_allclose(per_site, ps2, rtol=1e-12)
            P = ctx.get_pmatrices()
            case2 = cases.Case("p", f, bl, case.Q, case.pi, case.tips, case.rates, case.weights)
            check(per_site, total, oracle_per_site(case2, P))
        # an unrelated change invalidates everything
        ctx.set_categories(case.rates, case.weights)
        ctx.loglik()
        assert ctx.nodes_recomputed == f.ninternal


@pytest.mark.parametrize("pow2", [1, 0])
@pytest.mark.parametrize("block,spt", [(256, 1), (256, 2), (256, 4), (512, 2), (128, 4)])
@pytest.mark.parametrize("ntaxa,ncat", [(97, 3), (128, 4), (130, 4), (300, 3), (12, 4)])
def test_fused_constant_memory_variant(block, spt, ntaxa, ncat, pow2):
    """fused4c (program and inner-branch matrices through the constant bank, one launch per category): trees
    with up to 128 inner branches use it, larger ones fall back to the shared-memory kernel; both scale-factor
    rules (power of two / the reference's 1/max)."""
    case = cases.make_case("fusedc", "gtr", ntaxa, 2500, ncat=ncat, seed=4)
    P = case.host_P()
    with run_ctx(case, "direct", P=P) as ctx:
        ctx.set_option("fused_const", 1)
        ctx.set_option("fused_block", block)
        ctx.set_option("fused_spt", spt)
        ctx.set_option("fused_pow2", pow2)
        assert ctx.path_name == "fused4"
        total, per_site = ctx.loglik(per_site=True)
        check(per_site, total, oracle_per_site(case, P))
        total2 = ctx.loglik_host(case.tips)
        assert total2 == total
        if pow2:                                            # the default runs the program over the REDUCED tree (fused4t);
            ctx.set_option("fused_reduced", 0)              # look-ups inside the full program (fused4c): bit-identical
            total0, per_site0 = ctx.loglik(per_site=True)
            assert total0 == total and np.array_equal(per_site0, per_site)
            ctx.set_option("fused_reduced", 1)
            total0, per_site0 = ctx.loglik(per_site=True)
            assert total0 == total and np.array_equal(per_site0, per_site)
        ctx.set_option("fused_const_cats", 2)               # two categories per launch share the constant bank
        total3, per_site3 = ctx.loglik(per_site=True)
        assert total3 == total and np.array_equal(per_site3, per_site)
        ctx.set_option("fused_clade_force", 1)              # (with 2 500 sites the library would stop at 4 leaves)
        for leaves in (6, 5, 3, 2):                         # largest tabulated clade (default 5): bit-identical whatever it is
            ctx.set_option("fused_clade_leaves", leaves)
            total5, per_site5 = ctx.loglik(per_site=True)
            assert total5 == total and np.array_equal(per_site5, per_site), leaves
        try:                                                # tables carrying their consumer's matrix: only in a build
            ctx.set_option("fused_preapply", 1)             # with -DFUSED_PREAPPLY=1
        except _lib.PhyloB200InvalidArgument:
            pass
        else:
            ctx.set_option("fused_clade_leaves", 5)
            total6, per_site6 = ctx.loglik(per_site=True)
            assert total6 == total and np.array_equal(per_site6, per_site)
            ctx.set_option("fused_preapply", 0)
        ctx.set_option("fused_triple", 0)                   # cherries only
        total5, per_site5 = ctx.loglik(per_site=True)
        assert total5 == total and np.array_equal(per_site5, per_site)
        ctx.set_option("fused_cherry", 0)                   # leaf-leaf nodes computed in the kernel instead of tabulated
        total4, per_site4 = ctx.loglik(per_site=True)
        assert total4 == total and np.array_equal(per_site4, per_site)


@pytest.mark.parametrize("ntaxa,nsites,ncat,tree_kind", [(256, 3000, 2, "bd"), (512, 2100, 1, "bd"), (1024, 1500, 1, "bd"),
                                                        (1500, 1100, 1, "bd"), (200, 5000, 4, "random"), (2, 1300, 4, "bd"),
                                                        (3, 1300, 1, "bd"), (7, 70000, 2, "bd")])
def test_fused_reduced_program_tree_sizes(ntaxa, nsites, ncat, tree_kind):
    """fused4t on trees from 2 to 1 500 taxa (only the reduced tree's matrices go to the constant bank, so with
    clades of up to 7 leaves even 1 500 taxa fit; smaller clade limits leave more matrices and may hand a large
    tree to another kernel), every clade limit, against the oracle, and against the untabulated kernels bit for bit
    whenever the site-resident path runs."""
    case = cases.make_case("fusedt", "gtr", ntaxa, nsites, ncat=ncat, seed=11, tree_kind=tree_kind)
    P = case.host_P()
    expect = oracle_per_site(case, P)
    with run_ctx(case, "direct", P=P) as ctx:
        ctx.set_option("fused_clade_force", 1)              # (with so few sites the library would stop at 3 leaves)
        assert ctx.path_name == "fused4"
        total, per_site = ctx.loglik(per_site=True)
        check(per_site, total, expect)
        settings = [("fused_clade_leaves", 8), ("fused_clade_leaves", 6), ("fused_clade_leaves", 4), ("fused_clade_leaves", 2),
                    ("fused_clade_leaves", 7), ("fused_cherry", 0), ("fused_reduced", 0), ("fused_const", 0)]
        for name, value in settings:                        # cumulative: ends on the shared-memory kernel without tables
            ctx.set_option(name, value)
            path = ctx.path_name
            t2, p2 = ctx.loglik(per_site=True)
            check(p2, t2, expect)
            if path == "fused4":                            # the site-resident kernels agree to the bit among themselves
                assert t2 == total and np.array_equal(p2, per_site), (name, value)
            else:
                assert ntaxa >= 1024 and path == "level4", (name, value, path)


def test_fused_reduced_program_kary_tree_and_felsenstein_mode():
    """Arities 1..4 (left folds, unary chains, clades of mixed arity) and the Felsenstein conventions."""
    from phylogenetics_b200 import tree as T, synthetic
    from test_fused_program_cpu import random_kary_tree
    for seed in range(4):
        rng = np.random.default_rng(300 + seed)
        flat = T.flatten(random_kary_tree(rng, 40 + 17 * seed))
        bl = flat.branch_lengths(float)
        Q, pi = synthetic.random_gtr(rng, 10.0)
        tips = synthetic.simulate_alignment(rng, flat, synthetic.host_transition_matrices(Q, pi, bl), pi, 2300)
        case = cases.Case("kary%d" % seed, flat, bl, Q, pi, tips)
        P = case.host_P()
        for mode in (_lib.PB_MODE_PHYLO_CTMC, _lib.PB_MODE_FELSENSTEIN):
            with run_ctx(case, "direct", P=P, mode=mode) as ctx:
                assert ctx.path_name == "fused4"
                ctx.set_option("fused_clade_force", 1)
                total, per_site = ctx.loglik(per_site=True)
                check(per_site, total, oracle_per_site(case, P, mode))
                ctx.set_option("fused_reduced", 0)
                t2, p2 = ctx.loglik(per_site=True)
                assert t2 == total and np.array_equal(p2, per_site)


@pytest.mark.parametrize("n", [2, 3, 7, 16, 21, 33, 60, 64])
def test_any_state_count(n):
    """Phylo_ctmc.pruning is generic in ~nstates (lib/phylo_ctmc.mli:55-61): alphabets without a specialised kernel
    (16 states, 21 = amino acids + gap, a 60-codon code ...) run the runtime-n level kernel; with tip state sets too."""
    from phylogenetics_b200 import tree as T, synthetic
    rng = np.random.default_rng(n)
    flat = T.flatten(synthetic.birth_death_tree(rng, 19))
    bl = flat.branch_lengths(float)
    pi = rng.dirichlet(np.full(n, 5.0))
    S = rng.gamma(1.0, 1.0, size=(n, n))
    S = S + S.T
    Q = S * pi[None, :]
    np.fill_diagonal(Q, 0.0)
    Q -= np.diag(Q.sum(1))
    Q /= -(pi * np.diag(Q)).sum()
    P = synthetic.host_transition_matrices(Q, pi, bl)
    tips = synthetic.simulate_alignment(rng, flat, P, pi, 777)
    case = cases.Case("n%d" % n, flat, bl, Q, pi, tips)
    f = case.flat
    for keep in (False, True):
        with run_ctx(case, "direct", keep=keep, P=P) as ctx:
            total, per_site = ctx.loglik(per_site=True)
            check(per_site, total, oracle_per_site(case, P))
    masks = np.uint64(1) << tips.astype(np.uint64)
    masks[rng.random(masks.shape) < 0.2] = 0
    masks[:, 5] = 0
    expect = C.pruning(n, f.root, f.child_ptr, f.child_idx, f.tip_row, P, pi, masks=masks)
    with LikelihoodContext(n, 777, f.nleaves) as ctx:
        ctx.set_tree(f, bl)
        ctx.set_pmatrices(P)
        ctx.set_tip_masks(masks)
        ctx.set_root_frequencies(pi)
        total, per_site = ctx.loglik(per_site=True)
    nz = expect != 0.0
    assert per_site[5] == 0.0 and (np.abs(per_site[nz] - expect[nz]) / np.abs(expect[nz])).max() <= 1e-9
    # P(t) on the device for that alphabet as well (eigen route and Pade route)
    for route in ("eigen", "expm"):
        with run_ctx(case, route) as ctx:
            total, per_site = ctx.loglik(per_site=True)
            check(per_site, total, oracle_per_site(case, ctx.get_pmatrices()))


def test_cuda_graph_replay_of_one_evaluation():
    """Option "graph": the optimisation-loop pattern — same tree and tips, new branch lengths / model / categories /
    frequencies every evaluation — replays one captured evaluation; results equal the plain launches to the bit, and a
    call that reshapes the launches (a clade limit, new tips) drops the graph and captures again."""
    from phylogenetics_b200 import models
    case = cases.make_case("graph", "gtr", 64, 30000, ncat=4, seed=31)
    rng = np.random.default_rng(1)
    settings = []
    for k in range(6):
        Q, pi = __import__("phylogenetics_b200.synthetic", fromlist=["x"]).random_gtr(rng, 10.0)
        rates, weights = models.gamma_category_rates(0.3 + 0.2 * k, 4)
        settings.append((case.bl * (0.8 + 0.1 * k), Q, pi, rates, weights))

    def run(graph):
        out = []
        with run_ctx(case, "eigen") as ctx:
            ctx.set_option("graph", graph)
            for i, (bl, Q, pi, rates, weights) in enumerate(settings):
                ctx.set_rate_matrix(Q, pi)
                ctx.set_root_frequencies(pi)
                ctx.set_categories(rates, weights)
                ctx.set_branch_lengths(bl)
                out.append(ctx.loglik(per_site=True))
                if i == 3:
                    ctx.set_option("fused_clade_leaves", 4)     # reshapes the launches: plain evaluation, new capture
            replays = ctx.graph_replays
            bad = case.tips.copy()
            bad[3, 17] = 9
            ctx.set_tips(bad)                                   # the range check still works after graph mode
            with pytest.raises(_lib.PhyloB200InvalidArgument):
                ctx.loglik()
        return out, replays
    plain, r0 = run(0)
    graphed, r1 = run(1)
    assert r0 == 0 and r1 == 4                                 # evaluations 1, 2, 3 and 5 replay; 0 and 4 are plain
    for (t0, p0), (t1, p1) in zip(plain, graphed):
        assert t0 == t1 and np.array_equal(p0, p1)
    P = None
    with run_ctx(case, "eigen") as ctx:                         # and they are right
        bl, Q, pi, rates, weights = settings[2]
        ctx.set_rate_matrix(Q, pi); ctx.set_root_frequencies(pi); ctx.set_categories(rates, weights); ctx.set_branch_lengths(bl)
        P = ctx.get_pmatrices()
    f = case.flat
    expect = C.pruning(4, f.root, f.child_ptr, f.child_idx, f.tip_row, P, settings[2][2], tips=case.tips, weights=settings[2][4])
    check(graphed[2][1], graphed[2][0], expect)


def test_fused_large_tree_falls_back_to_shared_memory_matrices():
    """More than 128 inner branches: the shared-memory kernel takes over."""
    case = cases.make_case("fusedbig", "gtr", 600, 700, ncat=1, seed=4)
    P = case.host_P()
    with run_ctx(case, "direct", P=P) as ctx:
        assert ctx.path_name == "fused4"
        total, per_site = ctx.loglik(per_site=True)
    check(per_site, total, oracle_per_site(case, P))


@pytest.mark.parametrize("waves", [1, 4, 16])
@pytest.mark.parametrize("nsites", [777, 40000])
def test_level4_async_pipeline(waves, nsites):
    """The cp.async-pipelined level kernel (binary trees) against the oracle, for several tile
    granularities, ragged sizes, Felsenstein mode and with every CLV kept."""
    case = cases.make_case("lv4a", "gtr", 70, nsites, ncat=4, seed=9)
    P = case.host_P()
    for mode in (_lib.PB_MODE_PHYLO_CTMC, _lib.PB_MODE_FELSENSTEIN):
        expect = oracle_per_site(case, P, mode)
        for keep in (False, True):
            with run_ctx(case, "direct", level=True, keep=keep, mode=mode, P=P) as ctx:
                ctx.set_option("level4_async", 1)
                ctx.set_option("level4_waves", waves)
                assert ctx.path_name == "level4"
                total, per_site = ctx.loglik(per_site=True)
                check(per_site, total, expect)


def test_two_contexts_share_the_device_constant_bank():
    """Two contexts with different trees and models interleave evaluations on one GPU: the fused
    kernel's constant-bank staging must not leak from one to the other."""
    a = cases.make_case("ctxa", "gtr", 60, 4000, ncat=4, seed=41)
    b = cases.make_case("ctxb", "k80", 33, 6000, ncat=2, seed=42)
    Pa, Pb = a.host_P(), b.host_P()
    ea, eb = oracle_per_site(a, Pa), oracle_per_site(b, Pb)
    with run_ctx(a, "direct", P=Pa) as ca, run_ctx(b, "direct", P=Pb) as cb:
        for _ in range(3):
            ca.enqueue()
            cb.enqueue()
            ta, pa = ca.fetch(per_site=True)
            tb, pb_ = cb.fetch(per_site=True)
            check(pa, ta, ea)
            check(pb_, tb, eb)


def test_timing_ring_keeps_queued_evaluations():
    """pb_get_timing_at: phase times of evaluations queued without synchronising, oldest to newest."""
    case = cases.make_case("timing", "gtr", 30, 20000, ncat=2, seed=4)
    with run_ctx(case, "eigen") as ctx:
        ctx.set_timing(True)
        with pytest.raises(_lib.PhyloB200Error):
            ctx.get_timing()
        for _ in range(5):
            ctx.enqueue()
        total, _ = ctx.fetch(per_site=True)
        times = [ctx.get_timing(back) for back in range(5)]
        assert all(len(t) == 3 and all(x >= 0.0 for x in t) and t[1] > 0.0 for t in times)
        with pytest.raises(_lib.PhyloB200InvalidArgument):
            ctx.get_timing(5)
        with pytest.raises(_lib.PhyloB200InvalidArgument):
            ctx.get_timing(64)
        assert ctx.get_timing() == times[0]


@pytest.mark.parametrize("masked,every_category", [(False, False), (False, True), (True, True)])
def test_table_exponents_that_do_not_fold(masked, every_category):
    """Table entries whose scale exponent cannot be folded into the vector (|k| > 480: here leaf branches whose
    off-diagonal transition probabilities are 1e-70, so a clade of mismatching leaves is rescaled by 2^-232 several
    times): the table builder raises its device-side flag and the general build of fused4t (vector AND exponent
    gathers) does the work instead of the FOLD build.  Against the oracle, and bit for bit against the evaluation with
    folding switched off and against the untabulated kernels; also through the (value, exponent) root reduction with
    categories whose exponents are thousands apart."""
    case = cases.make_case("nofold", "gtr", 48, 6000, ncat=3, seed=21)
    P = case.host_P()
    f = case.flat
    eps = np.full((4, 4), 1e-70)
    np.fill_diagonal(eps, 1.0)
    for v in range(f.nnodes):
        if f.child_ptr[v] == f.child_ptr[v + 1]:            # every leaf branch (of every category but the last: that one
            P[:None if every_category else -1, v] = eps      # then carries the mixture, thousands of binary orders above)
    masks = None
    if masked:
        rng = np.random.default_rng(5)
        masks = np.uint64(1) << case.tips.astype(np.uint64)
        masks[rng.random(masks.shape) < 0.1] = 0
    expect = C.pruning(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi,
                       tips=None if masked else case.tips, masks=masks, weights=case.weights, threshold=1e-6, root_mode=1)
    assert not every_category or expect.min() < -400.0      # (the point of the case: likelihoods far below 2^-480)
    with run_ctx(case, "direct", P=P) as ctx:
        if masked:
            ctx.set_tip_masks(masks)
        ctx.set_option("fused_clade_force", 1)
        assert ctx.path_name == "fused4"
        total, per_site = ctx.loglik(per_site=True)
        check(per_site, total, expect)
        for name, value in (("fused_fold", 0), ("fused_clade_leaves", 3), ("fused_fold", 1)) + \
                           ((() if masked else (("fused_reduced", 0), ("fused_const", 0)))):
            ctx.set_option(name, value)
            t2, p2 = ctx.loglik(per_site=True)
            assert ctx.path_name == "fused4"
            assert t2 == total and np.array_equal(p2, per_site), (name, value)


def test_kernel_only_timing():
    """pb_set_timing(ctx, 2) + pb_get_kernel_timing: the dominant kernel alone, per launch, inside the pruning phase."""
    for model, ntaxa, nsites, ncat in (("gtr", 40, 60000, 4), ("wag", 12, 20000, 1)):
        case = cases.make_case("ktime", model, ntaxa, nsites, ncat=ncat, seed=3)
        with run_ctx(case, "eigen" if model == "gtr" else "expm") as ctx:
            ctx.set_timing(True)
            total = ctx.loglik()
            with pytest.raises(_lib.PhyloB200Error):
                ctx.get_kernel_timing()                     # kernel timing is off
            ctx.set_timing(2)
            for _ in range(3):
                assert ctx.loglik() == total
            k_ms, launches = ctx.get_kernel_timing()
            phases = ctx.get_timing()
            assert launches >= 1 and 0.0 < k_ms <= phases[1] * 1.001
            ctx.set_timing(True)
