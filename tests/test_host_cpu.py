"""Host layer (trees, model builders, synthetic inputs) against the oracle's independent
restatements — no GPU involved."""
import numpy as np
import pytest

import cases
from oracle import c_oracle as C
from oracle import models_ref as M
from oracle import phylo_ctmc_ref as R
from phylogenetics_b200 import models, synthetic, tree as T


def test_flatten_roundtrip_and_branch_order():
    t = T.node("r", [T.branch(0.1, T.leaf("a")),
                     T.branch(0.2, T.node("x", [T.branch(0.3, T.leaf("b")), T.branch(0.4, T.leaf("c")),
                                                T.branch(0.5, T.leaf("d"))])),
                     T.branch(0.6, T.node("u", [T.branch(0.7, T.leaf("e"))]))])
    f = T.flatten(t)
    assert f.nnodes == 8 and f.root == 0 and f.nleaves == 5
    assert list(f.children(0)) == [1, 2, 6]               # reference branch order is preserved
    assert f.leaf_data == ["a", "b", "c", "d", "e"]
    assert T.leaves(t) == f.leaf_data
    assert list(f.branch_lengths(float)) == [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7]
    back = T.unflatten(f)
    assert T.leaves(back) == f.leaf_data and back.data == "r" and len(back.branches) == 3
    # the oracle rebuilds the same shape from the CSR arrays
    ot = C.tree_from_csr(f.root, f.child_ptr, f.child_idx, f.tip_row)
    assert len(ot.branches) == 3 and len(ot.branches[1].tip.branches) == 3
    with pytest.raises(ValueError):
        T.flatten(T.leaf("x"))
    with pytest.raises(ValueError):
        T.node(None, [])


def test_of_preorder_and_branch_length_accessors():
    t = T.of_preorder("0.1;0.1;0.1;0.1;0;1;3.0;0.1;2;3")       # lib/mCMC.ml:37
    assert T.get_branch_lengths(t) == [0.1, 0.1, 0.1, 0.1, 3.0, 0.1]
    t2 = T.set_branch_lengths(t, [1, 2, 3, 4, 5, 6])
    assert T.get_branch_lengths(t2) == [1, 2, 3, 4, 5, 6]
    f = T.flatten(T.to_tree(t))
    assert [d[1] for d in f.leaf_data] == ["T0", "T1", "T2", "T3"]
    with pytest.raises(ValueError):
        T.of_preorder("0.1;x")
    with pytest.raises(RuntimeError):
        T.set_branch_lengths(t, [1, 2, 3])
    r = T.make_random(10)
    assert T.nb_branches(r) == 18 and all(1e-5 <= x <= 0.50001 for x in T.get_branch_lengths(r))


def test_model_builders_match_oracle_restatement():
    rng = np.random.default_rng(5)
    pi = rng.dirichlet(np.full(4, 10.0))
    vals = rng.gamma(1.0, 1.0, size=(4, 4))
    ex_p = models.Rate_matrix.make_symetric(4, lambda i, j: vals[i, j])
    ex_o = M.make_symetric(4, lambda i, j: vals[i, j])
    assert np.array_equal(ex_p, ex_o)
    assert np.array_equal(models.Rate_matrix.gtr(pi, ex_p), M.gtr(pi, ex_o))
    assert np.array_equal(models.Rate_matrix.jc69(4), M.jc69(4))
    assert np.array_equal(models.Rate_matrix.Nucleotide.k80(2.0), M.k80(2.0))
    assert np.array_equal(models.Rate_matrix.Nucleotide.hky85(pi, 0.3, 0.7), M.hky85(pi, 0.3, 0.7))
    Q = models.Rate_matrix.gtr(pi, ex_p)
    assert np.allclose(models.Rate_matrix.stationary_distribution(Q), pi, rtol=1e-9)
    assert np.array_equal(models.Rate_matrix.scale(Q), M.scale(Q))
    with pytest.raises(RuntimeError):
        models.Rate_matrix.make(3, lambda i, j: -1.0)
    # codon tables and models
    assert models.Codon.ns_triplets == M.NS_TRIPLETS and list(models.Codon.ns_aa) == M.NS_AA
    assert models.Codon.card == 61 and models.Codon.neighbours_diff(0, 5) is None     # lib/codon.ml:42
    p = dict(nucleotide_rates=Q, nucleotide_stat_dist=pi, omega=0.7)
    assert np.array_equal(models.MG94.rate_matrix(p), M.mg94_rate_matrix(Q, 0.7))
    assert np.allclose(models.MG94.stationary_distribution(p), M.mg94_stationary_distribution(pi), rtol=1e-15)
    F = models.Mutsel.fitness_of_profile(rng.dirichlet(np.full(20, 10.0)))
    pm = dict(nucleotide_rates=Q, nucleotide_stat_dist=pi, omega=1.3, scaled_fitness=F, gBGC=0.2, pps=0.1)
    assert np.array_equal(models.Mutsel.rate_matrix(pm), M.mutsel_rate_matrix(Q, 1.3, F, 0.2, 0.1))
    assert np.allclose(models.Mutsel.stationary_distribution(pm), M.mutsel_stationary_distribution(pi, F, 0.2), rtol=1e-14)
    for d in (0.0, 1e-31, 0.5, -0.5, 60.0, -60.0):
        assert models.Mutsel.fixation_probability(d) == M.fixation_probability(d)
    w = cases.wag()
    with open(cases.GOLDEN + "/wag.dat") as fh:
        rm, fr = M.parse_wag(fh.read())
    assert np.array_equal(w.rate_matrix, rm) and np.array_equal(w.freqs, fr)
    with pytest.raises(RuntimeError):
        models.Wag.from_string_exn("1 2 3\n")
    for kind in ("mean", "median"):
        r1, w1 = models.gamma_category_rates(0.5, 4, kind)
        r2, w2 = M.gamma_category_rates(0.5, 4, kind)
        assert np.allclose(r1, r2, rtol=1e-13) and np.allclose(w1, w2) and abs(r1.mean() - 1.0) < 1e-12
    U, lam, Ui = models.Site_evolution_model.K80.rate_matrix_reduction(2.0)
    Uo, lo, Uio = M.k80_reduction(2.0)
    assert np.array_equal(U, Uo) and np.array_equal(lam, lo) and np.array_equal(Ui, Uio)
    U, lam, Ui = models.Site_evolution_model.JC69.rate_matrix_reduction()
    Uo, lo, Uio = M.jc69_reduction()
    assert np.array_equal(U, Uo) and np.array_equal(lam, lo) and np.array_equal(Ui, Uio)


def test_matrix_decomposition_reduce_matches_oracle():
    from phylogenetics_b200.phylo_ctmc import matrix_decomposition_reduce
    rng = np.random.default_rng(8)
    A, B, d = rng.random((4, 4)), rng.random((4, 4)), rng.random(4)
    dec = [("Mat", A), ("Diag", d), ("Transpose", B)]
    assert np.allclose(matrix_decomposition_reduce(4, dec), R.matrix_decomposition_reduce(4, dec), rtol=1e-15)
    assert np.allclose(matrix_decomposition_reduce(4, dec), A @ np.diag(d) @ B.T)
    assert np.array_equal(matrix_decomposition_reduce(3, []), np.eye(3))
    # applying the factors right to left to a vector == applying the reduced matrix
    v = rng.random(4)
    sv = R.sv_decomp_vec_mul(dec, (v, 0.5))
    assert np.allclose(sv[0], matrix_decomposition_reduce(4, dec) @ v) and sv[1] == 0.5


def test_birth_death_tree_is_ultrametric_and_binary():
    rng = np.random.default_rng(synthetic.BASE_SEED)
    t = synthetic.birth_death_tree(rng, 128)
    f = T.flatten(t)
    assert f.nleaves == 128 and f.nnodes == 255
    bl = f.branch_lengths(float)
    depth = np.zeros(f.nnodes)
    for v in range(f.nnodes):
        for c in f.children(v):
            depth[c] = depth[v] + bl[c]
    leaf_depth = depth[[v for v in range(f.nnodes) if f.is_leaf(v)]]
    assert np.allclose(leaf_depth, leaf_depth[0], rtol=1e-12) and leaf_depth[0] <= 1.0 + 1e-12
    assert all(len(f.children(v)) in (0, 2) for v in range(f.nnodes)) and (bl[1:] >= 0).all()
    with pytest.raises(ValueError):
        synthetic.birth_death_tree(rng, 10, birth_rate=1.0, death_rate=1.0)


def test_simulated_alignment_has_model_frequencies():
    case = cases.make_case("freq", "gtr", 12, 60000, ncat=1, seed=9)
    freq = np.bincount(case.tips.ravel(), minlength=4) / case.tips.size
    assert np.abs(freq - case.pi).max() < 0.01
    assert case.tips.dtype == np.uint8 and case.tips.shape == (12, 60000) and case.tips.max() <= 3


def test_string_alignment_encoding():
    from phylogenetics_b200.felsenstein import StringAlignment
    A = StringAlignment("ACGT")
    al = A.of_string_list(["ACGT", "ttga"])
    assert A.length(al) == 4 and A.get_base(al, seq="T1", pos=2) == 2
    assert A.encode(al, ["T1", "T0"]).tolist() == [[3, 3, 2, 0], [0, 1, 2, 3]]
    from phylogenetics_b200 import _lib
    with pytest.raises(_lib.PhyloB200InvalidArgument):
        A.encode(A.of_string_list(["ACGN"]), ["T0"])


def test_compress_site_patterns_host():
    from phylogenetics_b200.engine import compress_site_patterns
    tips = np.array([[0, 1, 0, 3, 1, 0], [2, 2, 2, 0, 2, 2]], dtype=np.uint8)
    patterns, counts, inverse = compress_site_patterns(tips)
    assert patterns.shape == (2, 3) and sorted(counts.tolist()) == [1.0, 2.0, 3.0]
    assert np.array_equal(patterns[:, inverse], tips)


def test_pack_tips2_layout():
    """Host layout of pb_set_tips_packed2: site 4b+j of a row in bits 2j..2j+1 of byte b, ragged tail zero padded."""
    from phylogenetics_b200.engine import pack_tips2
    rng = np.random.default_rng(3)
    for nsites in (1, 4, 7, 1003):
        tips = rng.integers(0, 4, size=(5, nsites), dtype=np.uint8)
        packed = pack_tips2(tips)
        assert packed.shape == (5, (nsites + 3) // 4) and packed.dtype == np.uint8
        back = np.stack([(packed >> (2 * j)) & 3 for j in range(4)], axis=2).reshape(5, -1)[:, :nsites]
        assert np.array_equal(back, tips)
    with pytest.raises(ValueError):
        pack_tips2(np.array([[0, 4]], dtype=np.uint8))
