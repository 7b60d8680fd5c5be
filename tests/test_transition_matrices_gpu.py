"""GPU parity of the P(t) producers: K1 (eigensystem) and the batched Pade expm kernel against
the oracle's restatement of Matrix.expm / Make_diag."""
import numpy as np
import pytest

import cases
from oracle import c_oracle as C
from oracle import linalg_ref as la
from oracle import models_ref as M
from phylogenetics_b200 import engine, models
from phylogenetics_b200.engine import LikelihoodContext

pytestmark = pytest.mark.gpu


def test_make_diag_matrices_jc69_k80():
    """Site_evolution_model.JC69 / K80 (lib/site_evolution_model.ml:53-130) through kernel K1."""
    ts = np.array([0.0, 1e-5, 0.1, 0.5, 3.0, 50.0])
    for red in (M.jc69_reduction(), M.k80_reduction(2.0), M.k80_reduction(0.5)):
        got = engine.transition_matrices_eigen(*red, ts)
        for t, g in zip(ts, got):
            np.testing.assert_allclose(g, M.make_diag_tpm(red, t), rtol=0, atol=4e-16)
    # the host mirror's staged closures return the same matrices
    tpm = models.Site_evolution_model.K80.transition_probability_matrix(2.0)
    np.testing.assert_allclose(tpm(0.1), M.make_diag_tpm(M.k80_reduction(2.0), 0.1), atol=4e-16)
    # tests/test_site_evolution_model.ml:6-18 on the device: analytic vs numerical at 1e-4
    num = models.Site_evolution_model.K80_numerical.transition_probability_matrix(2.0)
    assert np.abs(tpm(0.1) - num(0.1)).max() <= 1e-4
    jc, jcn = models.Site_evolution_model.JC69, models.Site_evolution_model.JC69_numerical
    assert np.abs(jc.transition_probability_matrix()(0.1) - jcn.transition_probability_matrix()(0.1)).max() <= 1e-4


@pytest.mark.parametrize("kind,n", [("gtr", 4), ("wag", 20), ("mg94", 61), ("mutsel", 61)])
def test_expm_kernel_all_pade_orders(kind, n):
    """Every branch of Matrix.expm (lib/linear_algebra.ml:320-330,351-401): the four low-order Pade
    brackets, Pade-13 without scaling and with 1..8 squarings."""
    case = cases.make_case("e" + kind, kind, 4, 2, seed=1)
    Q = case.Q
    nq = la.norm1(Q)
    norms = [0.005, 0.1, 0.5, 1.5, 3.0, 5.0, 8.0, 40.0, 900.0]
    ts = np.array([x / nq for x in norms])
    got = engine.transition_matrices_expm(Q, ts)
    for t, g in zip(ts, got):
        ref = la.expm(t * Q)
        assert np.abs(g - ref).max() < 2e-13, (t * nq, np.abs(g - ref).max())
        assert np.abs(g.sum(axis=1) - 1.0).max() < 1e-12
    assert np.abs(models.Rate_matrix.transition_probability_matrix(0.3, Q) - la.expm(0.3 * Q)).max() < 2e-13


def test_expm_non_reversible_and_tiny():
    Qn = np.array([[-1., 1., 0., 0.], [0., -2., 2., 0.], [0., 0., -0.5, 0.5], [3., 0., 0., -3.]])
    got = engine.transition_matrices_expm(Qn, [0.01, 0.7, 6.0])
    for t, g in zip([0.01, 0.7, 6.0], got):
        assert np.abs(g - la.expm(t * Qn)).max() < 1e-13
    # n = 2 and n = 3 (padding to 4) and the zero matrix
    for n in (2, 3):
        Q = M.jc69(n)
        assert np.abs(engine.transition_matrices_expm(Q, [0.4])[0] - la.expm(0.4 * Q)).max() < 1e-14
    assert np.array_equal(engine.transition_matrices_expm(np.zeros((4, 4)), [1.0])[0], np.eye(4))


def test_general_rate_matrix_pipeline_matches_oracle():
    """pb_set_rate_matrix_general: Pade P(t) for every branch x category on the device, then pruning;
    Mutsel-style usage (lib/mutsel.ml:96-97) and a non-reversible nucleotide model."""
    for kind, ntaxa, nsites, ncat in [("mutsel", 8, 40, 1), ("gtr", 20, 500, 4)]:
        case = cases.make_case("gen" + kind, kind, ntaxa, nsites, ncat=ncat, seed=2)
        f = case.flat
        Pref = np.zeros((case.ncat, f.nnodes, case.n, case.n))
        for k, r in enumerate(case.rates):
            for v in range(f.nnodes):
                if v != f.root:
                    Pref[k, v] = la.expm((case.bl[v] * r) * case.Q)
        with LikelihoodContext(case.n, case.nsites, f.nleaves, ncat=case.ncat) as ctx:
            ctx.set_tree(f, case.bl)
            ctx.set_categories(case.rates, case.weights)
            ctx.set_rate_matrix_general(case.Q)
            ctx.set_tips(case.tips)
            ctx.set_root_frequencies(case.pi)
            total, per_site = ctx.loglik(per_site=True)
            Pdev = ctx.get_pmatrices()
        mask = np.arange(f.nnodes) != f.root
        assert np.abs(Pdev[:, mask] - Pref[:, mask]).max() < 2e-13
        expect = C.pruning(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, Pref, case.pi, tips=case.tips,
                           weights=case.weights)
        assert (np.abs(per_site - expect) / np.abs(expect)).max() < 1e-9
        assert abs(total - expect.sum()) < 1e-10 * abs(expect.sum())


def test_not_reversible_is_refused_by_the_eigen_path():
    from phylogenetics_b200 import _lib
    Qn = np.array([[-1., 1., 0., 0.], [0., -1., 1., 0.], [0., 0., -1., 1.], [1., 0., 0., -1.]])
    with LikelihoodContext(4, 10, 3) as ctx:
        with pytest.raises(_lib.PhyloB200Error) as e:
            ctx.set_rate_matrix(Qn, np.full(4, 0.25))
        assert e.value.code == _lib.PB_ERR_NUMERIC


@pytest.mark.gpu
def test_fp64_peak_probe():
    """pb_measure_fp64_peak: the DFMA and FP64 tensor-core peaks of the device are measured in the same range as the
    constants bench.py's roofline fractions use (34.8 / 36.8 TFLOP/s on a B200 at 1.965 GHz)."""
    from phylogenetics_b200.engine import measure_fp64_peak
    dfma, dmma = measure_fp64_peak(0)
    assert 12.0 < dfma < 40.0 and 12.0 < dmma < 40.0, (dfma, dmma)      # (wide: a power-capped box runs at lower clocks)
