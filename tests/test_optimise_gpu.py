"""BASELINE config 5 in miniature: Nelder_mead.minimize (mirror of lib/nelder_mead.ml) driving the GPU
evaluator as the opaque objective f = -log-likelihood(parameters).  One evaluation uploads O(branches)
scalars, rebuilds P(t) on the device (K1), prunes (K2), reduces (K3) and reads one double back."""
import math

import numpy as np
import pytest

import cases
from oracle import c_oracle as C
from oracle import linalg_ref as la
from phylogenetics_b200 import models, nelder_mead
from phylogenetics_b200.engine import LikelihoodContext

pytestmark = pytest.mark.gpu


def test_nelder_mead_recovers_kappa_and_tree_scale():
    true_kappa, true_scale = 3.0, 1.0
    rng = np.random.default_rng(5)
    case = cases.make_case("opt", "k80", 24, 40000, seed=12)      # simulated under kappa = 2
    # re-simulate under kappa = 3 so the optimum is not the generating default of make_case
    from phylogenetics_b200 import synthetic
    Q = models.Rate_matrix.Nucleotide.k80(true_kappa)
    P = synthetic.host_transition_matrices(Q, case.pi, case.bl)
    tips = synthetic.simulate_alignment(rng, case.flat, P, case.pi, case.nsites)
    f = case.flat
    evaluations = [0]
    with LikelihoodContext(4, case.nsites, f.nleaves) as ctx:
        ctx.set_tree(f, case.bl)
        ctx.set_tips(tips)
        ctx.set_root_frequencies(case.pi)

        def neg_loglik(x):
            kappa, scale = math.exp(x[0]), math.exp(x[1])
            evaluations[0] += 1
            ctx.set_eigensystem(*models.Site_evolution_model.K80.rate_matrix_reduction(kappa))
            ctx.set_branch_lengths(case.bl * scale)
            return -ctx.loglik()
        start = np.random.default_rng(1)
        best, x, its = nelder_mead.minimize(neg_loglik, lambda: [start.uniform(-1.0, 2.0), start.uniform(-1.0, 1.0)],
                                            tol=1e-7, maxit=200)
        at_truth = neg_loglik([math.log(true_kappa), math.log(true_scale)])
    kappa_hat, scale_hat = math.exp(x[0]), math.exp(x[1])
    assert best <= at_truth + 1e-6                       # the optimum is at least as good as the truth
    assert abs(kappa_hat - true_kappa) / true_kappa < 0.05
    assert abs(scale_hat - true_scale) < 0.05
    assert 10 < evaluations[0] < 600
    # the optimiser's value is a genuine log-likelihood: check it against the oracle at the optimum
    Qh = models.Rate_matrix.Nucleotide.k80(kappa_hat)
    Ph = np.array([la.expm((b * scale_hat) * Qh) for b in case.bl])[None]
    expect = C.pruning(4, f.root, f.child_ptr, f.child_idx, f.tip_row, Ph, case.pi, tips=tips)
    assert abs(-best - C.sum_exact(expect)) <= 1e-9 * abs(best)
