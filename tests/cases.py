"""Shared builders of (tree, model, alignment) cases for the parity tests.

Inputs come from the product's host layer (phylogenetics_b200.synthetic / models); the
expected values come from the oracle (oracle/), never the other way round.
"""
import os

import numpy as np

from phylogenetics_b200 import models, synthetic, tree as T

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def wag():
    return models.Wag.from_file_exn(os.path.join(GOLDEN, "wag.dat"))


class Case:
    """A fully specified likelihood problem."""

    def __init__(self, name, flat, bl, Q, pi, tips, rates=(1.0,), weights=(1.0,), reversible=True):
        self.name, self.flat, self.bl, self.Q, self.pi, self.tips = name, flat, bl, Q, pi, tips
        self.rates, self.weights = np.asarray(rates, float), np.asarray(weights, float)
        self.reversible = reversible
        self.n = Q.shape[0]
        self.ncat = len(self.rates)
        self.nsites = tips.shape[1]

    def host_P(self):
        """[ncat][nnodes][n][n] by scipy's expm — an independent third route to P(t)."""
        import scipy.linalg as sl
        P = np.zeros((self.ncat, self.flat.nnodes, self.n, self.n))
        for k, r in enumerate(self.rates):
            for v in range(self.flat.nnodes):
                if v != self.flat.root:
                    P[k, v] = sl.expm(self.Q * (self.bl[v] * r))
        return P


def make_case(name, n_kind, ntaxa, nsites, ncat=1, alpha=0.5, seed=0, tree_kind="bd", site_seed=None):
    """`seed` fixes the tree and the model; `site_seed` (optional) draws the alignment columns from a separate
    stream, so that several shards of ONE problem (same tree and model, different columns) can be generated."""
    rng = np.random.default_rng(synthetic.BASE_SEED + seed)
    if tree_kind == "bd":
        t = synthetic.birth_death_tree(rng, ntaxa)
    else:
        t = synthetic.random_binary_tree(rng, ntaxa)
    flat = T.flatten(t)
    bl = flat.branch_lengths(float)
    if n_kind == "jc69":
        Q, pi = models.Rate_matrix.jc69(4), np.full(4, 0.25)
    elif n_kind == "k80":
        Q, pi = models.Rate_matrix.Nucleotide.k80(2.0), np.full(4, 0.25)
    elif n_kind == "gtr":
        Q, pi = synthetic.random_gtr(rng, 10.0)
    elif n_kind == "wag":
        w = wag()
        pi = w.freqs / w.freqs.sum()
        Q = models.Rate_matrix.scaled_rate_matrix(pi, models.Rate_matrix.make(
            20, lambda i, j: w.rate_matrix[i, j] * pi[j]))
    elif n_kind in ("mg94", "mutsel"):
        np_ = models.Nucleotide_process.random_gtr(rng, 10.0)
        nr, ns = models.Nucleotide_process.rate_matrix(np_), models.Nucleotide_process.stationary_distribution(np_)
        if n_kind == "mg94":
            p = dict(nucleotide_rates=nr, nucleotide_stat_dist=ns, omega=0.7)
            Q, pi = models.MG94.rate_matrix(p), models.MG94.stationary_distribution(p)
        else:
            p = models.Mutsel.random_param(rng, np_, 10.0)
            Q, pi = models.Mutsel.rate_matrix(p), models.Mutsel.stationary_distribution(p)
    else:
        raise ValueError(n_kind)
    rates, weights = models.gamma_category_rates(alpha, ncat)
    P = synthetic.host_transition_matrices(Q, pi, bl, rates)
    site_rng = rng if site_seed is None else np.random.default_rng([synthetic.BASE_SEED + seed, int(site_seed)])
    tips = synthetic.simulate_alignment(site_rng, flat, P, pi, nsites, weights)
    return Case(name, flat, bl, Q, pi, tips, rates, weights)
