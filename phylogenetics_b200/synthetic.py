"""Synthetic benchmark inputs (host side, numpy): birth-death trees and alignments.

Distributionally equal to the reference's simulators, with our own seeded RNG
(PCG64), not GSL's stream:
  Birth_death.age_ntaxa_simulation   lib/birth_death.ml:35-86
  Simulator.site_exponential_method  lib/simulator.ml:35-40  (child ~ row of P(t);
      same distribution as the Gillespie direct method of :69-87, far cheaper)
"""
import numpy as np

from . import tree as T

BASE_SEED = 20240708


def birth_death_tree(rng, ntaxa, birth_rate=2.0, death_rate=1.0, age=1.0, rho=1.0):
    """Reconstructed birth-death tree conditioned on age and taxon count.

    Speciation times by inverse CDF (lib/birth_death.ml:68-83), then random
    coalescence of lineages backwards in time (sample_branch, :43-58).  Leaves carry
    their index, branches their length, the tree is binary and ultrametric.
    """
    b, d = float(birth_rate), float(death_rate)
    if b <= d:
        raise ValueError("expected birth_rate > death_rate")
    if ntaxa < 2:
        raise ValueError("not enough taxa")
    u = rng.random(ntaxa - 1)
    e = np.exp((d - b) * age)
    inner = 1.0 - u * (1.0 - ((b - d) * e) / (rho * b + (b * (1.0 - rho) - d) * e))
    times = age - (np.log(((b - d) / inner - (b * (1.0 - rho) - d)) / (rho * b)) + (d - b) * age) / (d - b)
    times.sort()
    particles = [(T.leaf(i), 0.0) for i in range(ntaxa)]
    for i in range(ntaxa, 1, -1):
        k = int(rng.integers(i))
        l = int(rng.integers(i - 1))
        l = l if l < k else l + 1
        k, l = min(k, l), max(k, l)
        t = float(times[ntaxa - i])
        left = T.branch(t - particles[k][1], particles[k][0])
        right = T.branch(t - particles[l][1], particles[l][0])
        particles[k] = (T.binary_node(None, left, right), t)
        particles[l] = particles[i - 1]
    return particles[0][0]


def random_binary_tree(rng, ntaxa, low=1e-5, high=0.5 + 1e-5):
    """Phylogenetic_tree.make_random-shaped tree (lib/phylogenetic_tree.ml:99-110) as a Tree.t:
    random joins, branch lengths U(0, 0.5) + 1e-5."""
    forest = [T.leaf(i) for i in range(ntaxa)]
    while len(forest) > 1:
        i, j = rng.choice(len(forest), size=2, replace=False)
        a, b = forest[i], forest[j]
        joined = T.binary_node(None, T.branch(float(rng.uniform(low, high)), a),
                               T.branch(float(rng.uniform(low, high)), b))
        forest = [x for q, x in enumerate(forest) if q not in (i, j)] + [joined]
    return forest[0]


def simulate_alignment(rng, flat, P, pi, nsites, weights=None):
    """Tip-state matrix uint8 [ntaxa][nsites]: root ~ pi, then each child ~ row of P(t).

    flat: tree.FlatTree; P: [ncat][nnodes][n][n] transition matrices (index = node at the
    lower end of the branch); sites are assigned a category ~ weights.
    """
    P = np.asarray(P)
    ncat, nnodes, n, _ = P.shape
    cat = None
    if ncat > 1:
        w = np.full(ncat, 1.0 / ncat) if weights is None else np.asarray(weights)
        cat = rng.choice(ncat, size=nsites, p=w / w.sum())
    cum_pi = np.cumsum(pi)
    cum_pi[-1] = 1.0
    states = {flat.root: np.searchsorted(cum_pi, rng.random(nsites), side="right").astype(np.uint8)}
    tips = np.zeros((flat.nleaves, nsites), dtype=np.uint8)
    for v in range(flat.nnodes):                       # pre-order ids: a parent precedes its children
        if flat.is_leaf(v):
            tips[flat.tip_row[v]] = states.pop(v)
            continue
        sv = states.pop(v)
        for c in flat.children(v):
            c = int(c)
            cum = np.cumsum(P[:, c], axis=2)           # [ncat][n][n]
            cum[:, :, -1] = 1.0
            rows = cum[cat, sv] if cat is not None else cum[0][sv]     # [nsites][n]
            u = rng.random(nsites)
            child = (u[:, None] >= rows).sum(axis=1)
            states[c] = np.minimum(child, n - 1).astype(np.uint8)
    return tips


def random_gtr(rng, alpha=10.0):
    """Nucleotide_process.Random.gtr (lib/nucleotide_process.ml:30-34): pi ~ Dirichlet(alpha),
    exchangeabilities ~ Gamma(1, 1).  Returns (Q, pi)."""
    from .models import Nucleotide_process
    p = Nucleotide_process.random_gtr(rng, alpha)
    return Nucleotide_process.rate_matrix(p), Nucleotide_process.stationary_distribution(p)


def host_transition_matrices(Q, pi, branch_lengths, rates=(1.0,)):
    """[ncat][nnodes][n][n] P(t) on the HOST, for input synthesis only (the product
    path builds its matrices on the GPU).  Reversible Q: symmetrise and numpy eigh."""
    Q = np.asarray(Q, dtype=np.float64)
    sq = np.sqrt(np.asarray(pi, dtype=np.float64))
    S = sq[:, None] * Q / sq[None, :]
    lam, V = np.linalg.eigh(0.5 * (S + S.T))
    U, Ui = V / sq[:, None], V.T * sq[None, :]
    t = np.asarray(rates, dtype=np.float64)[:, None] * np.asarray(branch_lengths, dtype=np.float64)[None, :]
    P = np.einsum("ik,cbk,kj->cbij", U, np.exp(t[:, :, None] * lam[None, None, :]), Ui)
    return np.clip(P, 0.0, None)
