"""Oracle restatement of the reference's substitution-model builders (Q, pi, P(t)).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

State orderings (part of the data contract):
  nucleotides A,C,G,T = 0..3                     lib/nucleotide.ml:3-6
  amino acids "ACDEFGHIKLMNPQRSTVWY"              lib/amino_acid.ml:4
  codons in TCAG triplet order, stops removed     lib/codon.ml:51-58,143-146
"""
import math

import numpy as np

from . import linalg_ref as la

NUCLEOTIDES = "ACGT"
AMINO_ACIDS = "ACDEFGHIKLMNPQRSTVWY"

# ---------------------------------------------------------------- Rate_matrix


def make(n, f):
    """Rate_matrix.make, lib/rate_matrix.ml:41-55 / :116-130.

    Off-diagonals f(i, j) >= 0, diagonal = -(row sum accumulated in j order).
    """
    r = np.zeros((n, n))
    for i in range(n):
        total = 0.0
        for j in range(n):
            if i != j:
                rij = f(i, j)
                if rij < 0.0:
                    raise RuntimeError("Rates should be positive")   # failwith
                total = rij + total
                r[i, j] = rij
        r[i, i] = -total
    return r


def make_symetric(n, f):
    """Rate_matrix.make_symetric, lib/rate_matrix.ml:57-71 (f called for i > j)."""
    # Quirk kept on purpose: the diagonal of row i is written right after row i's
    # lower part, when the entries r[i, j>i] are still zero, so it only sums the
    # lower triangle.  Harmless: callers use the result as exchangeabilities and
    # never read its diagonal (Rate_matrix.gtr goes through `make`).
    r = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i > j:
                rij = f(i, j)
                if rij < 0.0:
                    raise RuntimeError("Rates should be positive")
                r[i, j] = rij
                r[j, i] = rij
        total = 0.0
        for j in range(n):
            total = total + r[i, j]
        r[i, i] = -total
    return r


def jc69(n=4):
    """Rate_matrix.jc69, lib/rate_matrix.ml:37-39."""
    r = 1.0 / (float(n) - 1.0)
    return np.array([[-1.0 if i == j else r for j in range(n)] for i in range(n)])


def scaled_rate_matrix(profile, rate):
    """Rate_matrix.scaled_rate_matrix, lib/rate_matrix.ml:73-78."""
    n = rate.shape[0]
    acc = 0.0
    for i in range(n):
        acc = acc + profile[i] * rate[i, i]
    mu = -acc
    return make(n, lambda i, j: rate[i, j] / mu)


def scale(rate):
    """Rate_matrix.scale, lib/rate_matrix.ml:80-82 (divides by the trace)."""
    n = rate.shape[0]
    mu = 0.0
    for i in range(n):
        mu = mu + rate[i, i]
    return rate / mu


def gtr(stationary_distribution, exchangeabilities):
    """Rate_matrix.gtr, lib/rate_matrix.ml:101-105."""
    n = len(stationary_distribution)
    m = make(n, lambda i, j: exchangeabilities[i, j] * stationary_distribution[j])
    return scaled_rate_matrix(stationary_distribution, m)


def stationary_distribution(rate):
    """Rate_matrix.stationary_distribution, lib/rate_matrix.ml:33-35."""
    return la.zero_eigen_vector(rate)


def transversion(p, q):
    """Nucleotide.transversion, lib/nucleotide.ml:15-19 (A<->G, C<->T are transitions)."""
    return (p, q) not in ((0, 2), (2, 0), (1, 3), (3, 1))


def k80(kappa):
    """Rate_matrix.Nucleotide.k80, lib/rate_matrix.ml:143-148."""
    r = np.zeros((4, 4))
    for i in range(4):
        for j in range(4):
            if i == j:
                r[i, j] = -1.0
            elif transversion(i, j):
                r[i, j] = 1.0 / (kappa + 2.0)
            else:
                r[i, j] = kappa / (kappa + 2.0)
    return r


def hky85(stationary_distribution, transition_rate, transversion_rate):
    """Rate_matrix.Nucleotide.hky85, lib/rate_matrix.ml:150-162."""
    def f(i, j):
        coef = transversion_rate if transversion(i, j) else transition_rate
        return coef * stationary_distribution[j]
    return scaled_rate_matrix(stationary_distribution, make(4, f))


def transition_probability_matrix(tau, rates):
    """Rate_matrix.transition_probability_matrix, lib/rate_matrix.ml:132-138."""
    return la.expm(tau * np.asarray(rates, dtype=np.float64))


# ------------------------------------------------------- Site_evolution_model


def jc69_reduction():
    """Site_evolution_model.JC69.Base.rate_matrix_reduction, lib/site_evolution_model.ml:57-73."""
    p = np.array([[1., -1., -1., -1.],
                  [1., 1., 0., 0.],
                  [1., 0., 1., 0.],
                  [1., 0., 0., 1.]])
    d = np.array([0., -4. / 3., -4. / 3., -4. / 3.])
    pinv = np.array([[0.25, 0.25, 0.25, 0.25],
                     [-0.25, 0.75, -0.25, -0.25],
                     [-0.25, -0.25, 0.75, -0.25],
                     [-0.25, -0.25, -0.25, 0.75]])
    return p, d, pinv


def k80_reduction(k):
    """Site_evolution_model.K80.Base.rate_matrix_reduction, lib/site_evolution_model.ml:97-117."""
    p = np.array([[1., -1., 0., -1.],
                  [1., 1., -1., 0.],
                  [1., -1., 0., 1.],
                  [1., 1., 1., 0.]])
    l3 = (-2.0 * k - 2.0) / (k + 2.0)
    d = np.array([0.0, -4.0 / (k + 2.0), l3, l3])
    pinv = np.array([[0.25, 0.25, 0.25, 0.25],
                     [-0.25, 0.25, -0.25, 0.25],
                     [0., -0.5, 0., 0.5],
                     [-0.5, 0., 0.5, 0.]])
    return p, d, pinv


def make_diag_tpm(reduction, t):
    """Site_evolution_model.Make_diag.transition_probability_matrix, :43-47.

    (P . diagm(exp(t*lambda))) . P^-1, two dgemm in that association.
    """
    p, d, pinv = reduction
    return (p @ np.diag(np.exp(t * d))) @ pinv


def make_tpm(rate_matrix, t):
    """Site_evolution_model.Make.transition_probability_matrix, :31-32 (Pade expm)."""
    return la.expm(t * rate_matrix)


# --------------------------------------------------------------------- WAG


def parse_wag(text):
    """Wag.from_string_exn, lib/wag.ml:8-50.

    PAML layout: 19 lower-triangular rows, a frequency line, an amino-acid
    order line; both are remapped from the file's order to AMINO_ACIDS order.
    Returns (rate_matrix built by Rate_matrix.make from the symmetric
    exchangeabilities, freqs).
    """
    lines = [l for l in text.splitlines() if l != ""]
    if len(lines) < 21:
        raise RuntimeError("Syntax error: only %d line(s)" % len(lines))
    rows = [[float(x) for x in l.split(" ") if x != ""] for l in lines[:19]]
    freq_vals = [float(x) for x in lines[19].split(" ") if x != ""]
    chars = "".join(ch for ch in lines[20] if "A" <= ch <= "Y")
    order = [chars.index(aa) for aa in AMINO_ACIDS]
    freqs = np.array([freq_vals[order[a]] for a in range(20)])

    def f(a1, a2):
        i, j = order[a1], order[a2]
        return rows[max(i, j) - 1][min(i, j)]
    return make(20, f), freqs


# -------------------------------------------------------------------- Codon

ALL_TRIPLETS = [a + b + c for a in "TCAG" for b in "TCAG" for c in "TCAG"]  # lib/codon.ml:51-58
STANDARD_CODE = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"  # lib/codon.ml:73
NS_TRIPLETS = [t for t, a in zip(ALL_TRIPLETS, STANDARD_CODE) if a != "*"]  # lib/codon.ml:143
NS_AA = [AMINO_ACIDS.index(a) for a in STANDARD_CODE if a != "*"]           # lib/codon.ml:144-145


def ns_neighbours_diff(p, q):
    """Codon.Impl.neighbours_diff on NS codons, lib/codon.ml:34-40."""
    ps, qs = NS_TRIPLETS[p], NS_TRIPLETS[q]
    same = [ps[k] == qs[k] for k in range(3)]
    if same.count(False) != 1:
        return None
    pos = same.index(False)
    return pos, NUCLEOTIDES.index(ps[pos]), NUCLEOTIDES.index(qs[pos])


def ns_synonym(p, q):
    return NS_AA[p] == NS_AA[q]


def ns_nucleotides(c):
    return tuple(NUCLEOTIDES.index(ch) for ch in NS_TRIPLETS[c])


# --------------------------------------------------------------------- MG94


def mg94_rate_matrix(nucleotide_rates, omega):
    """MG94.rate_matrix, lib/mG94.ml:12-21."""
    def f(p, q):
        d = ns_neighbours_diff(p, q)
        if d is None:
            return 0.0
        _, xa, xb = d
        qab = nucleotide_rates[xa, xb]
        return qab if ns_synonym(p, q) else omega * qab
    return make(61, f)


def mg94_stationary_distribution(pi):
    """MG94.stationary_distribution, lib/mG94.ml:23-35 (stops TAG, TGA, TAA)."""
    a, g, t = 0, 2, 3
    pi_stop = pi[t] * pi[a] * pi[g] + pi[t] * pi[g] * pi[a] + pi[t] * pi[a] * pi[a]
    alpha = 1.0 / (1.0 - pi_stop)
    out = np.zeros(61)
    for c in range(61):
        i, j, k = ns_nucleotides(c)
        out[c] = pi[i] * pi[j] * pi[k] * alpha
    return out


# ------------------------------------------------------------------- Mutsel


def fixation_probability(delta):
    """Mutsel.fixation_probability, lib/mutsel.ml:51-56."""
    if abs(delta) < 1e-30:
        return 1.0 + delta / 2.0
    if delta > 50.0:
        return delta
    if delta < -50.0:
        return 0.0
    return delta / (1.0 - math.exp(-delta))


def mutsel_rate_matrix(nucleotide_rates, omega, scaled_fitness, gBGC=0.0, pps=0.0):
    """Mutsel.rate_matrix, lib/mutsel.ml:58-80."""
    at = (0, 3)

    def f(p, q):
        d = ns_neighbours_diff(p, q)
        if d is None:
            return 0.0
        _, xa, xb = d
        if xa in at and xb not in at:
            b = gBGC
        elif xa not in at and xb in at:
            b = -gBGC
        else:
            b = 0.0
        if ns_synonym(p, q):
            sel = b + 0.0
        else:
            sel = b + (scaled_fitness[NS_AA[q]] - scaled_fitness[NS_AA[p]] + pps)
        return nucleotide_rates[xa, xb] * omega * fixation_probability(sel)
    return make(61, f)


def mutsel_stationary_distribution(pi, scaled_fitness, gBGC=0.0):
    """Mutsel.stationary_distribution, lib/mutsel.ml:82-94."""
    out = np.zeros(61)
    for c in range(61):
        n1, n2, n3 = ns_nucleotides(c)
        bsum = sum((-gBGC if n in (0, 3) else gBGC) for n in (n1, n2, n3))
        out[c] = pi[n1] * pi[n2] * pi[n3] * math.exp(scaled_fitness[NS_AA[c]] + bsum / 2.0)
    return out / out.sum()


def fitness_of_profile(profile, beta=1.0):
    """Mutsel.fitness_of_profile, lib/mutsel.ml:18-19."""
    return beta * np.log(np.asarray(profile, dtype=np.float64))


# ------------------------------------------------- +Gamma (extension)


def gamma_category_rates(alpha, ncat, kind="mean"):
    """Discrete-Gamma rates (Yang 1994), NOT in the reference.

    Gamma(shape=alpha, rate=alpha) cut into ncat equiprobable classes; the class
    rate is its conditional mean ("mean") or its median ("median", renormalised
    to mean 1).  Weights are 1/ncat.
    """
    from scipy.special import gammainc, gammaincinv
    if ncat == 1:
        return np.ones(1), np.ones(1)
    k = np.arange(ncat + 1) / ncat
    if kind == "mean":
        cuts = gammaincinv(alpha, k[1:-1]) / alpha          # quantile boundaries
        upper = np.concatenate([gammainc(alpha + 1.0, cuts * alpha), [1.0]])
        lower = np.concatenate([[0.0], upper[:-1]])
        rates = (upper - lower) * ncat
    else:
        mid = (np.arange(ncat) + 0.5) / ncat
        rates = gammaincinv(alpha, mid) / alpha
        rates = rates / rates.mean()
    return rates, np.full(ncat, 1.0 / ncat)
