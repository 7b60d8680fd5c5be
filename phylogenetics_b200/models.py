"""Host-side substitution-model builders: the reference modules that produce the
rate matrix Q and stationary distribution pi handed to the GPU path.

In the OCaml drop-in these stay on the host unchanged; this
module is their mirror for the Python host layer.  Module-like classes keep the
reference names:

  Rate_matrix            lib/rate_matrix.ml
  Nucleotide_process     lib/nucleotide_process.ml
  Wag                    lib/wag.ml
  Codon (NS, universal)  lib/codon.ml
  MG94                   lib/mG94.ml
  Mutsel                 lib/mutsel.ml
  Site_evolution_model   lib/site_evolution_model.ml   (P(t) itself is computed on the GPU)

State orderings: nucleotides A,C,G,T (lib/nucleotide.ml:3-6); amino acids
"ACDEFGHIKLMNPQRSTVWY" (lib/amino_acid.ml:4); sense codons in TCAG order
(lib/codon.ml:51-58).
"""
import math

import numpy as np



class _LazyEngine:
    """The Q / pi builders of this module are pure host code; only the P(t) producers need the native library.
    Importing it on first use keeps `models` importable by processes that must not load the product library
    (the CPU reference arm of bench.py generates its inputs with these builders)."""

    def __getattr__(self, name):
        from . import engine as _engine
        return getattr(_engine, name)


engine = _LazyEngine()

NUCLEOTIDES = "ACGT"
AMINO_ACIDS = "ACDEFGHIKLMNPQRSTVWY"


def _seq_sum(xs):
    acc = 0.0
    for x in xs:
        acc = acc + x
    return acc


class Rate_matrix:
    """lib/rate_matrix.ml (functor Make(A) applied to an alphabet of size n)."""

    @staticmethod
    def make(n, f):
        """:41-55 / :116-130 — failwith "Rates should be positive" on a negative rate."""
        r = np.zeros((n, n))
        for i in range(n):
            total = 0.0
            for j in range(n):
                if i == j:
                    continue
                x = float(f(i, j))
                if x < 0.0:
                    raise RuntimeError("Rates should be positive")
                total = x + total
                r[i, j] = x
            r[i, i] = -total
        return r

    @staticmethod
    def make_symetric(n, f):
        """:57-71 — f is called for i > j only (row-major order of calls); as in the
        reference, the stored diagonal only sums the entries already filled."""
        r = np.zeros((n, n))
        for i in range(n):
            for j in range(i):
                x = float(f(i, j))
                if x < 0.0:
                    raise RuntimeError("Rates should be positive")
                r[i, j] = r[j, i] = x
            r[i, i] = -_seq_sum(r[i, :])
        return r

    @staticmethod
    def jc69(n=4):
        """:37-39"""
        r = np.full((n, n), 1.0 / (float(n) - 1.0))
        np.fill_diagonal(r, -1.0)
        return r

    @staticmethod
    def scaled_rate_matrix(profile, rate):
        """:73-78 — divide by mu = -sum_i profile_i q_ii (expected rate 1)."""
        n = rate.shape[0]
        mu = -_seq_sum(profile[i] * rate[i, i] for i in range(n))
        return Rate_matrix.make(n, lambda i, j: rate[i, j] / mu)

    @staticmethod
    def scale(rate):
        """:80-82"""
        mu = _seq_sum(rate[i, i] for i in range(rate.shape[0]))
        return rate / mu

    @staticmethod
    def gtr(stationary_distribution, exchangeabilities):
        """:101-105"""
        pi = np.asarray(stationary_distribution, dtype=np.float64)
        ex = np.asarray(exchangeabilities, dtype=np.float64)
        m = Rate_matrix.make(len(pi), lambda i, j: ex[i, j] * pi[j])
        return Rate_matrix.scaled_rate_matrix(pi, m)

    @staticmethod
    def stationary_distribution(rate):
        """:33-35 -> Matrix.zero_eigen_vector (lib/linear_algebra.ml:266-275): least squares
        on [Q^T; 1^T] x = [0; 1].  Host-side, once per model."""
        n = rate.shape[0]
        a = np.vstack([np.asarray(rate).T, np.ones((1, n))])
        b = np.concatenate([np.zeros(n), [1.0]])
        return np.linalg.lstsq(a, b, rcond=None)[0]

    @staticmethod
    def transition_probability_matrix(tau, rates, device=0):
        """:132-138 — expm(tau * rates); float array array in, float array array out.
        Runs the batched Pade kernel on the GPU."""
        q = np.asarray(rates, dtype=np.float64)
        return engine.transition_matrices_expm(q, [tau], device=device)[0]

    class Nucleotide:
        @staticmethod
        def transversion(p, q):
            """lib/nucleotide.ml:15-19"""
            return {p, q} not in ({0, 2}, {1, 3})      # A<->G and C<->T are transitions

        @staticmethod
        def k80(kappa):
            """:143-148"""
            r = np.empty((4, 4))
            for i in range(4):
                for j in range(4):
                    if i == j:
                        r[i, j] = -1.0
                    elif Rate_matrix.Nucleotide.transversion(i, j):
                        r[i, j] = 1.0 / (kappa + 2.0)
                    else:
                        r[i, j] = kappa / (kappa + 2.0)
            return r

        @staticmethod
        def hky85(stationary_distribution, transition_rate, transversion_rate):
            """:150-162"""
            pi = np.asarray(stationary_distribution, dtype=np.float64)
            tv = Rate_matrix.Nucleotide.transversion
            m = Rate_matrix.make(4, lambda i, j: (transversion_rate if tv(i, j) else transition_rate) * pi[j])
            return Rate_matrix.scaled_rate_matrix(pi, m)


class Nucleotide_process:
    """lib/nucleotide_process.ml:3-28 — tagged tuples ("JC69",), ("K80", kappa),
    ("HKY85", pi, transition_rate, transversion_rate), ("GTR", pi, exchangeabilities)."""

    @staticmethod
    def rate_matrix(p):
        kind = p[0]
        if kind == "JC69":
            return Rate_matrix.jc69(4)
        if kind == "K80":
            return Rate_matrix.Nucleotide.k80(p[1])
        if kind == "HKY85":
            return Rate_matrix.Nucleotide.hky85(p[1], p[2], p[3])
        if kind == "GTR":
            return Rate_matrix.gtr(p[1], p[2])
        raise ValueError("unknown nucleotide process")

    @staticmethod
    def stationary_distribution(p):
        if p[0] in ("JC69", "K80"):
            return np.full(4, 0.25)
        return np.asarray(p[1], dtype=np.float64)

    @staticmethod
    def random_gtr(rng, alpha):
        """Random.gtr (:30-34): pi ~ Dirichlet(alpha), exchangeabilities ~ Gamma(1,1).
        rng is a numpy Generator (own stream, not GSL's)."""
        pi = rng.dirichlet(np.full(4, alpha))
        ex = Rate_matrix.make_symetric(4, lambda i, j: rng.gamma(1.0, 1.0))
        return ("GTR", pi, ex)


class Wag:
    """lib/wag.ml — PAML-format exchangeabilities + frequencies."""

    def __init__(self, rate_matrix, freqs):
        self.rate_matrix = rate_matrix
        self.freqs = freqs

    @staticmethod
    def from_string_exn(s):
        lines = [l for l in s.splitlines() if l]
        if len(lines) < 21:
            raise RuntimeError("Syntax error: only %d line(s)" % len(lines))
        floats = [np.array(l.split(), dtype=np.float64) for l in lines[:20]]
        letters = [ch for ch in lines[20] if "A" <= ch <= "Y"]
        try:
            where = np.array([letters.index(a) for a in AMINO_ACIDS])
        except ValueError:
            raise RuntimeError(lines[20])
        freqs = floats[19][where]
        lower = floats[:19]
        rm = Rate_matrix.make(20, lambda a, b: lower[max(where[a], where[b]) - 1][min(where[a], where[b])])
        return Wag(rm, freqs)

    @staticmethod
    def from_file_exn(fn):
        with open(fn) as fh:
            return Wag.from_string_exn(fh.read())


class Codon:
    """Universal genetic code, non-stop codons (lib/codon.ml:51-58,73,131-199)."""
    all_triplets = [x + y + z for x in "TCAG" for y in "TCAG" for z in "TCAG"]
    code = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
    ns_triplets = [t for t, a in zip(all_triplets, code) if a != "*"]
    ns_aa = np.array([AMINO_ACIDS.index(a) for a in code if a != "*"])
    ns_nuc = np.array([[NUCLEOTIDES.index(ch) for ch in t] for t in ns_triplets])
    card = len(ns_triplets)

    @staticmethod
    def neighbours_diff(p, q):
        """lib/codon.ml:34-40: Some (position, nuc_p, nuc_q) when exactly one position differs."""
        d = np.nonzero(Codon.ns_nuc[p] != Codon.ns_nuc[q])[0]
        if len(d) != 1:
            return None
        k = int(d[0])
        return k, int(Codon.ns_nuc[p, k]), int(Codon.ns_nuc[q, k])

    @staticmethod
    def synonym(p, q):
        return Codon.ns_aa[p] == Codon.ns_aa[q]

    @staticmethod
    def of_string(s):
        return Codon.ns_triplets.index(s)


class MG94:
    """lib/mG94.ml — param = dict(nucleotide_rates, nucleotide_stat_dist, omega)."""

    @staticmethod
    def rate_matrix(p):
        """:12-21"""
        nuc, omega = np.asarray(p["nucleotide_rates"]), p["omega"]

        def f(a, b):
            d = Codon.neighbours_diff(a, b)
            if d is None:
                return 0.0
            q = nuc[d[1], d[2]]
            return q if Codon.synonym(a, b) else omega * q
        return Rate_matrix.make(61, f)

    @staticmethod
    def stationary_distribution(p):
        """:23-35"""
        pi = np.asarray(p["nucleotide_stat_dist"], dtype=np.float64)
        a, g, t = 0, 2, 3
        pi_stop = pi[t] * pi[a] * pi[g] + pi[t] * pi[g] * pi[a] + pi[t] * pi[a] * pi[a]
        alpha = 1.0 / (1.0 - pi_stop)
        n = Codon.ns_nuc
        return pi[n[:, 0]] * pi[n[:, 1]] * pi[n[:, 2]] * alpha


class Mutsel:
    """lib/mutsel.ml — param = dict(nucleotide_rates, nucleotide_stat_dist, omega,
    scaled_fitness, gBGC, pps)."""

    @staticmethod
    def fitness_of_profile(profile, beta=1.0):
        return beta * np.log(np.asarray(profile, dtype=np.float64))

    @staticmethod
    def flat_fitness():
        return Mutsel.fitness_of_profile(np.full(20, 1.0 / 20.0))

    @staticmethod
    def flat_param():
        """:37-48"""
        pi = np.full(4, 0.25)
        ex = Rate_matrix.make_symetric(4, lambda i, j: 1.0 / 6.0)
        return dict(nucleotide_rates=Rate_matrix.gtr(pi, ex), nucleotide_stat_dist=pi, omega=1.0,
                    scaled_fitness=Mutsel.flat_fitness(), gBGC=0.0, pps=0.0)

    @staticmethod
    def random_param(rng, nucleotide_process, alpha):
        """:25-35"""
        return dict(nucleotide_rates=Nucleotide_process.rate_matrix(nucleotide_process),
                    nucleotide_stat_dist=Nucleotide_process.stationary_distribution(nucleotide_process),
                    omega=1.0, scaled_fitness=Mutsel.fitness_of_profile(rng.dirichlet(np.full(20, alpha))),
                    gBGC=0.0, pps=0.0)

    @staticmethod
    def fixation_probability(delta):
        """:51-56"""
        if abs(delta) < 1e-30:
            return 1.0 + delta / 2.0
        if delta > 50.0:
            return delta
        if delta < -50.0:
            return 0.0
        return delta / (1.0 - math.exp(-delta))

    @staticmethod
    def rate_matrix(p):
        """:58-80"""
        nuc = np.asarray(p["nucleotide_rates"])
        F, omega, gbgc, pps = p["scaled_fitness"], p["omega"], p["gBGC"], p["pps"]
        weak = (0, 3)          # A, T

        def f(a, b):
            d = Codon.neighbours_diff(a, b)
            if d is None:
                return 0.0
            _, xa, xb = d
            B = gbgc if (xa in weak and xb not in weak) else (-gbgc if (xa not in weak and xb in weak) else 0.0)
            if Codon.synonym(a, b):
                s = B + 0.0
            else:
                s = B + (F[Codon.ns_aa[b]] - F[Codon.ns_aa[a]] + pps)
            return nuc[xa, xb] * omega * Mutsel.fixation_probability(s)
        return Rate_matrix.make(61, f)

    @staticmethod
    def stationary_distribution(p):
        """:82-94"""
        pi = np.asarray(p["nucleotide_stat_dist"], dtype=np.float64)
        n = Codon.ns_nuc
        sign = np.where((n == 0) | (n == 3), -p["gBGC"], p["gBGC"]).sum(axis=1)
        w = pi[n[:, 0]] * pi[n[:, 1]] * pi[n[:, 2]] * np.exp(np.asarray(p["scaled_fitness"])[Codon.ns_aa] + sign / 2.0)
        return w / w.sum()

    @staticmethod
    def transition_probability_matrix(p, t, device=0):
        """:96-97 — expm(t * rate_matrix p), on the GPU."""
        return engine.transition_matrices_expm(Mutsel.rate_matrix(p), [t], device=device)[0]


class Site_evolution_model:
    """lib/site_evolution_model.ml: models as objects with rate_matrix(param),
    transition_probability_matrix(param)(t) [staged like :43-47], stationary_distribution(param)."""

    class _Diag:
        """Make_diag (:38-51): P(t) = U diag(exp(t lambda)) U^-1, built by kernel K1."""
        nstates = 4

        @classmethod
        def transition_probability_matrix(cls, param=None, device=0):
            U, lam, Uinv = cls.rate_matrix_reduction(param)
            return lambda t: engine.transition_matrices_eigen(U, lam, Uinv, [t], device=device)[0]

    class _Numerical:
        """Make (:27-35): P(t) = expm(t Q) (batched Pade kernel), pi by least squares."""
        nstates = 4

        @classmethod
        def transition_probability_matrix(cls, param=None, device=0):
            Q = cls.rate_matrix(param)
            return lambda t: engine.transition_matrices_expm(Q, [t], device=device)[0]

        @classmethod
        def stationary_distribution(cls, param=None):
            return Rate_matrix.stationary_distribution(cls.rate_matrix(param))

    class JC69(_Diag):
        """:53-86"""
        @staticmethod
        def rate_matrix(_=None):
            return Rate_matrix.jc69(4)

        @staticmethod
        def rate_matrix_reduction(_=None):
            U = np.array([[1., -1., -1., -1.], [1., 1., 0., 0.], [1., 0., 1., 0.], [1., 0., 0., 1.]])
            lam = np.array([0., -4. / 3., -4. / 3., -4. / 3.])
            Uinv = np.full((4, 4), -0.25) + np.eye(4)
            Uinv[0, :] = 0.25
            return U, lam, Uinv

        @staticmethod
        def stationary_distribution(_=None):
            return np.full(4, 0.25)

    class K80(_Diag):
        """:93-130"""
        @staticmethod
        def rate_matrix(kappa):
            return Rate_matrix.Nucleotide.k80(kappa)

        @staticmethod
        def rate_matrix_reduction(k):
            U = np.array([[1., -1., 0., -1.], [1., 1., -1., 0.], [1., -1., 0., 1.], [1., 1., 1., 0.]])
            l3 = (-2. * k - 2.) / (k + 2.)
            lam = np.array([0., -4. / (k + 2.), l3, l3])
            Uinv = np.array([[.25, .25, .25, .25], [-.25, .25, -.25, .25], [0., -.5, 0., .5], [-.5, 0., .5, 0.]])
            return U, lam, Uinv

        @staticmethod
        def stationary_distribution(_=None):
            return np.full(4, 0.25)


class _JC69_numerical(Site_evolution_model._Numerical):
    rate_matrix = staticmethod(Site_evolution_model.JC69.rate_matrix)
    rate_matrix_reduction = staticmethod(Site_evolution_model.JC69.rate_matrix_reduction)


class _K80_numerical(Site_evolution_model._Numerical):
    rate_matrix = staticmethod(Site_evolution_model.K80.rate_matrix)
    rate_matrix_reduction = staticmethod(Site_evolution_model.K80.rate_matrix_reduction)


Site_evolution_model.JC69_numerical = _JC69_numerical     # :88-91
Site_evolution_model.K80_numerical = _K80_numerical       # :132-135


def gamma_category_rates(alpha, ncat, kind="mean"):
    """Discrete-Gamma rate categories (Yang 1994) — an extension on top of the reference
    (which has no rate heterogeneity): Gamma(alpha, alpha) cut into ncat
    equiprobable classes, class rate = conditional mean (or renormalised median)."""
    from scipy.special import gammainc, gammaincinv
    if ncat == 1:
        return np.ones(1), np.ones(1)
    if kind == "mean":
        cuts = gammaincinv(alpha, np.arange(1, ncat) / ncat) / alpha
        cdf1 = np.concatenate([[0.0], gammainc(alpha + 1.0, cuts * alpha), [1.0]])
        rates = np.diff(cdf1) * ncat
    elif kind == "median":
        rates = gammaincinv(alpha, (np.arange(ncat) + 0.5) / ncat) / alpha
        rates = rates / rates.mean()
    else:
        raise ValueError("kind must be 'mean' or 'median'")
    return rates, np.full(ncat, 1.0 / ncat)
