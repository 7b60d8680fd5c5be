"""Replays the reference's GSL-RNG-driven expect tests so that the oracle can be pinned against the
numbers the reference itself printed (tests/expect/*.expected).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference draws everything from `Gsl.Rng.(make (default ()))`: GSL's default generator
(mt19937) with the default seed 0.  This module restates, from their published algorithms,
exactly the GSL primitives those tests consume — they are NOT vendored in the reference
(ocaml-gsl is an opam dependency, version unpinned; GSL >= 2.x assumed):

  gsl_rng_mt19937            Matsumoto-Nishimura 2002 initialisation; seed 0 is replaced by 4357
  gsl_rng_uniform            one 32-bit output / 2^32
  gsl_rng_uniform_int        rejection on  k = output / (0xffffffff / n)
  gsl_ran_exponential        -mu * log1p(-u)
  gsl_ran_discrete_preproc / gsl_ran_discrete   Walker's alias method with Knuth's construction
                             (two LIFO stacks of "smalls" and "bigs"), F'[k] = (k + F[k]) / K
  gsl_ran_poisson_pdf        exp(k log mu - ln k! - mu)

and the OCaml code that calls them, in the reference's evaluation order (ocamlopt evaluates
function arguments right to left: `List1.map` = `cons (f h) (List.map t ~f)` visits the tail of a
branch list BEFORE its head):

  Birth_death.age_ntaxa_simulation                       lib/birth_death.ml:35-86
  Simulator.branch_gillespie_direct                      lib/simulator.ml:69-79
  Simulator.site_gillespie_first_reaction                lib/simulator.ml:43-66
  Phylo_ctmc.conditional_simulation_along_branch_by_rejection_sampling   lib/phylo_ctmc.ml:494-529
  Phylo_ctmc.conditional_simulation_along_branch_by_uniformization        lib/phylo_ctmc.ml:417-474
  Phylo_ctmc.Uniformized_process.make                    lib/phylo_ctmc.ml:385-411
"""
import math

import numpy as np

from . import linalg_ref as la
from . import models_ref as M
from . import phylo_ctmc_ref as R


class GslRng:
    """gsl_rng_mt19937 with gsl_rng_default_seed = 0 (-> 4357)."""

    def __init__(self, seed=0):
        bg = np.random.MT19937()
        bg._legacy_seeding(4357 if seed == 0 else seed)       # init_genrand, the 2002 initialisation GSL uses
        self._bg = bg
        self._buf = np.empty(0, dtype=np.uint64)
        self._pos = 0
        self.draws = 0

    def get(self):
        if self._pos >= len(self._buf):
            self._buf = self._bg.random_raw(1 << 16)
            self._pos = 0
        v = int(self._buf[self._pos])
        self._pos += 1
        self.draws += 1
        return v

    def uniform(self):
        return self.get() / 4294967296.0

    def uniform_int(self, n):
        scale = 0xFFFFFFFF // n
        while True:
            k = self.get() // scale
            if k < n:
                return k

    def exponential(self, mu):
        return -mu * math.log1p(-self.uniform())


class Discrete:
    """gsl_ran_discrete_preproc (Walker alias tables by Knuth's two-stack construction)."""

    def __init__(self, probs):
        K = len(probs)
        total = 0.0
        for p in probs:
            total += p
        E = [p / total for p in probs]
        mean = 1.0 / K
        A = [0] * K
        F = [0.0] * K
        bigs, smalls = [], []
        for k in range(K):
            (smalls if E[k] < mean else bigs).append(k)
        while smalls:
            s = smalls.pop()
            if not bigs:
                A[s] = s
                F[s] = 1.0
                continue
            b = bigs.pop()
            A[s] = b
            F[s] = K * E[s]
            d = mean - E[s]
            E[s] += d
            E[b] -= d
            if E[b] < mean:
                smalls.append(b)
            elif E[b] > mean:
                bigs.append(b)
            else:
                A[b] = b
                F[b] = 1.0
        while bigs:
            b = bigs.pop()
            A[b] = b
            F[b] = 1.0
        for k in range(K):                 # !KNUTH_CONVENTION: F'[k] = (k + F[k]) / K
            F[k] += k
            F[k] /= K
        self.K, self.A, self.F = K, A, F

    def draw(self, rng):
        u = rng.uniform()
        c = int(u * self.K)
        f = self.F[c]
        if f == 1.0:
            return c
        return c if u < f else self.A[c]


def poisson_pdf(k, mu):
    return math.exp(math.log(mu) * k - math.lgamma(k + 1.0) - mu)


# ------------------------------------------------------------------ Birth_death
def age_ntaxa_simulation(rng, birth_rate, death_rate, age, ntaxa, rho=1.0):
    """lib/birth_death.ml:63-86 -> a phylo_ctmc_ref tree (leaf data = index, branch data = length)."""
    b, d = birth_rate, death_rate
    u = [rng.uniform() for _ in range(ntaxa - 1)]
    times = []
    for x in u:
        e = math.exp((d - b) * age)
        times.append(age - (math.log(((b - d) / (1.0 - x * (1.0 - ((b - d) * e) / (rho * b + (b * (1.0 - rho) - d) * e)))
                                      - (b * (1.0 - rho) - d)) / (rho * b)) + (d - b) * age) / (d - b))
    times.sort()
    n = ntaxa
    particles = [(R.Leaf(i), 0.0) for i in range(n)]
    i = n
    while i != 1:
        a = rng.uniform_int(i)
        c = rng.uniform_int(i - 1)
        c = c if c < a else c + 1
        k, l = (a, c) if a < c else (c, a)
        t = times[n - i]
        branch_k = R.Branch(t - particles[k][1], particles[k][0])
        branch_l = R.Branch(t - particles[l][1], particles[l][0])
        particles[k] = (R.binary_node(None, branch_k, branch_l), t)
        particles[l] = particles[i - 1]
        i -= 1
    return particles[0][0]


def newick(t):
    """tests/expect/birth_death_simulator.ml:4-24 (Newick.to_string: %.12f-style lengths as printed there)."""
    def node(t, parent=None):
        lab = ("n%d" % t.data) if isinstance(t, R.Leaf) else \
            "(" + ",".join(node(b.tip, b.data) for b in t.branches) + ")"
        return lab + ((":%.12f" % parent) if parent is not None else "")
    return node(t) + ";"


# -------------------------------------------------------------------- Simulator
class AliasCache:
    """symbol_sample's tables depend only on the current state: build each once."""

    def __init__(self, rate_matrix):
        self.q = rate_matrix
        self.n = rate_matrix.shape[0]
        self._t = {}

    def table(self, state):
        if state not in self._t:
            self._t[state] = Discrete([0.0 if m == state else self.q[state, m] for m in range(self.n)])
        return self._t[state]


def branch_gillespie_direct(rng, cache, start_state, branch_length):
    """lib/simulator.ml:69-79: list of (state, time) events, newest first; plus the end state."""
    state, t, acc = start_state, 0.0, []
    while True:
        total_rate = -cache.q[state, state]
        t2 = t + rng.exponential(1.0 / total_rate)
        if t2 > branch_length:
            return acc, state
        state = cache.table(state).draw(rng)
        t = t2
        acc.append((state, t2))


def rejection_sampling_path(rng, cache, branch_length, start_state, end_state, max_tries):
    """lib/phylo_ctmc.ml:494-529: events of the first simulated path that ends in end_state."""
    for _ in range(max_tries):
        events, end = branch_gillespie_direct(rng, cache, start_state, branch_length)
        if end == end_state:
            return events
    raise RuntimeError("Reached max tries")


def site_gillespie_first_reaction(rng, tree, root, q):
    """lib/simulator.ml:43-66 through Simulator.evolve / Tree.propagate (lib/tree.ml:119-132).
    Returns {leaf data: state}.  Branches of a node are visited tail first (see module docstring)."""
    n = q.shape[0]
    out = {}

    def evolve_branch(x_in, length):
        state, remaining = x_in, length
        while True:
            best = None
            for m in range(n):
                if m == state:
                    cand = (math.inf, m)
                else:
                    rate = q[state, m]
                    cand = (math.inf, m) if rate < 1e-30 else (rng.exponential(1.0 / rate), m)
                if best is None or cand < best:         # Array.min_elt ~compare:Poly.compare: first minimum
                    best = cand
            tau, nxt = best
            if tau > remaining:
                return state
            state, remaining = nxt, remaining - tau

    def inner(state, t):
        if isinstance(t, R.Leaf):
            out[t.data] = state
            return
        order = [t.branches[0]] + list(t.branches[1:])
        for b in list(reversed(order[1:])) + [order[0]]:     # List.map over the tail (right to left), then the head
            inner(evolve_branch(state, b.data), b.tip)
    # NB: Base.List.map itself maps left to right over the tail; with two branches the tail has one element.
    inner(root, tree)
    return out


# --------------------------------------------------------- Uniformized process
class UniformizedProcess:
    """lib/phylo_ctmc.ml:369-414"""

    def __init__(self, q, branch_length):
        self.q = q
        n = q.shape[0]
        mu = -q[0, 0]
        for i in range(1, n):
            mu = max(mu, -q[i, i])
        self.mu = mu
        self.R1 = q / mu + np.eye(n)
        self._pow = {}
        self.P = la.expm(branch_length * q)
        self.branch_length = branch_length
        self.n = n

    def R(self, k):
        if k == 0:
            return np.eye(self.n)
        if k == 1:
            return self.R1
        if k not in self._pow:
            self._pow[k] = self.R(k - 1) @ self.R1
        return self._pow[k]


def uniformization_path(rng, up, max_path_length, start_state, end_state):
    """lib/phylo_ctmc.ml:417-474; returns the collapsed list of (state, time) events."""
    lam = up.branch_length
    p_ab = up.P[start_state, end_state]

    def p_n(k):
        return poisson_pdf(k, up.mu * lam)

    def p_b_given_n(k):
        if k == 0:
            return 1.0 if start_state == end_state else 0.0
        return up.R(k)[start_state, end_state]
    g = p_ab * rng.uniform()
    acc, n = 0.0, None
    for i in range(max_path_length + 1):
        acc2 = acc + p_n(i) * p_b_given_n(i)
        if acc2 > g:
            n = i
            break
        acc = acc2
    if n is None:
        n = Discrete([p_n(i) * p_b_given_n(i) for i in range(max_path_length + 1)]).draw(rng)
    path = []
    cur = start_state
    for i in range(n):
        second = up.R(n - i - 1)
        w = [up.R1[cur, s] * second[s, end_state] for s in range(up.n)]
        cur = Discrete(w).draw(rng)
        path.append(cur)
    times = sorted(rng.uniform() * lam for _ in range(n))
    events, prec = [], start_state
    for s, t in zip(path, times):                    # collapse_mapping :403-415
        if s != prec:
            events.append((s, t))
        prec = s
    return events


def histogram_line(counts_list):
    """render_int_histogram, tests/expect/phylo_ctmc_substitution_mappings.ml:53-67."""
    n = len(counts_list)
    vals, cnt = np.unique(np.asarray(counts_list), return_counts=True)
    return " | ".join("%02d %.3f" % (int(k), c / n) for k, c in zip(vals, cnt))


def wag_rates(wag_text):
    """transition_rates wag_param, tests/expect/phylo_ctmc_substitution_mappings.ml:24-50 (scale = 1)."""
    exch, freqs = M.parse_wag(wag_text)
    return M.make(20, lambda i, j: 1.0 * exch[i, j] * freqs[j]), freqs


def replay_substitution_mappings(wag_text, sample_size=10_000):
    """Replays tests/expect/phylo_ctmc_substitution_mappings.ml up to the pruning log-likelihood.

    Returns (histogram line of the rejection sampler, histogram line of the uniformization sampler,
    tree, leaf states, rate matrix, root frequencies).
    """
    rng = GslRng()
    q, freqs = wag_rates(wag_text)
    alanine, valine = M.AMINO_ACIDS.index("A"), M.AMINO_ACIDS.index("V")
    up = UniformizedProcess(q, 2.0)
    cache = AliasCache(q)
    rej = [len(rejection_sampling_path(rng, cache, 2.0, alanine, valine, 10_000)) for _ in range(sample_size)]
    uni = [len(uniformization_path(rng, up, 20, alanine, valine)) for _ in range(sample_size)]
    tree = age_ntaxa_simulation(rng, 2.0, 1.0, 1.0, 100)
    states = site_gillespie_first_reaction(rng, tree, valine, q)
    return histogram_line(rej), histogram_line(uni), tree, states, q, freqs
