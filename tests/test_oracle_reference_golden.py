"""The oracle PINNED against numbers the reference itself printed (its dune expect tests).

tests/expect/phylo_ctmc_substitution_mappings.expected:5 prints the pruning log-likelihood of one
WAG-simulated site on a 100-taxon birth-death tree: -117.326089.  Tree and site come from GSL's
default random stream after 20 000 path samples; oracle/gsl_replay.py replays that stream from
the published GSL algorithms, and reproduces along the way BOTH path-sampler histograms the test
prints and the Newick string of tests/expect/birth_death_simulator.expected — so the stream, the
tree and the site are the reference's, and the log-likelihood the oracle computes for them must
print as the reference's."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import gsl_replay as G
from oracle import linalg_ref as la
from oracle import phylo_ctmc_ref as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fixture():
    with open(os.path.join(GOLDEN, "reference_substitution_mappings.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def wag_text():
    with open(os.path.join(GOLDEN, "wag.dat")) as fh:
        return fh.read()


def test_birth_death_expect_newick(fixture):
    """tests/expect/birth_death_simulator.ml:26-30 / .expected:1."""
    tree = G.age_ntaxa_simulation(G.GslRng(), 2.0, 1.0, 5.0, 5)
    assert G.newick(tree) == fixture["birth_death_newick"]


def test_substitution_mappings_expect_replay(fixture, wag_text):
    """Histograms (.expected:2-3) and the pruning log-likelihood (.expected:5, "%f")."""
    h1, h2, tree, states, q, freqs = G.replay_substitution_mappings(wag_text)
    assert h1 == fixture["rejection_sampler_histogram"]
    assert h2 == fixture["uniformization_histogram"]
    ll = R.pruning(tree, 20, lambda bl: [("Mat", la.expm(bl * q))], lambda leaf: states[leaf], freqs)
    assert "%f" % ll == fixture["pruning_loglik_printed"] == "-117.326089"
    # the committed fixture is this very tree and site
    child_ptr, child_idx, tip_row, blen, tips = __import__("oracle.make_reference_fixture", fromlist=["flatten"]).flatten(tree, states)
    assert child_ptr == fixture["child_ptr"] and child_idx == fixture["child_idx"] and tips == fixture["tip_states"]
    assert [float(x).hex() for x in blen] == fixture["branch_len"]


def test_c_oracle_prints_the_reference_value(fixture, wag_text):
    """The C restatement (the checker of the GPU parity tests and the CPU baseline) on the fixture."""
    q, freqs = G.wag_rates(wag_text)
    bl = np.array([float.fromhex(x) for x in fixture["branch_len"]])
    P = np.array([la.expm(b * q) for b in bl])
    tips = np.zeros((len(fixture["tip_states"]), 1), dtype=np.uint8)
    tips[:, 0] = fixture["tip_states"]
    ll = C.pruning(20, fixture["root"], fixture["child_ptr"], fixture["child_idx"], fixture["tip_row"], P, freqs,
                   tips=tips)[0]
    assert "%f" % ll == "-117.326089"
