"""Writes tests/golden/reference_substitution_mappings.json: the 100-taxon tree and the WAG-simulated
site of the reference's expect test tests/expect/phylo_ctmc_substitution_mappings.ml, obtained by
replaying the reference's GSL random stream (oracle/gsl_replay.py), together with the numbers the
reference printed for it (tests/expect/phylo_ctmc_substitution_mappings.expected).

TEST INFRASTRUCTURE ONLY.  Run from the repository root:  python -m oracle.make_reference_fixture
"""
import json
import os

from . import gsl_replay as G
from . import phylo_ctmc_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EXPECTED = {
    "source": "tests/expect/phylo_ctmc_substitution_mappings.expected",
    "rejection_sampler_histogram": "01 0.488 | 02 0.150 | 03 0.221 | 04 0.081 | 05 0.039 | 06 0.015 | 07 0.004 | 08 0.001 | 09 0.001 | 10 0.000",
    "uniformization_histogram": "01 0.481 | 02 0.149 | 03 0.227 | 04 0.087 | 05 0.038 | 06 0.013 | 07 0.005 | 08 0.001 | 09 0.000",
    "pruning_loglik_printed": "-117.326089",
    "birth_death_newick_source": "tests/expect/birth_death_simulator.expected",
    "birth_death_newick": "(n0:4.926876639043,((n1:0.092501292168,n2:0.092501292168):2.151229766277,(n4:0.179025801152,n3:0.179025801152):2.064705257293):2.683145580598);",
}


def flatten(tree, states):
    """Pre-order CSR of a phylo_ctmc_ref tree (children in branch order)."""
    child_ptr, child_idx, tip_row, blen, tips = [0], [], [], [], []
    order = []

    def visit(t, bl):
        v = len(order)
        order.append((t, bl))
        return v
    # iterative pre-order keeping branch order
    nodes = []
    stack = [(tree, 0.0, None)]
    ids = {}
    while stack:
        t, bl, _ = stack.pop()
        ids[id(t)] = len(nodes)
        nodes.append((t, bl))
        if isinstance(t, R.Node):
            for b in reversed(t.branches):
                stack.append((b.tip, b.data, None))
    for t, bl in nodes:
        blen.append(bl)
        if isinstance(t, R.Leaf):
            tip_row.append(len(tips))
            tips.append(int(states[t.data]))
        else:
            tip_row.append(-1)
            child_idx.extend(ids[id(b.tip)] for b in t.branches)
        child_ptr.append(len(child_idx))
    return child_ptr, child_idx, tip_row, blen, tips


def main():
    with open(os.path.join(ROOT, "tests", "golden", "wag.dat")) as fh:
        wag = fh.read()
    h1, h2, tree, states, q, freqs = G.replay_substitution_mappings(wag)
    assert h1 == EXPECTED["rejection_sampler_histogram"], h1
    assert h2 == EXPECTED["uniformization_histogram"], h2
    child_ptr, child_idx, tip_row, blen, tips = flatten(tree, states)
    doc = dict(EXPECTED)
    doc.update({
        "nstates": 20, "root": 0, "child_ptr": child_ptr, "child_idx": child_idx, "tip_row": tip_row,
        "branch_len": [float(x).hex() for x in blen], "tip_states": tips,
        "model": "WAG: q_ij = exchangeability_ij * freq_j (unscaled, scale = 1), root frequencies = WAG freqs; "
                 "P(t) = Matrix.expm(t Q); tests/golden/wag.dat",
    })
    out = os.path.join(ROOT, "tests", "golden", "reference_substitution_mappings.json")
    with open(out, "w") as fh:
        json.dump(doc, fh, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
