"""The CUDA path on the reference's own golden vector: the 100-taxon tree and WAG-simulated site of
tests/expect/phylo_ctmc_substitution_mappings.ml (committed fixture, produced by replaying the
reference's GSL stream: oracle/make_reference_fixture.py) must print -117.326089 like
tests/expect/phylo_ctmc_substitution_mappings.expected:5."""
import json
import os

import numpy as np
import pytest

from oracle import gsl_replay as G
from oracle import linalg_ref as la
from phylogenetics_b200 import phylo_ctmc, tree as T
from phylogenetics_b200.engine import LikelihoodContext

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Flat:
    pass


@pytest.fixture(scope="module")
def problem():
    with open(os.path.join(GOLDEN, "reference_substitution_mappings.json")) as fh:
        fx = json.load(fh)
    with open(os.path.join(GOLDEN, "wag.dat")) as fh:
        q, freqs = G.wag_rates(fh.read())
    f = Flat()
    f.root, f.child_ptr, f.child_idx, f.tip_row = fx["root"], np.array(fx["child_ptr"], np.int32), \
        np.array(fx["child_idx"], np.int32), np.array(fx["tip_row"], np.int32)
    f.nnodes = len(f.tip_row)
    bl = np.array([float.fromhex(x) for x in fx["branch_len"]])
    tips = np.array(fx["tip_states"], dtype=np.uint8)[:, None]
    return fx, f, bl, q, freqs, tips


@pytest.mark.parametrize("source", ["expm_kernel", "host_matrices", "eigensystem"])
@pytest.mark.parametrize("opts", [{}, {"fused_dmma": 0}, {"fused_dmma": 0, "dmma_version": 1}, {"dense_scalar": 1}])
def test_gpu_prints_the_reference_value(problem, source, opts):
    fx, f, bl, q, freqs, tips = problem
    with LikelihoodContext(20, 1, len(fx["tip_states"])) as ctx:
        ctx.set_tree(f, bl)
        for k, v in opts.items():
            ctx.set_option(k, v)
        if source == "expm_kernel":                 # Matrix.expm per branch on the device, like the reference's closure
            ctx.set_rate_matrix_general(q)
        elif source == "eigensystem":               # WAG is reversible w.r.t. its (unnormalised) frequencies
            ctx.set_rate_matrix(q, freqs / freqs.sum())
        else:
            ctx.set_pmatrices(np.array([la.expm(b * q) for b in bl])[None])
        ctx.set_tips(tips)
        ctx.set_root_frequencies(freqs)
        total, per_site = ctx.loglik(per_site=True)
    assert "%f" % total == fx["pruning_loglik_printed"] == "-117.326089"
    assert "%f" % per_site[0] == "-117.326089"


def test_mirror_api_on_the_reference_tree(problem):
    """Through the Phylo_ctmc.pruning mirror with the reference test's closures
    (tests/expect/phylo_ctmc_substitution_mappings.ml:170-182)."""
    fx, f, bl, q, freqs, tips = problem
    f.leaf_data = list(range(len(fx["tip_states"])))
    f.node_data = [None] * f.nnodes
    f.branch_data = [None if v == f.root else float(bl[v]) for v in range(f.nnodes)]
    f.is_leaf = lambda v: f.child_ptr[v] == f.child_ptr[v + 1]
    f.children = lambda v: f.child_idx[f.child_ptr[v]:f.child_ptr[v + 1]]
    tree = T.unflatten(f)
    transition_probabilities = lambda l: [("Mat", la.expm(l * q))]
    ll = phylo_ctmc.pruning(tree, 20, transition_probabilities, lambda row: int(tips[row, 0]), freqs)
    assert "%f" % ll == "-117.326089"
