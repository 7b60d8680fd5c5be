import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from phylogenetics_b200.engine import LikelihoodContext

def run(case, P, opts):
    with LikelihoodContext(case.n, case.nsites, case.flat.nleaves, ncat=case.ncat) as ctx:
        ctx.set_tree(case.flat, case.bl); ctx.set_categories(case.rates, case.weights)
        ctx.set_pmatrices(P); ctx.set_tips(case.tips); ctx.set_root_frequencies(case.pi)
        for k, v in opts.items(): ctx.set_option(k, v)
        return ctx.loglik(per_site=True)[1]

for (model, ntaxa, nsites, ncat, kind) in [("wag", 150, 333, 2, "bd"), ("wag", 150, 333, 1, "bd"), ("wag", 100, 333, 1, "bd"), ("mg94", 32, 100000, 1, "bd")]:
    case = cases.make_case("fd" + model, model, ntaxa, nsites, ncat=ncat, seed=7, tree_kind=kind)
    P = case.host_P()
    a = run(case, P, {"fused_dmma_version": 1})
    b = run(case, P, {"fused_dmma_version": 2})
    b2 = run(case, P, {"fused_dmma_version": 2})
    bad = np.nonzero(np.abs(a - b) > 1e-9 * np.abs(a))[0]
    bad2 = np.nonzero(np.abs(a - b2) > 1e-9 * np.abs(a))[0]
    print(model, ntaxa, nsites, ncat, "nbad", len(bad), len(bad2), "first", bad[:40].tolist(), "units", sorted(set((bad // 8).tolist()))[:40])
    if len(bad): print("  values", a[bad[:5]], b[bad[:5]])
