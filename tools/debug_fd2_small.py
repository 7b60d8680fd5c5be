import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from phylogenetics_b200.engine import LikelihoodContext
model, ntaxa, nsites, ncat = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
case = cases.make_case("fd" + model, model, ntaxa, nsites, ncat=ncat, seed=7)
P = case.host_P()
def run(opts, reps=1):
    out = []
    with LikelihoodContext(case.n, case.nsites, case.flat.nleaves, ncat=case.ncat) as ctx:
        ctx.set_tree(case.flat, case.bl); ctx.set_categories(case.rates, case.weights)
        ctx.set_pmatrices(P); ctx.set_tips(case.tips); ctx.set_root_frequencies(case.pi)
        for k, v in opts.items(): ctx.set_option(k, v)
        for _ in range(reps): out.append(ctx.loglik(per_site=True)[1])
    return out
a = run({"fused_dmma_version": 1})[0]
flags = int(sys.argv[6]) if len(sys.argv) > 6 else 0
for b in run({"fused_dmma_version": 2, "fused_dmma_stagger": flags}, reps):
    bad = np.nonzero(np.abs(a - b) > 1e-9 * np.abs(a))[0]
    print("nbad", len(bad), bad[:24].tolist())
