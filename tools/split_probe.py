"""Fixed cost of a launch of the reduced 4-state program: the resident c2 evaluation cut into N site ranges (option
fused_split), with and without reloading the constant bank per launch.  Run on the GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from phylogenetics_b200.engine import LikelihoodContext
case = cases.make_case("c2", "gtr", 128, 1_000_000, ncat=4, seed=100)
f = case.flat
with LikelihoodContext(4, 1_000_000, f.nleaves, ncat=4) as ctx:
    ctx.set_tree(f, case.bl); ctx.set_categories(case.rates, case.weights)
    ctx.set_rate_matrix(case.Q, case.pi); ctx.set_root_frequencies(case.pi)
    ctx.set_tips(case.tips); ctx.set_timing(True)
    for keep in (1, 0):
        for split in (1, 2, 4, 5, 8, 16, 32):
            ctx.set_option("fused_const_keep", keep); ctx.set_option("fused_split", split)
            ms = []
            for _ in range(8):
                ctx.set_branch_lengths(case.bl)
                tot = ctx.loglik()
                ms.append(ctx.get_timing()[1])
            print(json.dumps({"fused_split": split, "fused_const_keep": keep, "pruning_ms_min": min(ms[2:]), "total": tot}), flush=True)
