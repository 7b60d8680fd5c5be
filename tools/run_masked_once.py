"""A few evaluations of the c2 workload with 10 % of the cells missing (pb_set_tip_masks), for ncu captures and A/B.
usage: run_masked_once.py [nsites]   (OPTS=opt=v,... sets options)"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from phylogenetics_b200.engine import LikelihoodContext
nsites = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
case = cases.make_case("c2", "gtr", 128, nsites, ncat=4, seed=100)
f = case.flat
rng = np.random.default_rng(7)
masks = np.uint64(1) << case.tips.astype(np.uint64)
masks[rng.random(masks.shape) < 0.1] = 0
with LikelihoodContext(4, nsites, f.nleaves, ncat=4) as ctx:
    ctx.set_tree(f, case.bl); ctx.set_categories(case.rates, case.weights); ctx.set_rate_matrix(case.Q, case.pi)
    ctx.set_root_frequencies(case.pi); ctx.set_tip_masks(masks)
    for kv in filter(None, os.environ.get("OPTS", "").split(",")):
        name, val = kv.split("=")
        ctx.set_option(name, float(val))
    ctx.set_timing(True)
    ms = []
    for _ in range(6):
        ctx.set_branch_lengths(case.bl)
        tot = ctx.loglik()
        ms.append(ctx.get_timing()[1])
    print(json.dumps({"masked": True, "opts": os.environ.get("OPTS", ""), "path": ctx.path_name, "pruning_ms_min": min(ms[2:]), "total": tot}))
