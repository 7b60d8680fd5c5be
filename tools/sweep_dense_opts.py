"""c3 / c4 pruning time of the default dense path under option sets (run on the GPU box).
usage: sweep_dense_opts.py c3|c4 ["opt=v,opt=v" ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from phylogenetics_b200.engine import LikelihoodContext
name = sys.argv[1]
model, ntaxa, nsites = {"c3": ("wag", 64, 200_000), "c4": ("mg94", 32, 100_000)}[name]
case = cases.make_case(name, model, ntaxa, nsites, seed=100)
f = case.flat
for opts in (sys.argv[2:] or [""]):
    with LikelihoodContext(case.n, nsites, f.nleaves) as ctx:
        ctx.set_tree(f, case.bl); ctx.set_rate_matrix(case.Q, case.pi); ctx.set_root_frequencies(case.pi)
        ctx.set_tips(case.tips); ctx.set_timing(True)
        for kv in filter(None, opts.split(",")):
            k, v = kv.split("=")
            ctx.set_option(k, float(v))
        ms = []
        for _ in range(10):
            ctx.set_branch_lengths(case.bl)
            tot = ctx.loglik()
            ms.append(ctx.get_timing()[1])
        print(json.dumps({"config": name, "opts": opts, "pruning_ms_min": min(ms[2:]), "pruning_ms_median": float(np.median(ms[2:])), "total": tot}), flush=True)
