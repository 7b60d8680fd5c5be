"""Runs a few fused4 evaluations of the c2 workload with the given variant (for ncu captures).
usage: run_fused_once.py BLOCK SPT [nsites]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from phylogenetics_b200.engine import LikelihoodContext
block, spt = int(sys.argv[1]), int(sys.argv[2])
nsites = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
case = cases.make_case("c2", "gtr", 128, nsites, ncat=4, seed=100)
f = case.flat
with LikelihoodContext(4, nsites, f.nleaves, ncat=4) as ctx:
    ctx.set_tree(f, case.bl); ctx.set_categories(case.rates, case.weights); ctx.set_rate_matrix(case.Q, case.pi)
    ctx.set_root_frequencies(case.pi); ctx.set_tips(case.tips)
    ctx.set_option("fused_block", block); ctx.set_option("fused_spt", spt)
    for kv in filter(None, os.environ.get("OPTS", "").split(",")):       # e.g. OPTS=fused_preapply=1
        name, val = kv.split("=")
        ctx.set_option(name, float(val))
    ctx.set_timing(True)
    for _ in range(4):
        tot = ctx.loglik()
    print(block, spt, ctx.get_timing(), tot)
