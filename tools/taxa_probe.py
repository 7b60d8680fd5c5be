"""Pruning time of the default 4-state path at a given number of taxa under option sets (run on the GPU box).
usage: taxa_probe.py NTAXA ["opt=v,opt=v" ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
ntaxa = int(sys.argv[1]); nsites = 250_000
case = bench.build_case("c2", nsites, seed=bench.CASE_SEED + ntaxa, ntaxa=ntaxa)
stream = torch.cuda.Stream()
for opts in (sys.argv[2:] or [""]):
    ctx = bench.new_context(case, nsites, 0, stream)
    ctx.set_tips(case.tips)
    for kv in filter(None, opts.split(",")):
        name, val = kv.split("=")
        ctx.set_option(name, float(val))
    ms = []
    l0 = ctx.kernel_launches
    for it in range(7):
        ctx.set_branch_lengths(case.bl)
        tot = ctx.loglik()
        ms.append(ctx.get_timing()[1])
    print(json.dumps({"ntaxa": ntaxa, "opts": opts, "pruning_ms_min": min(ms[2:]), "pruning_ms_mean": float(np.mean(ms[2:])), "total": tot,
                      "launches_per_eval": (ctx.kernel_launches - l0) / 7}), flush=True)
    ctx.close()
