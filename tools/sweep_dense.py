"""Times the dense-alphabet level kernels (DMMA vs scalar DFMA) on the c3 / c4 workloads."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from phylogenetics_b200.engine import LikelihoodContext

for name, model, ntaxa, nsites in [("c3", "wag", 64, 200_000), ("c4", "mg94", 32, 100_000)]:
    case = cases.make_case(name, model, ntaxa, nsites, seed=100)
    f = case.flat
    ref = None
    variants = [{"fused_dmma_version": 3}, {"fused_dmma_version": 2}, {"fused_dmma_version": 1}, {"fused_dmma": 0}, {"fused_dmma": 0, "dmma_version": 1}, {"fused_dmma": 0, "dense_scalar": 1}]
    if len(sys.argv) > 1:
        variants = [json.loads(a) for a in sys.argv[1:]]
    for opts in variants:
        with LikelihoodContext(case.n, nsites, f.nleaves) as ctx:
            ctx.set_tree(f, case.bl)
            ctx.set_rate_matrix(case.Q, case.pi)
            ctx.set_root_frequencies(case.pi)
            ctx.set_tips(case.tips)
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.set_timing(True)
            ms = []
            for _ in range(5):
                tot = ctx.loglik()
                ms.append(ctx.get_timing())
            ref = tot if ref is None else ref
            n, T = case.n, ntaxa
            byts = (8 * (n + 1) * (2 * T - 4) + T + 8) * nsites
            flops = (2 * n * n * (T - 2) + n * (T - 1)) * nsites
            best = min(m[1] for m in ms[1:])
            print(json.dumps({"config": name, "path": ctx.path_name, "opts": opts, "pmat_ms": ms[-1][0], "pruning_ms": best,
                              "root_ms": ms[-1][2], "GBps": byts / best / 1e6, "TFLOPs": flops / best / 1e9,
                              "total": tot, "rel_vs_first": abs(tot - ref) / abs(ref),
                              "launches": ctx.kernel_launches}), flush=True)
