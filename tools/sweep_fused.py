"""Times the fused4 kernel variants (threads per block x sites per thread, exact-log on/off)
and the level-scheduled path on the c2 workload.  Run on the GPU box."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from phylogenetics_b200.engine import LikelihoodContext

nsites = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
case = cases.make_case("c2", "gtr", 128, nsites, ncat=4, seed=100)
f = case.flat
out = []
ref = None
VARIANTS = [(256, 1, 0, 1), (256, 2, 0, 1), (512, 1, 0, 1), (512, 2, 0, 1), (128, 2, 0, 1), (128, 4, 0, 1), (256, 4, 0, 1), (256, 1, 0, 0), (256, 2, 0, 0), (512, 1, 0, 0),
            (128, 2, 0, 0), (256, 1, 1, 0), ("level", 4, 0, 0), ("level_async", 24, 0, 0), ("level_async", 32, 0, 0),
            ("level_async", 8, 0, 0), ("level_async", 16, 0, 0), ("level_async", 64, 0, 0)]
for block, spt, exact, const in VARIANTS:
    level = str(block).startswith("level")
    with LikelihoodContext(4, nsites, f.nleaves, ncat=4, level_kernels=level) as ctx:
        ctx.set_tree(f, case.bl)
        ctx.set_categories(case.rates, case.weights)
        ctx.set_rate_matrix(case.Q, case.pi)
        ctx.set_root_frequencies(case.pi)
        ctx.set_tips(case.tips)
        if block == "level_async":
            ctx.set_option("level4_async", 1)
            ctx.set_option("level4_waves", spt)
        elif level:
            ctx.set_option("level4_async", 0)
            ctx.set_option("level4_min_blocks", spt)
        else:
            ctx.set_option("fused_const", const)
            ctx.set_option("fused_block", block)
            ctx.set_option("fused_spt", spt)
            ctx.set_option("fused_exact_log", exact)
        ctx.set_timing(True)
        ms = []
        for _ in range(6):
            tot = ctx.loglik()
            ms.append(ctx.get_timing()[1])
        if ref is None:
            ref = tot
        rec = {"variant": "%sx%s exact=%d const=%d" % (block, spt, exact, const), "path": ctx.path_name,
               "pruning_ms_min": min(ms[1:]), "pruning_ms_first": ms[0], "total": tot,
               "rel_vs_first": abs(tot - ref) / abs(ref)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
