"""End-to-end time of pb_loglik_host on the dense path (c3 / c4) against host_chunks (run on the GPU box)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cases
from phylogenetics_b200.engine import LikelihoodContext, pinned_near_device
for name, model, ntaxa, nsites in (("c3", "wag", 64, 200_000), ("c4", "mg94", 32, 100_000)):
    case = cases.make_case(name, model, ntaxa, nsites, seed=100)
    f = case.flat
    tips = pinned_near_device(case.tips, 0)
    for opts in sys.argv[1:] or [""]:
        with LikelihoodContext(case.n, nsites, f.nleaves) as ctx:
            ctx.set_tree(f, case.bl); ctx.set_rate_matrix(case.Q, case.pi); ctx.set_root_frequencies(case.pi)
            for kv in filter(None, opts.split(",")):
                k, v = kv.split("=")
                ctx.set_option(k, float(v))
            ctx.set_option("host_trace", 1)
            call = lambda: ctx.loglik_host_ptr(tips.data_ptr(), nsites)
            call(); call()
            wall = []
            for rep in range(8):
                ctx.set_branch_lengths(case.bl)
                torch.cuda.synchronize(); t0 = time.perf_counter(); tot = call(); wall.append((time.perf_counter() - t0) * 1e3)
            try:
                chunks, end = ctx.host_trace()
            except Exception:
                chunks, end = [], None
            print(json.dumps({"config": name, "opts": opts, "wall_ms_min": min(wall), "wall_ms_median": float(np.median(wall)), "end_ms": end,
                              "chunks": [[int(a), int(b), round(c_, 3), round(d, 3)] for a, b, c_, d in chunks], "total": tot}), flush=True)
