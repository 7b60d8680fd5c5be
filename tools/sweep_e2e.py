"""End-to-end (pinned host tips -> total on the host) time of pb_loglik_host vs the number of copy chunks."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cases
from phylogenetics_b200.engine import LikelihoodContext
nsites = 1_000_000
case = cases.make_case("c2", "gtr", 128, nsites, ncat=4, seed=100)
f = case.flat
tips = torch.from_numpy(case.tips).pin_memory()
with LikelihoodContext(4, nsites, f.nleaves, ncat=4) as ctx:
    ctx.set_tree(f, case.bl); ctx.set_categories(case.rates, case.weights); ctx.set_rate_matrix(case.Q, case.pi)
    ctx.set_root_frequencies(case.pi)
    for chunks in (1, 2, 3, 4, 6, 8, 12, 16):
        ctx.set_option("host_chunks", chunks)
        for _ in range(2):
            ctx.loglik_host_ptr(tips.data_ptr(), nsites)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            ctx.set_branch_lengths(case.bl)
            tot = ctx.loglik_host_ptr(tips.data_ptr(), nsites)
        dt = (time.perf_counter() - t0) / 10
        print(json.dumps({"host_chunks": chunks, "ms": dt * 1e3, "site_node_per_s": nsites * f.ninternal / dt, "total": tot}), flush=True)
