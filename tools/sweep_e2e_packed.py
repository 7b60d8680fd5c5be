"""End-to-end time of one evaluation from a pinned host alignment (c2 shape), byte-per-site vs 2-bit packed,
as a function of the number of column chunks of the copy/compute pipeline."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cases
from phylogenetics_b200.engine import LikelihoodContext, pack_tips2
nsites = 1_000_000
LEADS = [int(x) for x in os.environ.get('LEADS', '1,2,3,4').split(',')]   # host_lead_cats values to sweep
case = cases.make_case("c2", "gtr", 128, nsites, ncat=4, seed=100)
f = case.flat
tips = torch.from_numpy(case.tips).pin_memory()
packed = torch.from_numpy(pack_tips2(case.tips)).pin_memory()
with LikelihoodContext(4, nsites, f.nleaves, ncat=4) as ctx:
    ctx.set_tree(f, case.bl); ctx.set_categories(case.rates, case.weights)
    ctx.set_rate_matrix(case.Q, case.pi); ctx.set_root_frequencies(case.pi)
    ctx.set_timing(True)
    for chunks, lead in [(c, l) for c in (1, 2, 3, 4, 5, 6, 8) for l in LEADS] + [(4, 0)]:
        ctx.set_option("host_chunks", chunks)
        if lead:
            ctx.set_option("host_lead_cats", lead)
        for name, call in (("packed2", lambda: ctx.loglik_host_packed2_ptr(packed.data_ptr(), packed.shape[1])),
                           ("bytes", lambda: ctx.loglik_host_ptr(tips.data_ptr(), nsites))):
            if (name == "bytes") != (lead == 0) or (chunks == 1 and lead > 1):
                continue
            call(); call()
            best = 1e9
            for rep in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                for _ in range(10):
                    ctx.set_branch_lengths(case.bl); call()
                torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) / 10)
            print(json.dumps({"host_chunks": chunks, "host_lead_cats": lead, "input": name, "ms_per_evaluation": best * 1e3,
                              "device_phases_ms_last": ctx.get_timing()}), flush=True)
