"""A/B of two builds of libphylo_b200.so on the c2 workload (run once per build with PHYLO_B200_LIB set):
pruning-phase time of the default path, end-to-end time from pinned host memory (2-bit packed and byte per
site), and the per-site agreement with the oracle on a slice.  Run on the GPU box."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cases
from oracle import c_oracle as C
from phylogenetics_b200.engine import LikelihoodContext, pack_tips2, pinned_near_device

nsites = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
case = cases.make_case("c2", "gtr", 128, nsites, ncat=4, seed=100)
f = case.flat
tips = pinned_near_device(case.tips, 0)
packed = pinned_near_device(pack_tips2(case.tips), 0)
rec = {"lib": os.environ.get("PHYLO_B200_LIB", "default"), "nsites": nsites}
with LikelihoodContext(4, nsites, f.nleaves, ncat=4) as ctx:
    ctx.set_tree(f, case.bl); ctx.set_categories(case.rates, case.weights)
    ctx.set_rate_matrix(case.Q, case.pi); ctx.set_root_frequencies(case.pi)
    ctx.set_tips(case.tips)
    for kv in filter(None, os.environ.get("OPTS", "").split(",")):       # e.g. OPTS=fused_const_group=0,fused_spt=2
        name, val = kv.split("=")
        ctx.set_option(name, float(val))
    rec["opts"] = os.environ.get("OPTS", "")
    ctx.set_timing(True)
    ms = []
    for _ in range(12):
        tot, ps = ctx.loglik(per_site=True)
        ms.append(ctx.get_timing()[1])
    P = ctx.get_pmatrices()
    chk = 20000
    expect = C.pruning(4, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi,
                       tips=np.ascontiguousarray(case.tips[:, :chk]), weights=case.weights)
    rec.update(path=ctx.path_name, pruning_ms_min=min(ms[2:]), pruning_ms_median=float(np.median(ms[2:])), total=tot,
               max_rel_vs_oracle=float(np.max(np.abs(ps[:chk] - expect) / np.abs(expect))))
    for name, call in (("packed2", lambda: ctx.loglik_host_packed2_ptr(packed.data_ptr(), packed.shape[1])),
                       ("bytes", lambda: ctx.loglik_host_ptr(tips.data_ptr(), nsites))):
        t = call(); call()
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(10):
                ctx.set_branch_lengths(case.bl); t = call()
            torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) / 10)
        rec["e2e_%s_ms" % name] = best * 1e3
        rec["e2e_%s_phases_ms" % name] = ctx.get_timing()
        rec["e2e_%s_rel_vs_resident" % name] = abs(t - tot) / abs(tot)
print(json.dumps(rec), flush=True)
