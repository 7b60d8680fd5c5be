"""fused4c with the reference scale factor (1/max) vs power-of-two scale factors, c2 workload; reports the
pruning time and the largest per-site relative difference between the two and against the oracle on a slice."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from oracle import c_oracle as C
from phylogenetics_b200.engine import LikelihoodContext

nsites = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
case = cases.make_case("c2", "gtr", 128, nsites, ncat=4, seed=100)
f = case.flat
res = {}
CATS = int(os.environ.get("CATS", "0"))
CHERRY = int(os.environ.get("CHERRY", "1"))
for block, spt, pow2 in [(256, 2, 0), (256, 2, 1), (256, 1, 1), (128, 2, 1), (256, 4, 1)]:
    with LikelihoodContext(4, nsites, f.nleaves, ncat=4) as ctx:
        ctx.set_tree(f, case.bl)
        ctx.set_categories(case.rates, case.weights)
        ctx.set_rate_matrix(case.Q, case.pi)
        ctx.set_root_frequencies(case.pi)
        ctx.set_tips(case.tips)
        ctx.set_option("fused_block", block)
        ctx.set_option("fused_spt", spt)
        ctx.set_option("fused_pow2", pow2)
        ctx.set_option("fused_cherry", CHERRY)
        if CATS:
            ctx.set_option("fused_const_cats", CATS)
        ctx.set_timing(True)
        ms = []
        for _ in range(6):
            tot, ps = ctx.loglik(per_site=True)
            ms.append(ctx.get_timing()[1])
        P = ctx.get_pmatrices()
    res[(block, spt, pow2)] = ps
    chk = 20000
    expect = C.pruning(4, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi, tips=np.ascontiguousarray(case.tips[:, :chk]),
                       weights=case.weights)
    print(json.dumps({"variant": "%dx%d pow2=%d" % (block, spt, pow2), "pruning_ms_min": min(ms[1:]), "total": tot,
                      "max_rel_vs_reference_factor": float(np.max(np.abs(ps - res[(256, 2, 0)]) / np.abs(ps))),
                      "max_rel_vs_oracle_20000_sites": float(np.max(np.abs(ps[:chk] - expect) / np.abs(expect)))}), flush=True)
