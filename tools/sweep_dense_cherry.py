"""Pruning time of the dense-alphabet site-resident kernel (fused_dmma3) with and without tabulated cherries,
BASELINE configs[2] (WAG, 64 taxa, 200k sites) and [3] (MG94, 32 taxa, 100k codons)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from oracle import c_oracle as C
from phylogenetics_b200.engine import LikelihoodContext
for name, model, ntaxa, nsites in (("c3", "wag", 64, 200_000), ("c4", "mg94", 32, 100_000)):
    case = cases.make_case(name, model, ntaxa, nsites, seed=100, site_seed=0)
    f = case.flat
    with LikelihoodContext(case.n, nsites, f.nleaves) as ctx:
        ctx.set_tree(f, case.bl); ctx.set_rate_matrix(case.Q, case.pi); ctx.set_root_frequencies(case.pi)
        ctx.set_tips(case.tips); ctx.set_timing(True)
        P = None
        for cherry, triple in ((1, 1), (1, 0), (0, 0), (1, 1)):
            ctx.set_option("fused_dmma_cherry", cherry)
            ctx.set_option("fused_dmma_triple", triple)
            ms = []
            for _ in range(10):
                ctx.set_branch_lengths(case.bl)
                tot, ps = ctx.loglik(per_site=True)
                ms.append(ctx.get_timing()[1])
            if P is None:
                P = ctx.get_pmatrices()
                chk = 5000
                expect = C.pruning(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi,
                                   tips=np.ascontiguousarray(case.tips[:, :chk]), weights=case.weights)
            rel = float(np.max(np.abs(ps[:chk] - expect) / np.abs(expect)))
            print(json.dumps({"config": name, "fused_dmma_cherry": cherry, "fused_dmma_triple": triple, "pruning_ms_min": min(ms[2:]),
                              "pruning_ms_median": float(np.median(ms[2:])), "total": tot, "max_rel_vs_oracle_5000_sites": rel,
                              "launches": ctx.kernel_launches}), flush=True)
