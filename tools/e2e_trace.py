"""Timeline of the pipelined host evaluation (pb_loglik_host_packed2) at c2 under a list of option sets: per chunk the
time its copy and its compute ended (option host_trace, pb_debug_host_trace), and the wall-clock time per call.
Run on the GPU box.  usage: e2e_trace.py ["opt=v,opt=v" ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cases
from phylogenetics_b200.engine import LikelihoodContext, pack_tips2, pinned_near_device

nsites = int(os.environ.get("NSITES", 1_000_000))
case = cases.make_case("c2", "gtr", 128, nsites, ncat=4, seed=100)
f = case.flat
packed = pinned_near_device(pack_tips2(case.tips), 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
optsets = sys.argv[1:] or [""]
# a plain copy of the same bytes for reference: one flat cudaMemcpyAsync and the library's 2-D form
dst = torch.empty_like(packed, device="cuda")
for name, fn in (("flat", lambda: dst.copy_(packed, non_blocking=True)),):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(6):
        flush.zero_(); torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    print(json.dumps({"copy": name, "bytes": packed.numel(), "ms_min": min(ms[1:]), "ms_median": float(np.median(ms[1:]))}), flush=True)
for opts in optsets:
    with LikelihoodContext(4, nsites, f.nleaves, ncat=4) as ctx:
        ctx.set_tree(f, case.bl); ctx.set_categories(case.rates, case.weights)
        ctx.set_rate_matrix(case.Q, case.pi); ctx.set_root_frequencies(case.pi)
        for kv in filter(None, opts.split(",")):
            name, val = kv.split("=")
            ctx.set_option(name, float(val))
        ctx.set_option("host_trace", 1)
        call = lambda: ctx.loglik_host_packed2_ptr(packed.data_ptr(), packed.shape[1])
        tot = call(); call()
        wall, traces = [], []
        for rep in range(8):
            flush.zero_(); torch.cuda.synchronize()
            ctx.set_branch_lengths(case.bl)
            t0 = time.perf_counter(); tot = call(); wall.append((time.perf_counter() - t0) * 1e3)
            traces.append(ctx.host_trace())
        best = int(np.argmin(wall))
        chunks, end = traces[best]
        print(json.dumps({"opts": opts, "total": tot, "wall_ms_min": min(wall), "wall_ms_median": float(np.median(wall)),
                          "end_ms": end, "launches_per_call": None,
                          "chunks": [{"site0": int(a), "sites": int(b), "copy_done_ms": round(c_, 4), "compute_done_ms": round(d, 4)}
                                     for a, b, c_, d in chunks]}), flush=True)
