#!/usr/bin/env python
"""Headline benchmark: pruning site.node updates / second (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--sites S] [--config c2|c3|c4] [--path auto|level] [--no-extras]

A "step" is one full log-likelihood evaluation of one alignment (shard): P(t) for every
branch x category, Felsenstein pruning over the whole tree, root reduction (and, for N > 1,
the all-reduce of the scalar).  Workload at N = 1: BASELINE configs[1], GTR+Gamma4, 128 taxa,
1M sites; for N > 1 every rank holds its own 1M-site shard of ONE problem (same tree and model,
own columns: weak scaling, sites are independent).  One process per GPU under torchrun.

What the one JSON line carries (N = 1): the device-timed value; `e2e` through the C-ABI call that
takes HOST buffers; `roofline` of the dominant kernel against its true bound, with the
level-scheduled HBM model beside it; `parity_gate`: the whole benchmarked alignment against the CPU
oracle (per site <= 1e-9, total <= 1e-10 relative, asserted before anything is timed), the same
oracle pass being the `cpu_baseline` timing; and the blocks `c3`, `c4` (BASELINE configs[2], [3] at
full size, each with its own parity gate, roofline, e2e and cpu_baseline), `level_scheduled`,
`reference_scale_factor`, `tip_masks`, `taxa_sweep`.

The reference arm (--impl reference) times the CPU restatement of the reference's pruning
(oracle/pruning_oracle.c, OpenMP over sites on all host cores) on the same workload: the OCaml
reference itself cannot be built in this image.  That arm never imports the product
library.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line, on the process's real stdout (libraries such as NCCL print banners on fd 1:
    main() points fd 1 at stderr for the duration of the run)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


METRIC = "pruning site-node updates/sec"
UNIT = "site-node updates/s"
PER_SITE_RTOL, TOTAL_RTOL = 1e-9, 1e-10          # BASELINE.json north_star tolerances

CONFIGS = {
    # name: (model, ntaxa, nsites, ncat, description)
    "c2": ("gtr", 128, 1_000_000, 4, "GTR+Gamma4 DNA, 128-taxon birth-death tree, 1M sites"),
    "c3": ("wag", 64, 200_000, 1, "WAG amino-acid 20-state, 64-taxon tree, 200k sites"),
    "c4": ("mg94", 32, 100_000, 1, "MG94 codon 61-state, 32-taxon tree, 100k codons"),
    "c1": ("jc69", 16, 1_000, 1, "JC69 DNA, 16-taxon tree, 1k sites"),
}
CASE_SEED = 100                                   # tree + model of every arm; site_seed = rank draws the columns


def algorithmic_bytes_per_site(n, ncat, ntaxa):
    """B_site = 8*C*(n+1)*(2T-4) + T + 8 (level-scheduled model)."""
    return 8 * ncat * (n + 1) * (2 * ntaxa - 4) + ntaxa + 8


def algorithmic_flops_per_site(n, ncat, ntaxa):
    """F_site = C*(2n^2*(T-2) + n*(T-1)) (tip children are look-ups)."""
    return ncat * (2 * n * n * (ntaxa - 2) + n * (ntaxa - 1))


FP64_DMMA_PEAK_TFLOPS = 36.8      # profiles/r01_fp64_peak_dmma_shapes.jsonl (tools/fp64_peak2.cu)
FP64_DFMA_PEAK_TFLOPS = 34.8      # profiles/r01_fp64_peak.json (tools/fp64_peak.cu)
FP64_PEAK_KIND = ("builder-measured on this pool's B200 (tools/fp64_peak*.cu, profiles/r01_fp64_peak*.json*); "
                  "MEASURED_PEAKS.json carries no FP64 figure; 148 SMs x 64 DFMA lanes x 2 x 1.965 GHz = 37.2")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def ncu_record(kernel):
    """Per-launch dram bytes and pipe activity of `kernel` from the committed ncu capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(kernel) or {}
    except (OSError, ValueError):
        return {}


def build_case(cfg, nsites, seed=CASE_SEED, site_seed=None, ntaxa=None):
    import cases
    model, T, _, ncat, _ = CONFIGS[cfg]
    return cases.make_case(cfg, model, ntaxa or T, nsites, ncat=ncat, seed=seed, site_seed=site_seed)


def oracle_pass(case, P, tips=None, masks=None, nthreads=None):
    """One pass of the CPU oracle over the case (all sites unless `tips` is a slice): (per-site values, seconds)."""
    from oracle import c_oracle as C
    f = case.flat
    t0 = time.perf_counter()
    if masks is not None:
        ps = C.pruning(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi, masks=masks,
                       weights=case.weights, nthreads=nthreads or (os.cpu_count() or 1))
    else:
        ps = C.pruning(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi,
                       tips=case.tips if tips is None else tips, weights=case.weights,
                       nthreads=nthreads or (os.cpu_count() or 1))
    return ps, time.perf_counter() - t0


def parity_gate(per_site, total, expect, what):
    """North-star tolerances on the benchmarked data, asserted: per site and on the total (against the exactly
    rounded sum of the oracle's per-site values)."""
    rel = float(np.max(np.abs(per_site - expect) / np.abs(expect)))
    truth = math.fsum(expect)
    tot_rel = abs(total - truth) / abs(truth)
    if not (rel <= PER_SITE_RTOL and tot_rel <= TOTAL_RTOL):
        raise SystemExit("parity gate failed (%s): per-site rel err %.3e, total rel err %.3e" % (what, rel, tot_rel))
    return {"per_site_max_rel_err": rel, "total_rel_err_vs_oracle_fsum": tot_rel, "sites_checked": int(len(expect)),
            "oracle_total": truth, "total": total, "tolerances": {"per_site": PER_SITE_RTOL, "total": TOTAL_RTOL}}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        inside = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm = sorted(float(r[1]) for r in inside if r[1].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            for nm, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(inside[0][2]) if inside else None,
                "samples": len(inside), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference's pruning, all host cores.  Never imports the product
# library (inputs come from the pure-host builders in tests/cases.py).
# ---------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg = args.config
    model, ntaxa, full_sites, ncat, desc = CONFIGS[cfg]
    cores = os.cpu_count() or 1
    probe = build_case(cfg, 4096)
    P = probe.host_P()
    oracle_pass(probe, P)
    per_site = oracle_pass(probe, P)[1] / 4096
    # the whole configuration when K + W steps of it fit ~3 minutes, else as many sites as do
    budget = 170.0 / max(1, args.steps + args.warmup)
    sample = full_sites if per_site * full_sites * 1.25 <= budget else int(max(4096, budget / per_site / 1.25))
    sample = args.sites or sample
    case = build_case(cfg, sample, site_seed=0)           # the very alignment rank 0 of our arm evaluates
    P = case.host_P()
    f = case.flat
    from oracle import c_oracle as C

    def step():
        ps, _ = oracle_pass(case, P, nthreads=cores)
        return C.sum_sequential(ps)                        # the reference's own left-to-right sum (lib/felsenstein.ml:75)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        total = step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sample * f.ninternal / dt
    sample_desc = ("all %d sites per step" % full_sites) if sample == full_sites else \
        "%d of %d sites per step, same tree and model" % (sample, full_sites)
    assert not any("libphylo_b200" in l for l in open("/proc/self/maps")), "the reference arm loaded the product library"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "sites_per_gpu": sample, "sites_per_step": sample, "ntaxa": ntaxa, "nstates": case.n, "ncat": ncat,
                   "internal_nodes": f.ninternal, "sample": sample_desc, "total_loglik": total,
                   "note": "CPU restatement of the reference algorithm (oracle/pruning_oracle.c, OpenMP over "
                           "sites, P(t) precomputed per branch - the CPU-favourable choice), not the OCaml/lacaml "
                           "binary, which cannot be built in this image"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def pinned_near_device(array, device):
    """Page-locked host copy of `array` on the GPU's NUMA node (product-arm only: imports the package)."""
    from phylogenetics_b200.engine import pinned_near_device as impl
    return impl(array, device)


def kernel_only(ctx, case, flush, flops, peak, reps=5):
    """The dominant kernel ALONE (pb_set_timing(ctx, 2): an event pair around each of its launches), outside the timed
    loop: the pruning phase of phase_ms also holds the per-evaluation table builders and matrix staging."""
    import torch
    ctx.set_timing(2)
    ms, n = [], 0
    for _ in range(reps + 1):
        flush.zero_()
        ctx.set_branch_lengths(case.bl)
        ctx.loglik()
        t, n = ctx.get_kernel_timing()
        ms.append(t)
    ctx.set_timing(True)
    if n < 1:
        return None
    k = float(np.mean(ms[1:]))
    tf = flops / (k * 1e-3) / 1e12
    return {"ms_per_step": k, "launches_per_step": n, "achieved": tf, "frac": tf / peak, "unit": "TFLOP/s",
            "note": "CUDA events around the kernel's own launches (mean of %d evaluations, L2 flushed between them)" % reps}


def new_context(case, nsites, device, stream, level=False):
    from phylogenetics_b200.engine import LikelihoodContext
    f = case.flat
    ctx = LikelihoodContext(case.n, nsites, f.nleaves, ncat=case.ncat, device=device, level_kernels=level)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_tree(f, case.bl)
    ctx.set_categories(case.rates, case.weights)
    if case.reversible:
        ctx.set_rate_matrix(case.Q, case.pi)          # eigensystem on the host once, P(t) on the device every step (K1)
    else:
        ctx.set_rate_matrix_general(case.Q)           # batched Pade kernel
    ctx.set_root_frequencies(case.pi)
    ctx.set_timing(True)
    return ctx


def timed_resident_steps(ctx, case, steps, warmup, stream, flush, after_enqueue=None, barrier=None):
    """K evaluations queued without host synchronisation, one CUDA-event pair each on the launching stream, L2
    flushed between the pairs; returns (per-step ms, per-step [pmat, pruning, root] ms, mid events' split)."""
    import torch
    barrier = barrier or torch.cuda.synchronize

    def step(mid=None):
        ctx.set_branch_lengths(case.bl)      # new branch lengths every evaluation: P(t) is rebuilt (kernel K1)
        ctx.enqueue()
        if mid is not None:
            mid.record(stream)
        if after_enqueue:
            after_enqueue()
    for _ in range(warmup):
        flush.zero_()
        step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    launches0 = ctx.kernel_launches
    t0 = time.time()
    for (a, b), m in zip(ev, mids):
        flush.zero_()                        # L2 flush, outside the event pair
        a.record(stream)
        step(m)
        b.record(stream)
    barrier()
    t1 = time.time()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    local_ms = [a.elapsed_time(m) for (a, b), m in zip(ev, mids)]
    coll_ms = [m.elapsed_time(b) for (a, b), m in zip(ev, mids)]
    phases = np.array([ctx.get_timing(back) for back in range(min(steps, 64) - 1, -1, -1)])
    return step_ms, phases, local_ms, coll_ms, (t0, t1), ctx.kernel_launches - launches0


def dense_block(cfg, device, stream, flush, steps):
    """BASELINE configs[2] / [3] at full size on the FP64 tensor-core path: parity gate over every site, device-timed
    value, roofline against the measured DMMA peak, e2e from pinned host memory, CPU baseline."""
    import torch
    model, ntaxa, nsites, ncat, desc = CONFIGS[cfg]
    case = build_case(cfg, nsites, site_seed=0)
    f, n = case.flat, case.n
    ctx = new_context(case, nsites, device, stream)
    tips_pinned = pinned_near_device(case.tips, device)
    ctx.set_tips_ptr(tips_pinned.data_ptr(), nsites)
    total0, per_site0 = ctx.loglik(per_site=True)
    P_host = case.host_P()
    oracle_pass(case, P_host, tips=np.ascontiguousarray(case.tips[:, :2048]))        # spin the OpenMP team up
    expect, cpu_s = oracle_pass(case, P_host)
    gate = parity_gate(per_site0, total0, expect, cfg)
    step_ms, phases, _, _, _, launches = timed_resident_steps(ctx, case, steps, 3, stream, flush)
    ms = float(np.mean(step_ms))
    prune_ms = float(phases[:, 1].mean())
    flops = algorithmic_flops_per_site(n, ncat, ntaxa) * nsites
    tf = flops / (prune_ms * 1e-3) / 1e12
    rec = ncu_record("fused_dmma3_%s" % cfg)
    konly = kernel_only(ctx, case, flush, flops, FP64_DMMA_PEAK_TFLOPS)
    # end to end: pb_set_branch_lengths + pb_loglik_host (copy from pinned host memory, evaluation, total read back)
    for _ in range(2):
        ctx.loglik_host_ptr(tips_pinned.data_ptr(), nsites)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        ctx.set_branch_lengths(case.bl)
        tot = ctx.loglik_host_ptr(tips_pinned.data_ptr(), nsites)
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / steps
    if tot != total0:
        raise SystemExit("%s: host-buffer evaluation disagrees with the resident one" % cfg)
    block = {
        "config": {"workload": desc, "sites": nsites, "ntaxa": ntaxa, "nstates": n, "ncat": ncat,
                   "internal_nodes": f.ninternal, "path": ctx.path_name,
                   "p_of_t": "eigensystem kernel" if case.reversible else "batched Pade kernel"},
        "value": nsites * f.ninternal / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
        "phase_ms": {"pmat": float(phases[:, 0].mean()), "pruning": prune_ms, "root": float(phases[:, 2].mean())},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": tf, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                     "frac": tf / FP64_DMMA_PEAK_TFLOPS, "traffic": rec.get("dram_bytes_per_launch"),
                     "kernel": "fused_dmma3", "launches_per_step": 1, "kernel_ms_per_step": prune_ms, "kernel_only": konly,
                     "algorithmic_flops_per_launch": flops, "peak_kind": FP64_PEAK_KIND,
                     "pipe_active": rec.get("pipe_active"), "ncu_source": rec.get("source"),
                     "hbm_model": {"algorithmic_bytes_per_launch": algorithmic_bytes_per_site(n, ncat, ntaxa) * nsites,
                                   "achieved_gbs": algorithmic_bytes_per_site(n, ncat, ntaxa) * nsites / (prune_ms * 1e-3) / 1e9},
                     "note": "algorithmic flops (SURVEY 8d: tip children are look-ups, padding not counted) / pruning-phase time"},
        "e2e": {"value": nsites * f.ninternal / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(case.tips.nbytes), "d2h_bytes_per_step": 8, "call": "pb_loglik_host"},
        "parity_gate": gate,
        "cpu_baseline": {"value": nsites * f.ninternal / cpu_s, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                         "sample": "all %d sites, one pass (%.2f s) - the parity gate's oracle pass" % (nsites, cpu_s)},
    }
    ctx.close()
    return block


def masked_block(case, P_host, device, stream, flush, base_prune_ms, gap_frac=0.10):
    """the c2 alignment with 10 % of its cells missing (Missing_values.pruning), through
    pb_set_tip_masks: which kernel runs, its pruning time against the unmasked one, parity on a slice."""
    import torch
    f, n, nsites = case.flat, case.n, case.nsites
    rng = np.random.default_rng(7)
    masks = (np.uint64(1) << case.tips.astype(np.uint64))
    masks[rng.random(masks.shape) < gap_frac] = 0
    ctx = new_context(case, nsites, device, stream)
    ctx.set_tip_masks(masks)
    total, per_site = ctx.loglik(per_site=True)
    chk = min(nsites, 100_000)
    expect, _ = oracle_pass(case, P_host, masks=np.ascontiguousarray(masks[:, :chk]))
    rel = float(np.max(np.abs(per_site[:chk] - expect) / np.maximum(np.abs(expect), 1e-300)))
    if not rel <= PER_SITE_RTOL:
        raise SystemExit("parity gate failed (tip masks): per-site rel err %.3e" % rel)
    ms = []
    for it in range(8):
        flush.zero_()
        ctx.set_branch_lengths(case.bl)
        ctx.loglik()
        if it >= 3:
            ms.append(ctx.get_timing()[1])
    prune = float(np.mean(ms))
    block = {"kernel_path": ctx.path_name, "missing_fraction": gap_frac, "kernel_ms_per_step": prune,
             "value": nsites * f.ninternal / (prune * 1e-3), "unit": UNIT,
             "slowdown_vs_unmasked": prune / base_prune_ms,
             "parity_gate": {"per_site_max_rel_err": rel, "sites_checked": chk},
             "note": "pb_set_tip_masks: state sets per (taxon, site), 0 = missing (lib/phylo_ctmc.ml:145-172)"}
    ctx.close()
    return block


def taxa_sweep_block(device, stream, flush, nsites=250_000):
    """Pruning time of the default 4-state path against the number of taxa (GTR+G4): site-node updates/s should not
    fall off a cliff above the benchmarked 128 taxa."""
    out = []
    for ntaxa in (128, 256, 512, 1024):
        case = build_case("c2", nsites, seed=CASE_SEED + ntaxa, ntaxa=ntaxa)
        ctx = new_context(case, nsites, device, stream)
        ctx.set_tips(case.tips)
        ms = []
        for it in range(7):
            flush.zero_()
            ctx.set_branch_lengths(case.bl)
            tot, ps = ctx.loglik(per_site=True)
            if it >= 2:
                ms.append(ctx.get_timing()[1])
        chk = 4096
        expect, _ = oracle_pass(case, case.host_P(), tips=np.ascontiguousarray(case.tips[:, :chk]))
        rel = float(np.max(np.abs(ps[:chk] - expect) / np.abs(expect)))
        if not rel <= PER_SITE_RTOL:
            raise SystemExit("parity gate failed (%d taxa): %.3e" % (ntaxa, rel))
        prune = float(np.mean(ms))
        out.append({"ntaxa": ntaxa, "sites": nsites, "path": ctx.path_name, "pruning_ms": prune,
                    "site_node_updates_per_s": nsites * case.flat.ninternal / (prune * 1e-3),
                    "per_site_max_rel_err_first_4096": rel})
        ctx.close()
    base = out[0]["site_node_updates_per_s"]
    for r in out:
        r["rate_vs_128_taxa"] = r["site_node_updates_per_s"] / base
    return out


def single_process_block(case, world, tips_pinned, total_local, steps):
    """the same N-GPU evaluation driven by ONE host thread through pb_create_sharded — one context
    per device, ncclAllReduce of the scalar inside the library.  Run by rank 0 while the other ranks wait; every shard
    holds rank 0's columns (the timing does not depend on which columns a device holds), so the total must be
    N x rank 0's local total."""
    from phylogenetics_b200.engine import ShardedContext
    f, nsites = case.flat, case.nsites
    s = ShardedContext(list(range(world)), case.n, nsites * world, f.nleaves, ncat=case.ncat)
    s.set_tree(f, case.bl)
    s.set_categories(case.rates, case.weights)
    s.set_rate_matrix(case.Q, case.pi)
    s.set_root_frequencies(case.pi)
    for d in range(world):
        assert s.shard_bounds(d) == (d * nsites, (d + 1) * nsites)
        s.shard_set_tips_ptr(d, tips_pinned.data_ptr(), nsites)
    for _ in range(3):
        s.set_branch_lengths(case.bl)
        total = s.loglik()
    if not abs(total - world * total_local) <= 1e-12 * abs(total):
        raise SystemExit("single-process sharded total %r != %d x %r" % (total, world, total_local))
    t0 = time.perf_counter()
    for _ in range(steps):
        s.set_branch_lengths(case.bl)
        total = s.loglik()                 # blocking: P(t) + pruning + reduction on every device, all-reduce, read-back
    dt = (time.perf_counter() - t0) / steps
    s.set_reduce(1)
    total_ordered = s.loglik()
    s.close()
    return {"call": "pb_create_sharded + pb_sharded_set_branch_lengths + pb_sharded_loglik (one host thread, %d devices, "
                    "ncclAllReduce(ncclDouble, 1) inside the library)" % world,
            "ms_per_step_wall": dt * 1e3, "value": world * nsites * f.ninternal / dt, "unit": UNIT, "steps": steps,
            "total": total, "total_ordered_reduce": total_ordered,
            "note": "wall clock around blocking calls (host launch latency of N devices from one thread and the final "
                    "synchronisation included), L2 not flushed; every shard holds rank 0's columns"}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = args.config
    model, ntaxa, full_sites, ncat, desc = CONFIGS[cfg]
    nsites = args.sites or full_sites
    # one problem, sharded by site columns: every rank has the SAME tree and model and simulates its own columns
    case = build_case(cfg, nsites, site_seed=rank)
    f = case.flat
    n = case.n
    ninternal = f.ninternal
    cores = os.cpu_count() or 1

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = new_context(case, nsites, local_rank, stream, level=(args.path == "level"))
    tips_pinned = pinned_near_device(case.tips, local_rank)
    ctx.set_tips_ptr(tips_pinned.data_ptr(), nsites)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gate, before anything is timed.  N = 1: the WHOLE benchmarked alignment against the oracle with
    # host-built matrices (scipy expm: independent of the device's P(t) kernel); the oracle pass is also the CPU
    # baseline's timing.  N > 1: a slice per rank (all ranks share the host cores).
    ctx.enqueue()
    total0, per_site0 = ctx.fetch(per_site=True)
    P_host = case.host_P()
    chk = nsites if world == 1 else min(nsites, 20000)
    oracle_pass(case, P_host, tips=np.ascontiguousarray(case.tips[:, :2048]))          # spin the OpenMP team up
    expect, cpu_s = oracle_pass(case, P_host, tips=None if chk == nsites else np.ascontiguousarray(case.tips[:, :chk]),
                                nthreads=max(1, cores // world))
    if chk == nsites:
        gate = parity_gate(per_site0, total0, expect, cfg)
    else:
        rel = float(np.max(np.abs(per_site0[:chk] - expect) / np.abs(expect)))
        if not rel <= PER_SITE_RTOL:
            raise SystemExit("parity gate failed: per-site rel err %.3e" % rel)
        gate = {"per_site_max_rel_err": rel, "sites_checked": chk, "note": "per-rank slice; the full-size gate runs at N = 1"}
    log("parity gate ok", gate)

    total_t = None

    def collective():
        nonlocal total_t
        if world > 1:
            if total_t is None:
                total_t = torch.as_tensor(ctx.total_as_cuda_array(), device="cuda")
            dist.all_reduce(total_t)            # in place on the device-resident total

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    step_ms, phases, local_ms, collective_ms, (t_wall0, t_wall1), launches = timed_resident_steps(
        ctx, case, args.steps, max(args.warmup, 3), stream, flush, after_enqueue=collective, barrier=barrier)
    per_rank = None
    if world > 1:                          # where the step time goes on every rank: local work vs waiting in the collective
        mine = torch.tensor([sum(local_ms) / len(local_ms), sum(collective_ms) / len(collective_ms)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"local_ms": [float(x[0]) for x in allr], "collective_ms": [float(x[1]) for x in allr]}
    t = torch.tensor([sum(step_ms) / 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_max = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- end to end through the C-ABI call that takes HOST buffers: every step copies the step's tips from pinned host
    # memory, evaluates, and reads the (all-reduced) total back.  Nucleotide alignments cross PCIe 2-bit packed
    # (pb_loglik_host_packed2, the layout the binding fills from `leaf_state` for 4-state alphabets,
    # INTEGRATION.md; the packing itself is host work outside the loop, like the CPU arm's ready tips); the
    # byte-per-site call is timed beside it.  N > 1: the non-blocking form of the same call
    # (pb_loglik_host_enqueue), the NCCL all-reduce of the device-resident total queued behind it on the same
    # stream, one read-back of the reduced scalar: one host synchronisation per step.
    e2e_steps = max(2, min(args.steps, 10))
    e2e_reps = []

    def e2e_loop(ptr, ld, packed2):
        def one():
            ctx.set_branch_lengths(case.bl)
            if world == 1:
                return ctx.loglik_host_packed2_ptr(ptr, ld) if packed2 else ctx.loglik_host_ptr(ptr, ld)
            ctx.host_enqueue_ptr(ptr, ld, packed2)
            collective()
            return float(total_t.item())            # D2H of the reduced total: the step's one synchronisation
        for _ in range(2):
            one()
        reps = []                                       # three repetitions of e2e_steps steps; the median one is reported
        for _ in range(3):                              # (host-to-device bandwidth is shared with the box's other GPUs)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(e2e_steps):
                tot = one()
            e1.record(stream)
            barrier()
            tt = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            reps.append(float(tt.item()))
        e2e_reps.append([r / e2e_steps * 1e3 for r in reps])
        t_med = sorted(reps)[1]
        return world * nsites * ninternal * e2e_steps / t_med, tot, t_med / e2e_steps * 1e3

    if world > 1:
        collective()                                   # total_t exists before the e2e loops use it
    e2e_bytes_val, tot_bytes, e2e_bytes_ms = e2e_loop(tips_pinned.data_ptr(), nsites, False)
    e2e_val, e2e_ms, e2e_h2d, e2e_call = e2e_bytes_val, e2e_bytes_ms, int(case.tips.nbytes), "pb_loglik_host (one byte per site)"
    if n == 4:
        from phylogenetics_b200.engine import pack_tips2
        packed_pinned = pinned_near_device(pack_tips2(case.tips), local_rank)
        e2e_val, tot_packed, e2e_ms = e2e_loop(packed_pinned.data_ptr(), packed_pinned.shape[1], True)
        if not abs(tot_packed - tot_bytes) <= 1e-12 * abs(tot_bytes):
            raise SystemExit("packed and byte-per-site evaluations disagree: %r %r" % (tot_packed, tot_bytes))
        e2e_h2d, e2e_call = int(packed_pinned.numel()), "pb_loglik_host_packed2 (2-bit packed nucleotides, 4 sites per byte)"
        if os.environ.get("PB_E2E_TRACE") and world == 1:      # timeline of one more call (pb_debug_host_trace), to stderr
            ctx.set_option("host_trace", 1)
            ctx.loglik_host_packed2_ptr(packed_pinned.data_ptr(), packed_pinned.shape[1])
            print("[bench] e2e timeline", ctx.host_trace(), file=sys.stderr, flush=True)
            ctx.set_option("host_trace", 0)
    if world == 1 and not abs(tot_bytes - total0) <= 1e-12 * abs(total0):
        raise SystemExit("host-buffer and resident evaluations disagree: %r %r" % (tot_bytes, total0))

    c5 = None
    if world > 1 and not args.no_extras and cfg == "c2" and not args.sites:
        # BASELINE configs[4]: the Nelder-Mead loop on 10M sites over 8 GPUs = 1.25M sites per GPU (the same per-GPU
        # shard at other N), 263 parameters, fixed evaluation budget, one captured evaluation replayed (CUDA graph)
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from optimise_c5 import c5_loop
        c5_sites = 1_250_000
        c5_case = build_case("c2", c5_sites, site_seed=1000 + rank)
        c5_ctx = new_context(c5_case, c5_sites, local_rank, stream)
        c5_ctx.set_timing(False)
        c5_ctx.set_tips(c5_case.tips)
        barrier()
        try:
            c5 = c5_loop(c5_case, c5_ctx, world, dist, 400)
        except Exception as e:
            c5 = {"error": repr(e)}
        c5_ctx.close()
        del c5_case
        barrier()
    single = None
    if world > 1:
        # one host thread driving all N devices through the C ABI (pb_create_sharded), while the other ranks idle.
        # They must idle ON THE HOST: a rank waiting in an NCCL barrier keeps a spinning kernel on its GPU, and rank 0's
        # kernels on that GPU would be time-sliced against it (measured: 5.2 ms per step instead of 0.7).
        barrier()
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                single = single_process_block(case, world, tips_pinned, total0, max(5, min(args.steps, 20)))
            except SystemExit:
                store.set("pb_single_done", "failed")
                raise
            except Exception as e:
                single = {"error": repr(e)}
            store.set("pb_single_done", "1")
        else:
            store.wait(["pb_single_done"])
        barrier()
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    try:        # the FP64 peaks of THIS device, measured now (pb_measure_fp64_peak), beside the constants the fractions use
        from phylogenetics_b200.engine import measure_fp64_peak
        dfma_live, dmma_live = measure_fp64_peak(local_rank)
        fp64_live = {"dfma_tflops": dfma_live, "dmma_tflops": dmma_live,
                     "constants_used": {"dfma_tflops": FP64_DFMA_PEAK_TFLOPS, "dmma_tflops": FP64_DMMA_PEAK_TFLOPS},
                     "note": "register-resident dependent chains on every SM (8 per thread, 4 blocks of 256 threads per SM), best of "
                             "three repetitions; the roofline fractions use the constants (measured the same way in round 1)"}
    except Exception as e:
        fp64_live = {"error": repr(e)}
    units = world * nsites * ninternal * args.steps
    value = units / t_max
    prune_avg_ms = float(phases[:, 1].mean())
    bytes_eval = algorithmic_bytes_per_site(n, ncat, ntaxa) * nsites
    flops_eval = algorithmic_flops_per_site(n, ncat, ntaxa) * nsites
    hbm_ach = bytes_eval / (prune_avg_ms * 1e-3) / 1e9
    tf = flops_eval / (prune_avg_ms * 1e-3) / 1e12
    path = ctx.path_name
    if path == "fused4":
        kernel_name, nlaunch = "fused4t", max(1, launches // args.steps - 5)     # (P(t), compaction, tables, K3 x 2 are the other five)
        rec = ncu_record(kernel_name)
        konly = kernel_only(ctx, case, flush, flops_eval, FP64_DFMA_PEAK_TFLOPS)
        if konly:
            nlaunch = konly["launches_per_step"]
        roof = {"bound": "fp64", "achieved": tf, "peak": FP64_DFMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": tf / FP64_DFMA_PEAK_TFLOPS, "traffic": rec.get("dram_bytes_per_launch"),
                "kernel": kernel_name, "launches_per_step": nlaunch, "kernel_ms_per_step": prune_avg_ms, "kernel_only": konly,
                "algorithmic_flops_per_launch": flops_eval / nlaunch, "peak_kind": FP64_PEAK_KIND,
                "pipe_active": rec.get("pipe_active"), "ncu_source": rec.get("source"),
                "hbm_model": {"bound": "hbm", "achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                              "frac": hbm_ach / peaks["hbm_gbs"], "peak_kind": "of " + peak_kind,
                              "algorithmic_bytes_per_launch": bytes_eval / nlaunch,
                              "traffic": rec.get("dram_bytes_per_launch"),
                              "note": "SURVEY 8(d) bytes of the LEVEL-SCHEDULED model / pruning time: above 1 because the "
                                      "site-resident kernel never writes a conditional-likelihood array to HBM (it moves "
                                      "ntaxa/4 + 8C bytes per site); not a roofline - see level_scheduled for the kernel "
                                      "that the HBM roofline bounds"},
                "note": "site-resident kernel: conditional likelihoods stay on chip, so its bound is the FP64 pipe (and "
                        "the L1 gather path of its clade tables), not HBM.  achieved = SURVEY 8(d) algorithmic flops / "
                        "pruning-phase time (P(t) compaction, table build and the C category launches included); the "
                        "tables remove part of those flops, so pipe_active (ncu) is below frac"}
    elif path == "fused_dmma":
        rec = ncu_record("fused_dmma3_%s" % cfg)
        roof = {"bound": "tensor", "achieved": tf, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": tf / FP64_DMMA_PEAK_TFLOPS, "traffic": rec.get("dram_bytes_per_launch"),
                "peak_kind": FP64_PEAK_KIND, "kernel": "fused_dmma3", "launches_per_step": 1,
                "kernel_ms_per_step": prune_avg_ms, "kernel_only": kernel_only(ctx, case, flush, flops_eval, FP64_DMMA_PEAK_TFLOPS),
                "algorithmic_flops_per_launch": flops_eval,
                "pipe_active": rec.get("pipe_active"), "ncu_source": rec.get("source"),
                "hbm_model": {"algorithmic_bytes_per_launch": bytes_eval, "achieved_gbs": hbm_ach}}
    else:
        nlaunch = max(1, launches // args.steps - 3)
        rec = ncu_record("level4_async")
        roof = {"bound": "hbm", "achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": hbm_ach / peaks["hbm_gbs"], "traffic": rec.get("dram_bytes_per_step"), "peak_kind": "of " + peak_kind,
                "kernel": path, "launches_per_step": nlaunch, "kernel_ms_per_step": prune_avg_ms,
                "algorithmic_bytes_per_step": bytes_eval,
                "note": "level-scheduled: every inner CLV written once and read once"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc + (" per GPU" if world > 1 else ""), "sites_per_gpu": nsites,
                   "ntaxa": ntaxa, "nstates": n, "ncat": ncat, "internal_nodes": ninternal, "path": path,
                   "rescaling": "reference trigger (min v <= threshold), exact power-of-two factor; see reference_scale_factor",
                   "l2": "flushed between steps (256 MiB memset outside the event pairs)",
                   "timing": "sum of per-step CUDA-event pairs on the launching stream, max over ranks",
                   "parity_gate": gate},
        "roofline": roof,
        "fp64_peaks_live": fp64_live,
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms, "reps_ms_per_step": e2e_reps[-1], "h2d_bytes_per_step": e2e_h2d,
                "d2h_bytes_per_step": 8, "steps": e2e_steps, "call": e2e_call,
                "byte_per_site": {"value": e2e_bytes_val, "ms_per_step": e2e_bytes_ms,
                                  "h2d_bytes_per_step": int(case.tips.nbytes), "call": "pb_loglik_host"},
                "note": "pb_set_branch_lengths + one host-alignment evaluation call per step: tips copied from pinned host "
                        "memory in column chunks overlapped with the pruning, total read back to the host (N > 1: "
                        "pb_loglik_host_enqueue + NCCL all-reduce of the device-resident total + one read-back); the "
                        "2-bit packing of the host alignment is done once outside the loop; median of three repetitions of "
                        "`steps` steps (reps_ms_per_step), host buffers page-locked on the GPU's NUMA node"},
        "collective_ms_per_step": (sum(collective_ms) / len(collective_ms)) if world > 1 else 0.0,
        "per_rank_ms": per_rank,
        "single_process_sharded": single,
        "c5_nelder_mead": c5,
        "gpu_launches": int(launches),
        "phase_ms": {"pmat": float(phases[:, 0].mean()), "pruning": prune_avg_ms, "root": float(phases[:, 2].mean())},
        "clocks": clocks,
    }
    line["cpu_baseline"] = None
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = {
            "value": chk * ninternal / cpu_s, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "all %d sites, one pass (%.2f s) - the parity gate's oracle pass; CPU restatement of the reference "
                      "algorithm (oracle/pruning_oracle.c, OpenMP over sites), not the OCaml binary" % (chk, cpu_s)}
    extras = world == 1 and not args.no_extras and not args.sites and cfg == "c2"
    if world == 1 and path in ("fused4", "fused_dmma"):
        # the same site-resident kernel with the reference's own scale factor (v := v / max v) instead of
        # the default exact power of two ("Rescaling"): time and agreement
        ctx.set_option("fused_pow2", 0)
        if n == 4:
            ctx.set_option("fused_spt", 2)          # that mode's fastest shape (profiles/r01f_sweep_c2_fused4c.jsonl)
        ms = []
        for it in range(8):
            flush.zero_()
            ctx.set_branch_lengths(case.bl)
            tot_ref, ps_ref = ctx.loglik(per_site=True)
            if it >= 3:
                ms.append(sum(ctx.get_timing()))
        ref_ms = float(np.mean(ms))
        line["reference_scale_factor"] = {
            "ms_per_step_device_phases": ref_ms, "value": nsites * ninternal / (ref_ms * 1e-3), "unit": UNIT,
            "per_site_max_rel_diff_vs_default": float(np.max(np.abs(ps_ref - per_site0) / np.abs(per_site0))),
            "per_site_max_rel_err_vs_oracle": float(np.max(np.abs(ps_ref[:chk] - expect) / np.abs(expect))),
            "note": "option fused_pow2=0: rescaling multiplies by 1/max v exactly as SV.shift does; the default "
                    "multiplies by 2^-exponent(max v), which is exact in floating point"}
        ctx.set_option("fused_pow2", 1)
        if n == 4:
            ctx.set_option("fused_spt", 4)
    ctx.close()
    if world == 1 and n == 4 and path == "fused4":
        # the level-scheduled kernel the north star describes (CLVs materialised in HBM), same inputs
        lv = new_context(case, nsites, local_rank, stream, level=True)
        lv.set_tips_ptr(tips_pinned.data_ptr(), nsites)
        ms = []
        for it in range(8):
            flush.zero_()
            lv.set_branch_lengths(case.bl)
            tot_lv = lv.loglik()
            if it >= 3:
                ms.append(lv.get_timing()[1])
        lv_ms = float(np.mean(ms))
        lv_ach = bytes_eval / (lv_ms * 1e-3) / 1e9
        rec = ncu_record("level4_async")
        line["level_scheduled"] = {
            "kernel": "level4_async", "path": lv.path_name, "kernel_ms_per_step": lv_ms,
            "value": nsites * ninternal / (lv_ms * 1e-3), "unit": UNIT,
            "roofline": {"bound": "hbm", "achieved": lv_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": lv_ach / peaks["hbm_gbs"], "peak_kind": "of " + peak_kind,
                         "frac_of_8TBs_datasheet": lv_ach / 8000.0,
                         "traffic": rec.get("dram_bytes_per_step"), "ncu_source": rec.get("source"),
                         "algorithmic_bytes_per_step": bytes_eval},
            "total_rel_diff_vs_fused": abs(tot_lv - total0) / abs(total0),
            "note": "PB_FLAG_LEVEL_KERNELS: one launch per tree level, every inner CLV written to and read "
                    "from HBM once (the algorithmic-bytes model of SURVEY.md 8d)"}
        lv.close()
    if extras:
        t0 = time.time()
        try:
            line["tip_masks"] = masked_block(case, P_host, local_rank, stream, flush, prune_avg_ms)
        except Exception as e:                      # an extra block must not take the headline down
            line["tip_masks"] = {"error": repr(e)}
        log("tip_masks block %.1f s" % (time.time() - t0))
        del tips_pinned
        for extra in ("c3", "c4"):
            t0 = time.time()
            try:
                line[extra] = dense_block(extra, local_rank, stream, flush, steps=10)
            except SystemExit:
                raise
            except Exception as e:
                line[extra] = {"error": repr(e)}
            log("%s block %.1f s" % (extra, time.time() - t0))
        t0 = time.time()
        try:
            line["taxa_sweep"] = taxa_sweep_block(local_rank, stream, flush)
        except SystemExit:
            raise
        except Exception as e:
            line["taxa_sweep"] = {"error": repr(e)}
        log("taxa_sweep block %.1f s" % (time.time() - t0))
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--sites", type=int, default=0, help="override sites per GPU (testing)")
    ap.add_argument("--path", default="auto", choices=["auto", "level"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the c3 / c4 / tip-mask / taxa-sweep blocks")
    args = ap.parse_args()
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                       # anything else written to fd 1 (NCCL banner, ...) goes to stderr
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
