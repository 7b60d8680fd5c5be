#!/usr/bin/env python
"""Headline benchmark: pruning site.node updates / second (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--sites S] [--config c2|c3|c4] [--path auto|level]

A "step" is one full log-likelihood evaluation of one alignment (shard): P(t) for every
branch x category, Felsenstein pruning over the whole tree, root reduction (and, for N > 1,
the all-reduce of the scalar).  Workload at N = 1: BASELINE configs[1], GTR+Gamma4, 128 taxa,
1M sites; for N > 1 every rank holds its own 1M-site shard (weak scaling, sites are
independent).  One process per GPU under torchrun.

The reference arm (--impl reference) times the CPU restatement of the reference's pruning
(oracle/pruning_oracle.c, OpenMP over sites on all host cores) on a bounded sample of the
same workload: the OCaml reference itself cannot be built in this image.
"""
import argparse
import json
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


METRIC = "pruning site-node updates/sec"
UNIT = "site-node updates/s"

CONFIGS = {
    # name: (model, ntaxa, nsites, ncat, description)
    "c2": ("gtr", 128, 1_000_000, 4, "GTR+Gamma4 DNA, 128-taxon birth-death tree, 1M sites"),
    "c3": ("wag", 64, 200_000, 1, "WAG amino-acid 20-state, 64-taxon tree, 200k sites"),
    "c4": ("mg94", 32, 100_000, 1, "MG94 codon 61-state, 32-taxon tree, 100k codons"),
    "c1": ("jc69", 16, 1_000, 1, "JC69 DNA, 16-taxon tree, 1k sites"),
}


def algorithmic_bytes_per_site(n, ncat, ntaxa):
    """B_site = 8*C*(n+1)*(2T-4) + T + 8 (level-scheduled model)."""
    return 8 * ncat * (n + 1) * (2 * ntaxa - 4) + ntaxa + 8


def algorithmic_flops_per_site(n, ncat, ntaxa):
    """F_site = C*(2n^2*(T-2) + n*(T-1)) (tip children are look-ups)."""
    return ncat * (2 * n * n * (ntaxa - 2) + n * (ntaxa - 1))


FP64_DMMA_PEAK_TFLOPS = 36.8      # profiles/r01_fp64_peak_dmma_shapes.jsonl
FP64_DFMA_PEAK_TFLOPS = 34.8      # profiles/r01_fp64_peak.json (measured DFMA rate; algorithmic flops exclude rescaling)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def build_case(cfg, nsites, seed, site_seed=None):
    import cases
    model, ntaxa, _, ncat, _ = CONFIGS[cfg]
    return cases.make_case(cfg, model, ntaxa, nsites, ncat=ncat, seed=seed, site_seed=site_seed)


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


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's pruning, all host cores, bounded sample."""
    if rank != 0:
        return
    from oracle import c_oracle as C
    cfg = args.config
    model, ntaxa, full_sites, ncat, desc = CONFIGS[cfg]
    cores = os.cpu_count() or 1
    # size the per-step sample for roughly 3 s of CPU work per step
    probe = build_case(cfg, 4096, seed=1)
    P = probe.host_P()
    f = probe.flat
    t0 = time.perf_counter()
    C.pruning(probe.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P, probe.pi, tips=probe.tips,
              weights=probe.weights, nthreads=cores)
    per_site = (time.perf_counter() - t0) / 4096
    budget = 180.0 / max(1, args.steps + args.warmup)
    sample = int(min(full_sites, max(4096, min(3.0, budget) / per_site)))
    case = build_case(cfg, sample, seed=2)
    P = case.host_P()
    f = case.flat

    def step():
        ps = C.pruning(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi, tips=case.tips,
                       weights=case.weights, nthreads=cores)
        return C.sum_sequential(ps)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    ninternal = f.ninternal
    value = sample * ninternal / dt
    sample_desc = "%d of %d sites per step, same tree/model shape" % (sample, full_sites)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "sample": sample_desc,
                   "note": "CPU restatement of the reference algorithm (oracle/pruning_oracle.c, OpenMP over "
                           "sites, P(t) precomputed per branch), not the OCaml/lacaml binary"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline_block(cfg, seconds=12.0):
    """Oracle port timed on this box's host cores on a bounded sample (rank 0, N = 1 only)."""
    from oracle import c_oracle as C
    cores = os.cpu_count() or 1
    _, _, full_sites, _, _ = CONFIGS[cfg]
    probe = build_case(cfg, 4096, seed=1)
    P = probe.host_P()
    f = probe.flat

    def run(case, P):
        t0 = time.perf_counter()
        C.pruning(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi, tips=case.tips,
                  weights=case.weights, nthreads=cores)
        return time.perf_counter() - t0
    run(probe, P)
    per_site = run(probe, P) / 4096
    sample = int(min(full_sites, max(4096, seconds / per_site)))
    case = build_case(cfg, sample, seed=2)
    P = case.host_P()
    f = case.flat
    dt = run(case, P)
    return {"value": sample * f.ninternal / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d of %d sites, one pass (%.1f s), same tree/model shape; CPU restatement of the "
                      "reference algorithm, not the OCaml binary" % (sample, full_sites, dt)}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from phylogenetics_b200.engine import LikelihoodContext

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = args.config
    model, ntaxa, full_sites, ncat, desc = CONFIGS[cfg]
    nsites = args.sites or full_sites
    # one problem, sharded by site columns: every rank has the SAME tree and model and simulates its own columns
    case = build_case(cfg, nsites, seed=100, site_seed=rank)
    f = case.flat
    n = case.n
    ninternal = f.ninternal

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = LikelihoodContext(n, nsites, f.nleaves, ncat=ncat, device=local_rank, level_kernels=(args.path == "level"))
    ctx.set_stream(stream.cuda_stream)
    ctx.set_tree(f, case.bl)
    ctx.set_categories(case.rates, case.weights)
    ctx.set_rate_matrix(case.Q, case.pi)
    ctx.set_root_frequencies(case.pi)
    tips_pinned = torch.from_numpy(case.tips).pin_memory()
    ctx.set_tips_ptr(tips_pinned.data_ptr(), nsites)
    ctx.set_timing(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    total_t = None

    def step(mid=None):
        nonlocal total_t
        ctx.set_branch_lengths(case.bl)      # new branch lengths every evaluation: P(t) is rebuilt (kernel K1)
        ctx.enqueue()
        if mid is not None:
            mid.record(stream)               # local work of the step ends here; what follows is the collective
        if world > 1:
            if total_t is None:
                total_t = torch.as_tensor(ctx.total_as_cuda_array(), device="cuda")
            dist.all_reduce(total_t)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness gate on a slice of this rank's shard before timing anything
    ctx.enqueue()
    total0, per_site0 = ctx.fetch(per_site=True)
    from oracle import c_oracle as C
    Pdev = ctx.get_pmatrices()
    chk = min(nsites, 20000)
    expect = C.pruning(n, f.root, f.child_ptr, f.child_idx, f.tip_row, Pdev, case.pi,
                       tips=np.ascontiguousarray(case.tips[:, :chk]), weights=case.weights)
    rel = float(np.max(np.abs(per_site0[:chk] - expect) / np.abs(expect)))
    if not rel < 1e-9:
        raise SystemExit("parity gate failed: per-site rel err %.3e" % rel)
    tot_rel = abs(total0 - float(np.sum(per_site0))) / abs(total0)

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = ctx.kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    prune_ms = []
    barrier()
    t_wall0 = time.time()
    # The K steps are queued without host synchronisation (a host sync per step would put launch latency
    # and, for N > 1, the Python-side collective overhead inside every event pair); each step keeps its own
    # event pair, the L2 flush sits between the pairs, and the per-phase events of every step are read back
    # from the library's timing ring afterwards.
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for (a, b), m in zip(ev, mids):
        flush.zero_()                    # L2 flush, outside the event pair
        a.record(stream)
        step(m)
        b.record(stream)
    barrier()
    collective_ms = [m.elapsed_time(b) for (a, b), m in zip(ev, mids)]
    local_ms = [a.elapsed_time(m) for (a, b), m in zip(ev, mids)]
    per_rank = None
    if world > 1:                          # where the step time goes on every rank: local work vs waiting in the collective
        mine = torch.tensor([sum(local_ms) / len(local_ms), sum(collective_ms) / len(collective_ms)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"local_ms": [float(x[0]) for x in allr], "collective_ms": [float(x[1]) for x in allr]}
    ring = 64
    prune_ms = [ctx.get_timing(back) for back in range(min(args.steps, ring) - 1, -1, -1)]
    t_wall1 = time.time()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    launches = ctx.kernel_launches - launches0
    t_local = sum(step_ms) / 1e3
    t = torch.tensor([t_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_max = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # end-to-end through the public API with HOST buffers: H2D of the step's tips from pinned
    # memory + evaluation + D2H of the total, every step.  Nucleotide alignments cross PCIe 2-bit packed
    # (pb_loglik_host_packed2, the layout the binding fills from `leaf_state` for 4-state alphabets,
    # INTEGRATION.md); the byte-per-site call (pb_loglik_host) is timed beside it.
    e2e_steps = max(2, min(args.steps, 10))

    def e2e_loop(call):
        for _ in range(2):
            call()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(e2e_steps):
            ctx.set_branch_lengths(case.bl)
            tot = call()
            if world > 1:
                tt = torch.tensor([tot], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt)
                tot = float(tt.item())
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * nsites * ninternal * e2e_steps / float(t.item()), tot

    e2e_bytes_val, tot_bytes = e2e_loop(lambda: ctx.loglik_host_ptr(tips_pinned.data_ptr(), nsites))
    e2e_val, e2e_h2d, e2e_call = e2e_bytes_val, int(case.tips.nbytes), "pb_loglik_host (one byte per site)"
    if n == 4:
        from phylogenetics_b200.engine import pack_tips2
        packed_pinned = torch.from_numpy(pack_tips2(case.tips)).pin_memory()
        e2e_val, tot_packed = e2e_loop(lambda: ctx.loglik_host_packed2_ptr(packed_pinned.data_ptr(), packed_pinned.shape[1]))
        if not abs(tot_packed - tot_bytes) <= 1e-12 * abs(tot_bytes):
            raise SystemExit("packed and byte-per-site evaluations disagree: %r %r" % (tot_packed, tot_bytes))
        e2e_h2d, e2e_call = int(packed_pinned.numel()), "pb_loglik_host_packed2 (2-bit packed nucleotides, 4 sites per byte)"

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        units = world * nsites * ninternal * args.steps
        value = units / t_max
        pm = np.array(prune_ms)                                    # [steps][3] ms
        prune_avg_ms = float(pm[:, 1].mean())
        fused_const = ctx.path_name == "fused4" and n == 4 and ninternal - 1 <= 128
        kernel_name = "fused4c" if fused_const else ctx.path_name
        nlaunch_prune = (ncat if fused_const else 1) if ctx.path_name == "fused4" else max(1, launches // args.steps - 3)
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                tj = json.load(fh)
            if cfg == "c2" and not args.sites and tj.get(kernel_name, {}).get("dram_bytes_per_launch"):
                traffic, traffic_src = tj[kernel_name]["dram_bytes_per_launch"], tj[kernel_name]["source"]
        except (OSError, ValueError):
            pass
        bytes_eval = algorithmic_bytes_per_site(n, ncat, ntaxa) * nsites
        flops_eval = algorithmic_flops_per_site(n, ncat, ntaxa) * nsites
        achieved = bytes_eval / (prune_avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "true_bound": {"fused4": "fp64 pipe + L1 table gathers", "fused_dmma": "fp64 tensor pipe"}.get(ctx.path_name, "hbm"), "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                "peak_kind": "of " + peak_kind,
                "kernel": kernel_name, "launches_per_step": nlaunch_prune,
                "algorithmic_bytes_per_launch": bytes_eval / nlaunch_prune,
                "kernel_ms_per_step": prune_avg_ms,
                "algorithmic_bytes_per_step": bytes_eval,
                "fp64_tflops_achieved": flops_eval / (prune_avg_ms * 1e-3) / 1e12,
                "fp64_peak_tflops": FP64_DFMA_PEAK_TFLOPS,
                "fp64_frac": flops_eval / (prune_avg_ms * 1e-3) / 1e12 / FP64_DFMA_PEAK_TFLOPS,
                "note": ("fused4 keeps conditional likelihoods on chip: it moves ntaxa+8C bytes/site of HBM, "
                         "so against the level-scheduled algorithmic bytes the fraction can exceed 1; its own "
                         "bounds are the FP64 pipe and, since small clades are table look-ups, the L1 gather path "
                         "(profiles/r01n_ncu_full_fused4c_256x4_pow2_clades5.txt)" if ctx.path_name in ("fused4", "fused_dmma") else
                         "level-scheduled: every inner CLV written once and read once")}
        if ctx.path_name == "fused_dmma":
            # dense alphabets on the FP64 tensor cores: the roofline is the DMMA pipe.  Peak = this pool's
            # measured mma.sync f64 throughput (tools/fp64_peak2.cu, profiles/r01_fp64_peak_dmma_shapes.jsonl);
            # MEASURED_PEAKS.json carries no FP64 figure.
            tf = flops_eval / (prune_avg_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": tf, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                    "frac": tf / FP64_DMMA_PEAK_TFLOPS, "traffic": None,
                    "peak_kind": "measured FP64 DMMA (own micro-benchmark; not in MEASURED_PEAKS.json)",
                    "kernel": "fused_dmma3", "launches_per_step": 1, "kernel_ms_per_step": prune_avg_ms,
                    "algorithmic_flops_per_launch": flops_eval,
                    "hbm_gbs_vs_level_scheduled_bytes": achieved,
                    "note": "algorithmic flops (SURVEY 8d: tip children are look-ups, padding not counted) / pruning-phase time; "
                            "ncu DMMA-pipe activity in profiles/r01j_ncu_full_fused_dmma3_c3_c4.txt"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc + (" per GPU" if world > 1 else ""), "sites_per_gpu": nsites,
                       "ntaxa": ntaxa, "nstates": n, "ncat": ncat, "internal_nodes": ninternal,
                       "path": ctx.path_name,
                       "rescaling": "reference trigger (min v <= threshold), exact power-of-two factor; see reference_scale_factor", "l2": "flushed between steps (256 MiB memset outside the event pairs)",
                       "timing": "sum of per-step CUDA-event pairs on the launching stream, max over ranks",
                       "parity_gate": {"per_site_max_rel_err": rel, "sites_checked": chk,
                                       "total_vs_sum_of_sites_rel": tot_rel}},
            "roofline": roof,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d,
                    "d2h_bytes_per_step": 8, "steps": e2e_steps, "call": e2e_call,
                    "byte_per_site": {"value": e2e_bytes_val, "h2d_bytes_per_step": int(case.tips.nbytes), "call": "pb_loglik_host"},
                    "note": "pb_set_branch_lengths + one host-alignment evaluation call per step: tips copied from pinned host memory in column chunks overlapped with the pruning, total read back to the host"},
            "collective_ms_per_step": (sum(collective_ms) / len(collective_ms)) if world > 1 else 0.0,
            "per_rank_ms": per_rank,
            "gpu_launches": int(launches),
            "phase_ms": {"pmat": float(pm[:, 0].mean()), "pruning": prune_avg_ms, "root": float(pm[:, 2].mean())},
            "clocks": clocks,
        }
        if world == 1 and ctx.path_name in ("fused4", "fused_dmma"):
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
                "note": "option fused_pow2=0: rescaling multiplies by 1/max v exactly as SV.shift does; the default "
                        "multiplies by 2^-exponent(max v), which is exact in floating point"}
            ctx.set_option("fused_pow2", 1)
            if n == 4:
                ctx.set_option("fused_spt", 4)
        if world == 1 and n == 4 and ctx.path_name == "fused4":
            # the level-scheduled kernel the north star describes (CLVs materialised in HBM), same inputs
            ctx.close()
            lv = LikelihoodContext(n, nsites, f.nleaves, ncat=ncat, device=local_rank, level_kernels=True)
            lv.set_stream(stream.cuda_stream)
            lv.set_tree(f, case.bl)
            lv.set_categories(case.rates, case.weights)
            lv.set_rate_matrix(case.Q, case.pi)
            lv.set_root_frequencies(case.pi)
            lv.set_tips_ptr(tips_pinned.data_ptr(), nsites)
            lv.set_timing(True)
            ms = []
            for it in range(8):
                flush.zero_()
                lv.set_branch_lengths(case.bl)
                tot_lv = lv.loglik()
                if it >= 3:
                    ms.append(lv.get_timing()[1])
            lv_ms = float(np.mean(ms))
            lv_ach = bytes_eval / (lv_ms * 1e-3) / 1e9
            line["level_scheduled"] = {
                "kernel": lv.path_name, "kernel_ms_per_step": lv_ms, 
                "value": nsites * ninternal / (lv_ms * 1e-3), "unit": UNIT,
                "roofline": {"bound": "hbm", "achieved": lv_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": lv_ach / peaks["hbm_gbs"], "peak_kind": "of " + peak_kind},
                "total_rel_diff_vs_fused": abs(tot_lv - total0) / abs(total0),
                "note": "PB_FLAG_LEVEL_KERNELS: one launch per tree level, every inner CLV written to and read "
                        "from HBM once (the algorithmic-bytes model of SURVEY.md 8d)"}
            lv.close()
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(cfg)
        else:
            line["cpu_baseline"] = None
        emit(line)
    ctx.close()
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
