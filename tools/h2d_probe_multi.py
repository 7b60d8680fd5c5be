"""Concurrent host-to-device bandwidth of N ranks of one box (torchrun): every rank copies a pinned 32 MB buffer
(the 2-bit packed c2 alignment of one shard) and a pinned 128 MB buffer (one byte per site) to its own GPU, first
alone (the others idle at a barrier), then all at once.  What the end-to-end scaling of bench.py can reach is bounded
by the 'all at once' figure: the copies of all ranks leave the same host memory.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/h2d_probe_multi.py
"""
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(h, d, reps=20):
    st = torch.cuda.current_stream()
    d.copy_(h, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record(st)
    e1.synchronize()
    return h.numel() * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


out = {"n_ranks": world}
for name, nbytes in (("32MB", 32_000_000), ("128MB", 128_000_000)):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    alone = torch.zeros(world, dtype=torch.float64, device="cuda")
    for r in range(world):                      # one rank at a time
        barrier()
        if r == rank:
            st = torch.cuda.current_stream()
            d.copy_(h, non_blocking=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(20):
                d.copy_(h, non_blocking=True)
            e1.record(st)
            e1.synchronize()
            alone[r] = h.numel() * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        barrier()
    together = torch.zeros(world, dtype=torch.float64, device="cuda")
    together[rank] = timed(h, d)
    if world > 1:
        dist.all_reduce(alone)
        dist.all_reduce(together)
    out[name] = {"gbs_per_rank_alone": [round(float(x), 2) for x in alone.tolist()],
                 "gbs_per_rank_all_at_once": [round(float(x), 2) for x in together.tolist()],
                 "aggregate_gbs_all_at_once": round(float(together.sum()), 1)}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
