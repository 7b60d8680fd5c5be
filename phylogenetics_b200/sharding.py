"""Site sharding across the GPUs of one box: one process per GPU (torchrun), every rank
evaluates a contiguous block of alignment columns, and only the scalar log-likelihood is
combined (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests).

Sites are independent given the tree and the model (the reference sums per-site values,
lib/felsenstein.ml:73-75), so there is no exchange during pruning: tree, branch lengths,
eigensystem and category rates are replicated, P(t) is rebuilt redundantly on every rank.
"""
import numpy as np


def shard_bounds(nsites, world, rank):
    """Columns [lo, hi) of rank `rank`: contiguous, sizes differ by at most one, every site once."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    return (nsites * rank) // world, (nsites * (rank + 1)) // world


class ShardedLikelihood:
    """One rank's share of a site-sharded likelihood.

    context_factory(nstates, nsites_local, ntaxa, ncat=..., device=...) builds the evaluator of
    the local shard (LikelihoodContext on a GPU; the CPU tests inject an oracle-backed fake).
    """

    def __init__(self, nstates, nsites_total, ntaxa, ncat=1, rank=None, world=None, device=None,
                 context_factory=None, reduce="allreduce", **ctx_kwargs):
        import torch.distributed as dist
        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.rank = rank if rank is not None else (self.dist.get_rank() if self.dist else 0)
        self.world = world if world is not None else (self.dist.get_world_size() if self.dist else 1)
        self.lo, self.hi = shard_bounds(nsites_total, self.world, self.rank)
        self.nsites_total = nsites_total
        if self.hi <= self.lo:
            raise ValueError("more ranks than sites")
        if context_factory is None:
            from .engine import LikelihoodContext as context_factory
        if device is None:
            device = 0
        self.device = device
        self.reduce = reduce
        self.ctx = context_factory(nstates, self.hi - self.lo, ntaxa, ncat=ncat, device=device, **ctx_kwargs)

    def __getattr__(self, name):
        # model / tree setters are replicated: forward them to the local context
        if name.startswith("set_") or name in ("get_pmatrices", "path_name", "kernel_launches", "close"):
            return getattr(self.ctx, name)
        raise AttributeError(name)

    def set_tips(self, tips_full):
        """tips_full: uint8 [ntaxa][nsites_total]; this rank keeps its own columns."""
        tips_full = np.asarray(tips_full)
        if tips_full.shape[1] != self.nsites_total:
            raise ValueError("expected the full alignment; use set_tips_shard for a pre-cut shard")
        self.ctx.set_tips(np.ascontiguousarray(tips_full[:, self.lo:self.hi]))

    def set_tips_shard(self, tips_shard):
        self.ctx.set_tips(tips_shard)

    def _backend_device(self):
        import torch
        if self.dist and self.dist.get_backend() == "nccl":
            return torch.device("cuda", self.device)
        return torch.device("cpu")

    def combine(self, local_total):
        """Sum of the per-rank totals, identical on every rank."""
        if not self.dist or self.world == 1:
            return float(local_total)
        import torch
        dev = self._backend_device()
        if self.reduce == "ordered":
            # all-gather + sum in rank order: bit-reproducible whatever the collective algorithm
            mine = torch.tensor([local_total], dtype=torch.float64, device=dev)
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            self.dist.all_gather(parts, mine)
            acc = 0.0
            for p in parts:
                acc = acc + float(p.item())
            return acc
        t = torch.tensor([local_total], dtype=torch.float64, device=dev)
        self.dist.all_reduce(t)
        return float(t.item())

    def loglik(self):
        """Total log-likelihood of the whole alignment (all ranks return the same value)."""
        return self.combine(self.ctx.loglik())

    def loglik_per_site(self, gather=False):
        """(total, per-site values).  Per-site values stay sharded unless gather=True, in which case
        every rank receives the full [nsites_total] vector."""
        total, local = self.ctx.loglik(per_site=True)
        total = self.combine(total)
        if not gather or not self.dist or self.world == 1:
            return total, local
        import torch
        dev = self._backend_device()
        sizes = [shard_bounds(self.nsites_total, self.world, r) for r in range(self.world)]
        width = max(hi - lo for lo, hi in sizes)
        buf = torch.zeros(width, dtype=torch.float64, device=dev)
        buf[:len(local)] = torch.from_numpy(local).to(dev)
        parts = [torch.empty_like(buf) for _ in range(self.world)]
        self.dist.all_gather(parts, buf)
        full = np.concatenate([p[:hi - lo].cpu().numpy() for p, (lo, hi) in zip(parts, sizes)])
        return total, full
