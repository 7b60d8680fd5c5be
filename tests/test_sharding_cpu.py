"""Site sharding + the scalar combine, world_size 2 over gloo on the CPU.  The per-rank
evaluator is an oracle-backed stand-in (there is no GPU here): what is under test is the host
logic — shard bounds, replication of the model, the all-reduce / ordered reduction, the gather."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from oracle import c_oracle as C
from phylogenetics_b200.sharding import ShardedLikelihood, shard_bounds


def test_shard_bounds_cover_every_site_once():
    for nsites, world in [(10, 3), (1_000_000, 8), (7, 7), (1001, 2)]:
        cuts = [shard_bounds(nsites, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == nsites
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        sizes = [hi - lo for lo, hi in cuts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


class OracleContext:
    """Duck-typed LikelihoodContext whose loglik() is the CPU oracle (test stand-in only)."""

    def __init__(self, nstates, nsites, ntaxa, ncat=1, device=0):
        self.n, self.nsites, self.ncat = nstates, nsites, ncat

    def set_tree(self, flat, bl=None):
        self.flat = flat

    def set_pmatrices(self, P):
        self.P = P

    def set_root_frequencies(self, pi):
        self.pi = pi

    def set_categories(self, rates, weights):
        self.weights = weights

    def set_tips(self, tips):
        assert tips.shape[1] == self.nsites
        self.tips = tips

    def loglik(self, per_site=False):
        f = self.flat
        ps = C.pruning(self.n, f.root, f.child_ptr, f.child_idx, f.tip_row, self.P, self.pi, tips=self.tips,
                       weights=self.weights)
        tot = C.sum_exact(ps)
        return (tot, ps) if per_site else tot


def _worker(rank, world, port, reduce, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = cases.make_case("shard", "gtr", 14, 1001, ncat=4, seed=21)      # same seed: replicated inputs
        P = case.host_P()
        sl = ShardedLikelihood(4, case.nsites, case.flat.nleaves, ncat=4, context_factory=OracleContext,
                               reduce=reduce)
        sl.set_tree(case.flat, case.bl)
        sl.set_pmatrices(P)
        sl.set_categories(case.rates, case.weights)
        sl.set_root_frequencies(case.pi)
        sl.set_tips(case.tips)
        total = sl.loglik()
        total2, full = sl.loglik_per_site(gather=True)
        out[rank] = (total, total2, full, sl.lo, sl.hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("reduce", ["allreduce", "ordered"])
def test_two_rank_sharded_total_matches_unsharded(reduce):
    world = 2
    port = 29500 + (os.getpid() % 2000) + (0 if reduce == "allreduce" else 1)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, reduce, out), nprocs=world, join=True)
    case = cases.make_case("shard", "gtr", 14, 1001, ncat=4, seed=21)
    f = case.flat
    expect = C.pruning(4, f.root, f.child_ptr, f.child_idx, f.tip_row, case.host_P(), case.pi, tips=case.tips,
                       weights=case.weights)
    truth = math.fsum(expect)
    assert out[0][3:] == (0, 500) and out[1][3:] == (500, 1001)
    for r in range(world):
        total, total2, full, _, _ = out[r]
        assert abs(total - truth) <= 1e-12 * abs(truth) and total == total2
        np.testing.assert_array_equal(full, expect)
    assert out[0][0] == out[1][0]           # every rank holds the same combined value
