"""BASELINE config 5 in miniature, through the multi-process path: two ranks (torch.distributed, gloo for the
scalar because both ranks share the one GPU of the test box; NCCL is what bench.py uses across GPUs), each
with a real CUDA context on its shard of the site columns; Nelder_mead.minimize drives the summed
log-likelihood.  Every rank sees the same objective values, hence takes the same steps, and the optimum
equals the single-context one."""
import math
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from phylogenetics_b200 import models, nelder_mead
from phylogenetics_b200.engine import LikelihoodContext
from phylogenetics_b200.sharding import ShardedLikelihood

pytestmark = pytest.mark.gpu


def objective_factory(sl_or_ctx, case, log=None):
    def neg_loglik(x):
        kappa, scale = math.exp(x[0]), math.exp(x[1])
        sl_or_ctx.set_eigensystem(*models.Site_evolution_model.K80.rate_matrix_reduction(kappa))
        sl_or_ctx.set_branch_lengths(case.bl * scale)
        v = -sl_or_ctx.loglik()
        if log is not None:
            log.append(v)
        return v
    return neg_loglik


def sampler():
    start = np.random.default_rng(1)
    return lambda: [start.uniform(-1.0, 2.0), start.uniform(-1.0, 1.0)]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = cases.make_case("shardopt", "k80", 20, 30001, seed=31)          # same seed: replicated inputs
        sl = ShardedLikelihood(4, case.nsites, case.flat.nleaves, device=0, reduce="ordered")
        sl.set_tree(case.flat, case.bl)
        sl.set_root_frequencies(case.pi)
        sl.set_tips(case.tips)                                                  # keeps its own columns
        values = []
        best, x, its = nelder_mead.minimize(objective_factory(sl, case, values), sampler(), tol=1e-6, maxit=60)
        total, full = None, None
        sl.set_eigensystem(*models.Site_evolution_model.K80.rate_matrix_reduction(2.0))
        sl.set_branch_lengths(case.bl)
        total, full = sl.loglik_per_site(gather=True)
        out[rank] = (best, x, its, values, total, full, sl.lo, sl.hi, sl.ctx.kernel_launches)
        sl.close()
    finally:
        dist.destroy_process_group()


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def test_two_ranks_optimise_like_one_context():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
            assert p.exitcode == 0
        res = [out[r] for r in range(world)]
    # every rank saw the same objective values and took the same path
    assert res[0][3] == res[1][3] and res[0][1] == res[1][1] and res[0][2] == res[1][2]
    assert (res[0][6], res[0][7], res[1][6], res[1][7]) == (0, 15000, 15000, 30001)
    assert all(r[8] > 0 for r in res)
    # single context on the whole alignment
    case = cases.make_case("shardopt", "k80", 20, 30001, seed=31)
    with LikelihoodContext(4, case.nsites, case.flat.nleaves) as one:
        one.set_tree(case.flat, case.bl)
        one.set_root_frequencies(case.pi)
        one.set_tips(case.tips)
        values = []
        best, x, its = nelder_mead.minimize(objective_factory(one, case, values), sampler(), tol=1e-6, maxit=60)
        one.set_eigensystem(*models.Site_evolution_model.K80.rate_matrix_reduction(2.0))
        one.set_branch_lengths(case.bl)
        total, per_site = one.loglik(per_site=True)
    # per-site values are identical (the same kernels on the same columns); totals agree to rounding of the sum
    assert np.array_equal(res[0][5], per_site) and np.array_equal(res[1][5], per_site)
    assert abs(res[0][4] - total) <= 1e-12 * abs(total) and res[0][4] == res[1][4]
    # the optimisation lands on the same optimum (objective values differ only by summation order)
    assert abs(res[0][0] - best) <= 1e-9 * abs(best)
    assert np.allclose(res[0][1], x, rtol=0, atol=1e-3)
