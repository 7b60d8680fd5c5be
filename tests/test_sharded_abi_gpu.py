"""Multi-GPU INSIDE the C ABI (pb_create_sharded, include/phylo_b200.h): one host thread, one context per device,
the alignment's columns sharded, the total combined by ncclAllReduce(ncclDouble, 1) — SURVEY.md 8(b), 8(e).

On a one-GPU box the NCCL leg is exercised where it can be: ndev = 1 (no communicator) must equal the plain
context bit for bit, and asking for the same device twice must come back as PB_ERR_NCCL, not as a crash.
The two-device tests skip below 2 GPUs."""
import numpy as np
import pytest

import cases
from oracle import c_oracle as C
from phylogenetics_b200 import _lib
from phylogenetics_b200.engine import LikelihoodContext, ShardedContext, pack_tips2

pytestmark = pytest.mark.gpu


def ngpus():
    import torch
    return torch.cuda.device_count()


def plain(case, P):
    with LikelihoodContext(case.n, case.nsites, case.flat.nleaves, ncat=case.ncat) as ctx:
        ctx.set_tree(case.flat, case.bl)
        ctx.set_categories(case.rates, case.weights)
        ctx.set_pmatrices(P)
        ctx.set_tips(case.tips)
        ctx.set_root_frequencies(case.pi)
        return ctx.loglik(per_site=True)


def sharded(case, devices, P=None):
    s = ShardedContext(devices, case.n, case.nsites, case.flat.nleaves, ncat=case.ncat)
    s.set_tree(case.flat, case.bl)
    s.set_categories(case.rates, case.weights)
    if P is None:
        s.set_rate_matrix(case.Q, case.pi)
    else:
        s.set_pmatrices(P)
    s.set_root_frequencies(case.pi)
    return s


@pytest.mark.parametrize("model,ntaxa,nsites,ncat", [("gtr", 40, 30001, 4), ("wag", 12, 1501, 1)])
def test_one_device_equals_the_plain_context(model, ntaxa, nsites, ncat):
    case = cases.make_case("sh1", model, ntaxa, nsites, ncat=ncat, seed=21)
    P = case.host_P()
    total, per_site = plain(case, P)
    with sharded(case, [0], P) as s:
        assert s.ndev == 1 and s.shard_bounds(0) == (0, nsites)
        s.set_tips(case.tips)
        t2, p2 = s.loglik(per_site=True)
        assert t2 == total and np.array_equal(p2, per_site)
        t3, p3 = s.loglik_host_ptr(case.tips.ctypes.data, nsites, per_site=True)
        assert t3 == total and np.array_equal(p3, per_site)
        if case.n == 4:
            packed = pack_tips2(case.tips)
            t4 = s.loglik_host_ptr(packed.ctypes.data, packed.shape[1], packed2=True)
            assert t4 == total


def test_nccl_failures_are_reported_as_pb_err_nccl():
    """One rank per device is NCCL's rule: the same device twice is refused with PB_ERR_NCCL (and a message),
    before any communicator exists."""
    with pytest.raises(_lib.PhyloB200Error) as e:
        ShardedContext([0, 0], 4, 5000, 8, ncat=1)
    assert e.value.code == _lib.PB_ERR_NCCL and "NCCL" in str(e.value)
    with pytest.raises(_lib.PhyloB200InvalidArgument):
        ShardedContext([], 4, 5000, 8)
    with pytest.raises(_lib.PhyloB200Error):
        ShardedContext([ngpus() + 3], 4, 5000, 8)           # no such device: the shard's pb_create fails, reported with its device


@pytest.mark.parametrize("model,ntaxa,nsites,ncat", [("gtr", 64, 100003, 4), ("gtr", 16, 4099, 1), ("wag", 16, 2003, 1)])
def test_two_devices_allreduce(model, ntaxa, nsites, ncat):
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    case = cases.make_case("sh2", model, ntaxa, nsites, ncat=ncat, seed=22)
    P = case.host_P()
    f = case.flat
    expect = C.pruning(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi, tips=case.tips, weights=case.weights)
    total1, per_site1 = plain(case, P)
    with sharded(case, [0, 1], P) as s:
        lo0, hi0 = s.shard_bounds(0)
        lo1, hi1 = s.shard_bounds(1)
        assert (lo0, hi1) == (0, nsites) and hi0 == lo1 and hi0 % 4 == 0
        s.set_tips(case.tips)
        total, per_site = s.loglik(per_site=True)                    # NCCL all-reduce of the two device-resident totals
        assert np.array_equal(per_site, per_site1)                   # sites are independent: same values wherever they run
        assert np.max(np.abs(per_site - expect) / np.abs(expect)) <= 1e-9
        assert abs(total - total1) <= 1e-13 * abs(total1)
        s.set_reduce(_lib.PB_REDUCE_ORDERED)                         # local totals summed in shard order on the host
        total_o = s.loglik()
        assert abs(total_o - total) <= 1e-13 * abs(total)
        assert total_o == s.loglik()                                 # reproducible to the bit
        s.set_reduce(_lib.PB_REDUCE_NCCL)
        t3 = s.loglik_host_ptr(case.tips.ctypes.data, nsites)
        assert t3 == total
    # the optimisation-loop pattern on the device-built matrices: new branch lengths, P(t) rebuilt on both devices
    if case.n == 4:
        with sharded(case, [0, 1]) as s:
            s.set_tips(case.tips)
            a = s.loglik()
            s.set_branch_lengths(case.bl * 1.1)
            b = s.loglik()
            s.set_branch_lengths(case.bl)
            assert b != a and s.loglik() == a


def test_two_devices_eigen_route_and_masks():
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    case = cases.make_case("sh3", "gtr", 24, 9001, ncat=4, seed=23)
    with sharded(case, [1, 0]) as s:                                 # any order of devices
        s.set_tips(case.tips)
        total = s.loglik()
    with LikelihoodContext(4, case.nsites, case.flat.nleaves, ncat=4) as ctx:
        ctx.set_tree(case.flat, case.bl)
        ctx.set_categories(case.rates, case.weights)
        ctx.set_rate_matrix(case.Q, case.pi)
        ctx.set_tips(case.tips)
        ctx.set_root_frequencies(case.pi)
        assert abs(ctx.loglik() - total) <= 1e-13 * abs(total)
