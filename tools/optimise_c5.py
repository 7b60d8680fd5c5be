"""BASELINE config 5: Nelder-Mead over the 254 branch lengths (log scale), 5 GTR exchangeabilities, 3 free base
frequencies and the Gamma shape — 263 parameters — of a GTR+G4 model on a 128-taxon alignment sharded by site
columns across the GPUs of one box, with a fixed evaluation budget.  One process per GPU under
torchrun; per evaluation every rank uploads the parameters (O(branches) scalars), rebuilds P(t), prunes its shard
and the scalar log-likelihood is all-reduced over NCCL in place on the device.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
        tools/optimise_c5.py --sites-per-gpu 1250000 --evaluations 400

Rank 0 prints one JSON line.  `c5_loop` is also what bench.py runs as its `c5_nelder_mead` block at N > 1."""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def c5_loop(case, ctx, world, dist, evaluations, graph=True):
    """The optimisation loop on an existing context (tree, tips and stream already set).  Returns a dict."""
    import torch
    from phylogenetics_b200 import models, nelder_mead
    f = case.flat
    nsites, ncat = case.nsites, case.ncat
    ninternal = f.nnodes - f.nleaves
    branches = np.array([v for v in range(f.nnodes) if v != f.root])
    nb = len(branches)                                                      # 254 at 128 taxa
    total_t = None
    count = [0]
    if graph:
        ctx.set_option("graph", 1)                                           # one captured evaluation, replayed

    def unpack(x):
        bl = np.zeros(f.nnodes)
        bl[branches] = np.exp(x[:nb])
        ex = np.exp(np.concatenate([x[nb:nb + 5], [0.0]]))                  # the sixth exchangeability fixed to 1
        w = np.exp(np.concatenate([x[nb + 5:nb + 8], [0.0]]))
        pi = w / w.sum()                                                    # 3 free frequencies
        alpha = math.exp(x[nb + 8])
        return bl, ex, pi, alpha

    def neg_loglik(x):
        nonlocal total_t
        bl, ex, pi, alpha = unpack(np.asarray(x))
        S = np.zeros((4, 4))
        S[np.triu_indices(4, 1)] = ex
        S = S + S.T
        Q = S * pi[None, :]
        np.fill_diagonal(Q, 0.0)
        Q -= np.diag(Q.sum(1))
        Q /= -(pi * np.diag(Q)).sum()                                       # Rate_matrix.scaled_rate_matrix
        rates, weights = models.gamma_category_rates(alpha, ncat)
        ctx.set_rate_matrix(Q, pi)                                          # host Jacobi on the symmetrised 4x4, device P(t)
        ctx.set_root_frequencies(pi)
        ctx.set_categories(rates, weights)
        ctx.set_branch_lengths(bl)
        ctx.enqueue()
        count[0] += 1
        if world > 1:
            if total_t is None:
                total_t = torch.as_tensor(ctx.total_as_cuda_array(), device="cuda")
            dist.all_reduce(total_t)                                        # in place on the device-resident total
            return -float(total_t.item())
        return -ctx.fetch()

    class Budget(Exception):
        pass

    def budgeted(x):
        if count[0] >= evaluations:
            raise Budget()
        return neg_loglik(x)

    tri = np.triu_indices(4, 1)
    ex_true = case.Q[tri] / case.pi[tri[1]]
    truth = np.concatenate([np.log(case.bl[branches]), np.log(ex_true[:5] / ex_true[5]),
                            np.log(case.pi[:3] / case.pi[3]), [math.log(0.5)]])
    start_rng = np.random.default_rng(3)                                    # same on every rank: identical simplex
    first = neg_loglik(truth + 0.05)
    at_truth = neg_loglik(truth)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    count[0] = 0
    values = []
    t0 = time.time()
    try:
        nelder_mead.minimize_np(lambda x: values.append(budgeted(x)) or values[-1],
                                lambda: truth + start_rng.uniform(-0.05, 0.05, size=len(truth)), tol=1e-12, maxit=10 ** 9)
    except Budget:
        pass
    torch.cuda.synchronize()
    dt = time.time() - t0
    if graph:
        ctx.set_option("graph", 0)
    n_eval = count[0]
    return {
        "config": "c5: GTR+G4, %d taxa, %d sites over %d GPU(s); Nelder-Mead over %d log branch lengths + 5 exchangeabilities "
                  "+ 3 free frequencies + alpha = %d parameters" % (f.nleaves, nsites * world, world, nb, len(truth)),
        "n_gpus": world, "parameters": int(len(truth)), "evaluations": n_eval, "seconds": dt,
        "ms_per_evaluation": dt / max(n_eval, 1) * 1e3, "evaluations_per_s": n_eval / dt,
        "value": world * nsites * ninternal * n_eval / dt, "unit": "site-node updates/s",
        "cuda_graph": bool(graph),
        "neg_loglik_first": first, "neg_loglik_best": min(values), "neg_loglik_at_generating_parameters": at_truth,
        "note": "wall clock around the whole optimisation loop: host simplex logic (numpy mirror of lib/nelder_mead.ml:53-132), "
                "4x4 eigen-decomposition and Gamma quantiles on the host, parameter upload, P(t) + pruning + reduction on "
                "every GPU, NCCL all-reduce, scalar read-back; a simplex of %d vertices takes %d evaluations to initialise"
                % (len(truth) + 1, len(truth) + 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites-per-gpu", type=int, default=1_250_000)
    ap.add_argument("--evaluations", type=int, default=400)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import cases
    from phylogenetics_b200.engine import LikelihoodContext

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ncat, ntaxa, nsites = 4, 128, args.sites_per_gpu
    case = cases.make_case("c5", "gtr", ntaxa, nsites, ncat=ncat, seed=100, site_seed=rank)   # one tree, own columns
    f = case.flat
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = LikelihoodContext(4, nsites, f.nleaves, ncat=ncat, device=local)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_tree(f, case.bl)
    ctx.set_root_frequencies(case.pi)
    ctx.set_tips(case.tips)
    out = c5_loop(case, ctx, world, dist, args.evaluations, graph=not args.no_graph)
    if rank == 0:
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
