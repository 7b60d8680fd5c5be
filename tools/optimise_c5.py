"""BASELINE config 5: Nelder-Mead over GTR exchangeabilities, Gamma shape and tree scale on an alignment
sharded by site columns across the GPUs of one box (torchrun, one process per GPU, NCCL all-reduce of the
log-likelihood per evaluation), with a fixed evaluation budget.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
        tools/optimise_c5.py --sites-per-gpu 1250000 --evaluations 200

Rank 0 prints one JSON line: evaluations/s, aggregate site-node updates/s inside the loop, the log-likelihood
at the start and at the end."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites-per-gpu", type=int, default=1_250_000)
    ap.add_argument("--evaluations", type=int, default=200)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import cases
    from phylogenetics_b200 import models, nelder_mead
    from phylogenetics_b200.engine import LikelihoodContext

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ncat, ntaxa, nsites = 4, 128, args.sites_per_gpu
    case = cases.make_case("c5", "gtr", ntaxa, nsites, ncat=ncat, seed=100, site_seed=rank)   # one tree, own columns
    f = case.flat
    ninternal = f.nnodes - f.nleaves
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = LikelihoodContext(4, nsites, f.nleaves, ncat=ncat, device=local)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_tree(f, case.bl)
    ctx.set_root_frequencies(case.pi)
    ctx.set_tips(case.tips)
    total_t = None
    count = [0]

    def neg_loglik(x):
        """x = log of 5 exchangeabilities (the sixth fixed to 1), log alpha, log tree scale."""
        nonlocal total_t
        ex = np.exp(np.concatenate([x[:5], [0.0]]))
        S = np.zeros((4, 4))
        S[np.triu_indices(4, 1)] = ex
        S = S + S.T
        Q = S * case.pi[None, :]
        np.fill_diagonal(Q, 0.0)
        Q -= np.diag(Q.sum(1))
        Q /= -(case.pi * np.diag(Q)).sum()                  # Rate_matrix.scaled_rate_matrix
        rates, weights = models.gamma_category_rates(math.exp(x[5]), ncat)
        ctx.set_rate_matrix(Q, case.pi)                     # host Jacobi on the symmetrised 4x4, device P(t)
        ctx.set_categories(rates, weights)
        ctx.set_branch_lengths(case.bl * math.exp(x[6]))
        ctx.enqueue()
        count[0] += 1
        if world > 1:
            if total_t is None:
                total_t = torch.as_tensor(ctx.total_as_cuda_array(), device="cuda")
            dist.all_reduce(total_t)                        # in place on the device-resident total
            return -float(total_t.item())
        return -ctx.fetch()

    class Budget(Exception):
        pass

    def budgeted(x):
        if count[0] >= args.evaluations:
            raise Budget()
        return neg_loglik(x)

    truth = np.concatenate([np.log(case.Q[np.triu_indices(4, 1)] / case.pi[np.triu_indices(4, 1)[1]])[:5] -
                            math.log(case.Q[2, 3] / case.pi[3]), [math.log(0.5), 0.0]])
    start_rng = np.random.default_rng(3)                    # same on every rank: identical simplex
    first = neg_loglik(truth + 0.3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    count[0] = 0
    values = []
    t0 = time.time()
    try:
        nelder_mead.minimize(lambda x: values.append(budgeted(x)) or values[-1],
                             lambda: list(truth + start_rng.uniform(-0.3, 0.3, size=7)), tol=1e-12, maxit=10**9)
    except Budget:
        pass
    torch.cuda.synchronize()
    dt = time.time() - t0
    at_truth = neg_loglik(truth)
    if rank == 0:
        print(json.dumps({
            "config": "c5: GTR+G4, %d taxa, %d sites over %d GPU(s), Nelder-Mead over 5 exchangeabilities + alpha + tree scale"
                      % (ntaxa, nsites * world, world),
            "n_gpus": world, "evaluations": count[0] - 1, "seconds": dt, "evaluations_per_s": (count[0] - 1) / dt,
            "ms_per_evaluation": dt / max(count[0] - 1, 1) * 1e3,
            "site_node_updates_per_s": world * nsites * ninternal * (count[0] - 1) / dt,
            "neg_loglik_first": first, "neg_loglik_best": min(values), "neg_loglik_at_generating_parameters": at_truth,
            "note": "wall clock around the whole optimisation loop: host simplex logic, host eigen-decomposition, "
                    "parameter upload, P(t) + pruning + reduction on every GPU, NCCL all-reduce, scalar read-back"}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
