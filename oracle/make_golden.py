"""Writes tests/golden/known_answers.json: high-precision log-likelihoods that certify the
oracle (and, through it, the CUDA path) at the 1e-10 level.

TEST INFRASTRUCTURE ONLY.  Run from the repository root:  python -m oracle.make_golden

The reference's own tests pin this path to 1e-5 at best and the OCaml reference cannot be run
here, so the fixtures are computed independently of every double-precision
code path in this repository: a 60-digit mpmath restatement of Phylo_ctmc.pruning
(lib/phylo_ctmc.ml:83-98 — exact arithmetic needs no rescaling) with P(t) = mp.expm(tQ).
The model parameters (Q, pi, rates) are stored in the fixture as exact doubles.
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

mp.mp.dps = 60


def mp_site_logliks(case):
    f = case.flat
    n = case.n
    Q = mp.matrix(case.Q.tolist())
    P = {}
    for k, r in enumerate(case.rates):
        for v in range(f.nnodes):
            if v != f.root:
                P[k, v] = mp.expm(Q * (mp.mpf(float(case.bl[v])) * mp.mpf(float(r))), method="taylor")
    pi = [mp.mpf(float(x)) for x in case.pi]
    w = [mp.mpf(float(x)) for x in case.weights]
    out = []
    for s in range(case.nsites):
        lik = mp.mpf(0)
        for k in range(case.ncat):
            def clv(v):
                if f.is_leaf(v):
                    st = int(case.tips[f.tip_row[v], s])
                    return [mp.mpf(1) if i == st else mp.mpf(0) for i in range(n)]
                acc = None
                for c in f.children(v):
                    c = int(c)
                    sub = clv(c)
                    term = [sum(P[k, c][i, j] * sub[j] for j in range(n)) for i in range(n)]
                    acc = term if acc is None else [a * b for a, b in zip(acc, term)]
                return acc
            root = clv(f.root)
            lik += w[k] * sum(p * x for p, x in zip(pi, root))
        out.append(mp.log(lik))
    return out


def closed_forms():
    """JC69, two leaves at distances t1, t2 from the root: the pair likelihood is
    1/4 (1/4 + 3/4 e^{-4(t1+t2)/3}) for equal states, 1/4 (1/4 - 1/4 e^{-4(t1+t2)/3}) otherwise."""
    e = mp.exp(-mp.mpf(4) / 3 * (mp.mpf("0.3") + mp.mpf("0.7")))
    return {
        "jc69_pair_same": mp.nstr(mp.log(mp.mpf(1) / 4 * (mp.mpf(1) / 4 + mp.mpf(3) / 4 * e)), 25),
        "jc69_pair_diff": mp.nstr(mp.log(mp.mpf(1) / 4 * (mp.mpf(1) / 4 - mp.mpf(1) / 4 * e)), 25),
    }


def main():
    import cases
    specs = [
        ("jc69_6x12", "jc69", 6, 12, 1),
        ("k80_9x10", "k80", 9, 10, 1),
        ("gtr_g4_10x16", "gtr", 10, 16, 4),
        ("wag_6x6", "wag", 6, 6, 1),
        ("mg94_4x3", "mg94", 4, 3, 1),
    ]
    doc = {"closed_forms": closed_forms(), "cases": []}
    for name, model, ntaxa, nsites, ncat in specs:
        case = cases.make_case(name, model, ntaxa, nsites, ncat=ncat, seed=7)
        ll = mp_site_logliks(case)
        f = case.flat
        doc["cases"].append({
            "name": name, "nstates": case.n, "ncat": case.ncat,
            "root": int(f.root), "child_ptr": f.child_ptr.tolist(), "child_idx": f.child_idx.tolist(),
            "tip_row": f.tip_row.tolist(), "branch_len": [float(x).hex() for x in case.bl],
            "Q": [[float(x).hex() for x in row] for row in case.Q],
            "pi": [float(x).hex() for x in case.pi],
            "rates": [float(x).hex() for x in case.rates], "weights": [float(x).hex() for x in case.weights],
            "tips": case.tips.tolist(),
            "site_loglik": [mp.nstr(x, 25) for x in ll],
            "total_loglik": mp.nstr(sum(ll), 25),
        })
        print(name, mp.nstr(sum(ll), 20))
    # Fixed RNG-free fixture of the reference: lib/mCMC.ml:22,36-38 / tests/test_MCMC.ml:11-12,
    # K80 kappa = 2 on ((T0:0.1,T1:0.1):0.1,(T2:x,T3:0.1):0.1) with columns A,A,A,T.
    doc["mcmc_fixture"] = {}
    for x in ("3.0", "2.5"):
        k = mp.mpf(2)
        Q = mp.matrix(4, 4)
        for i in range(4):
            for j in range(4):
                if i == j:
                    Q[i, j] = -1
                elif {i, j} in ({0, 2}, {1, 3}):
                    Q[i, j] = k / (k + 2)
                else:
                    Q[i, j] = 1 / (k + 2)

        def Pm(t):
            return mp.expm(Q * mp.mpf(t), method="taylor")
        A, Tn = 0, 3

        def col(P, s):
            return [P[i, s] for i in range(4)]
        left = [a * b for a, b in zip(col(Pm("0.1"), A), col(Pm("0.1"), A))]
        right = [a * b for a, b in zip(col(Pm(x), A), col(Pm("0.1"), Tn))]
        P01 = Pm("0.1")
        tl = [sum(P01[i, j] * left[j] for j in range(4)) for i in range(4)]
        tr = [sum(P01[i, j] * right[j] for j in range(4)) for i in range(4)]
        doc["mcmc_fixture"][x] = mp.nstr(mp.log(sum(mp.mpf(1) / 4 * a * b for a, b in zip(tl, tr))), 25)
    out = os.path.join(ROOT, "tests", "golden", "known_answers.json")
    with open(out, "w") as fh:
        json.dump(doc, fh, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
