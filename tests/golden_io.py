"""Reads tests/golden/known_answers.json back into problem instances."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Flat:
    pass


def load():
    with open(os.path.join(GOLDEN, "known_answers.json")) as fh:
        return json.load(fh)


def fx(x):
    return float.fromhex(x)


def case_arrays(c):
    f = Flat()
    f.root = c["root"]
    f.child_ptr = np.array(c["child_ptr"], dtype=np.int32)
    f.child_idx = np.array(c["child_idx"], dtype=np.int32)
    f.tip_row = np.array(c["tip_row"], dtype=np.int32)
    f.nnodes = len(f.tip_row)
    f.nleaves = int((f.tip_row >= 0).sum())
    out = dict(flat=f, n=c["nstates"], ncat=c["ncat"],
               bl=np.array([fx(x) for x in c["branch_len"]]),
               Q=np.array([[fx(x) for x in row] for row in c["Q"]]),
               pi=np.array([fx(x) for x in c["pi"]]),
               rates=np.array([fx(x) for x in c["rates"]]),
               weights=np.array([fx(x) for x in c["weights"]]),
               tips=np.array(c["tips"], dtype=np.uint8),
               site_loglik=np.array([float(x) for x in c["site_loglik"]]),
               total=float(c["total_loglik"]))
    return out
