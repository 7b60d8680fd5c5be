"""ctypes front end of oracle/pruning_oracle.c (TEST INFRASTRUCTURE ONLY).

The flattened tree is the same CSR description the C-ABI takes
(include/phylo_b200.h): child_ptr[nnodes+1], child_idx[], tip_row[nnodes]
(-1 for internal nodes); branch b is identified by the node at its lower end.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import phylo_ctmc_ref as ref

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """Compile the C oracle in-tree (oracle/_build/liboracle.so)."""
    if force or not os.path.exists(_SO) or \
            os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "pruning_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        i32p = ctypes.POINTER(ctypes.c_int32)
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_pruning.restype = ctypes.c_int
        L.oracle_pruning.argtypes = [ctypes.c_int] * 4 + [i32p, i32p, i32p, dp, dp,
                                                          ctypes.c_void_p, ctypes.c_void_p,
                                                          ctypes.c_int64, ctypes.c_int64, dp,
                                                          ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                                          dp, dp]
        L.oracle_sum_sequential.restype = ctypes.c_double
        L.oracle_sum_sequential.argtypes = [dp, ctypes.c_int64]
        L.oracle_sum_exact.restype = ctypes.c_double
        L.oracle_sum_exact.argtypes = [dp, ctypes.c_int64]
        L.oracle_conditional_likelihoods.restype = ctypes.c_int
        L.oracle_conditional_likelihoods.argtypes = [ctypes.c_int] * 3 + [i32p, i32p, i32p, dp,
                                                                          ctypes.c_void_p, ctypes.c_void_p,
                                                                          ctypes.c_int64, ctypes.c_int64,
                                                                          ctypes.c_double, dp, dp]
        _lib = L
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def _colmajor(P, ncat, nnodes, n):
    """[ncat][nnodes][n][n] (row, col) -> the oracle's column-major storage."""
    P = np.asarray(P, dtype=np.float64).reshape(ncat, nnodes, n, n)
    return np.ascontiguousarray(np.swapaxes(P, 2, 3))


def pruning(nstates, root, child_ptr, child_idx, tip_row, P, pi, tips=None, masks=None,
            weights=None, threshold=1e-6, root_mode=1, nthreads=0, want_cat=False):
    """Per-site log-likelihoods.

    P: [ncat][nnodes][n][n] (or [nnodes][n][n]) indexed [row=parent state, col=child state].
    tips: uint8 [ntaxa][nsites]; masks: uint64 [ntaxa][nsites] state sets.
    """
    nnodes = len(tip_row)
    P = np.asarray(P, dtype=np.float64)
    if P.ndim == 3:
        P = P[None]
    ncat = P.shape[0]
    Pc = _colmajor(P, ncat, nnodes, nstates)
    w = np.ascontiguousarray(np.ones(1) if weights is None else weights, dtype=np.float64)
    pi = np.ascontiguousarray(pi, dtype=np.float64)
    cp, ci, tr = _i32(child_ptr), _i32(child_idx), _i32(tip_row)
    tp = mp = None
    if masks is not None:
        masks = np.ascontiguousarray(masks, dtype=np.uint64)
        ld, nsites = masks.shape[1], masks.shape[1]
        mp = masks.ctypes.data
    else:
        tips = np.ascontiguousarray(tips, dtype=np.uint8)
        ld, nsites = tips.shape[1], tips.shape[1]
        tp = tips.ctypes.data
    out = np.empty(nsites)
    out_cat = np.empty((ncat, nsites)) if want_cat else None
    rc = lib().oracle_pruning(nstates, ncat, nnodes, int(root),
                              _ptr(cp, ctypes.c_int32), _ptr(ci, ctypes.c_int32),
                              _ptr(tr, ctypes.c_int32), _ptr(Pc, ctypes.c_double),
                              _ptr(w, ctypes.c_double), tp, mp, ld, nsites,
                              _ptr(pi, ctypes.c_double), threshold, root_mode, nthreads,
                              _ptr(out, ctypes.c_double),
                              _ptr(out_cat, ctypes.c_double) if want_cat else None)
    if rc != 0:
        raise ValueError("oracle_pruning: bad arguments")
    return (out, out_cat) if want_cat else out


def sum_sequential(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return lib().oracle_sum_sequential(_ptr(x, ctypes.c_double), x.size)


def sum_exact(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return lib().oracle_sum_exact(_ptr(x, ctypes.c_double), x.size)


def conditional_likelihoods(nstates, root, child_ptr, child_idx, tip_row, P, site,
                            tips=None, masks=None, threshold=1e-6):
    """(clv [nnodes][n], carry [nnodes]) for one site and one category."""
    nnodes = len(tip_row)
    Pc = _colmajor(P, 1, nnodes, nstates)
    cp, ci, tr = _i32(child_ptr), _i32(child_idx), _i32(tip_row)
    tp = mp = None
    if masks is not None:
        masks = np.ascontiguousarray(masks, dtype=np.uint64)
        ld, mp = masks.shape[1], masks.ctypes.data
    else:
        tips = np.ascontiguousarray(tips, dtype=np.uint8)
        ld, tp = tips.shape[1], tips.ctypes.data
    clv = np.full((nnodes, nstates), np.nan)
    car = np.full(nnodes, np.nan)
    rc = lib().oracle_conditional_likelihoods(nstates, nnodes, int(root),
                                              _ptr(cp, ctypes.c_int32), _ptr(ci, ctypes.c_int32),
                                              _ptr(tr, ctypes.c_int32), _ptr(Pc, ctypes.c_double),
                                              tp, mp, ld, site, threshold,
                                              _ptr(clv, ctypes.c_double), _ptr(car, ctypes.c_double))
    if rc != 0:
        raise ValueError("oracle_conditional_likelihoods: bad arguments")
    return clv, car


def tree_from_csr(root, child_ptr, child_idx, tip_row):
    """Rebuild an oracle Tree.t (phylo_ctmc_ref) from the CSR arrays.

    Leaf data = tip row; branch data = the child's node id (index into P).
    """
    def build_node(v):
        c0, c1 = child_ptr[v], child_ptr[v + 1]
        if c0 == c1:
            return ref.Leaf(int(tip_row[v]))
        return ref.Node(int(v), [ref.Branch(int(child_idx[c]), build_node(int(child_idx[c])))
                                 for c in range(c0, c1)])
    return build_node(int(root))
