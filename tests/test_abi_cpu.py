"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares,
and refuses to compute without a GPU (no CPU fallback)."""
import os
import re
import subprocess

import pytest

from phylogenetics_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "phylo_b200.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    declared = header_symbols()
    assert len(declared) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (pb_[a-z0-9_]+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == declared       # the ctypes table covers the whole header


def test_no_torch_or_python_dependency_in_the_abi_library():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out


def test_sm100a_cubin_embedded():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the behaviour on a box without a GPU")
def test_fails_loudly_without_a_gpu():
    from phylogenetics_b200.engine import LikelihoodContext, transition_matrices_eigen
    import numpy as np
    with pytest.raises(_lib.PhyloB200Error) as e:
        LikelihoodContext(4, 100, 8)
    assert e.value.code == _lib.PB_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.PhyloB200Error):
        transition_matrices_eigen(np.eye(4), np.zeros(4), np.eye(4), [0.1])
    # the multi-GPU entry point as well: no device, no silent fallback; its argument errors need no device at all
    from phylogenetics_b200.engine import ShardedContext
    with pytest.raises(_lib.PhyloB200Error) as e:
        ShardedContext([0, 1], 4, 1000, 8)
    assert e.value.code == _lib.PB_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.PhyloB200Error) as e:
        ShardedContext([0, 0], 4, 1000, 8)
    assert e.value.code == _lib.PB_ERR_NCCL
    with pytest.raises(_lib.PhyloB200InvalidArgument):
        ShardedContext([0, 1, 2], 4, 2, 8)                  # more devices than sites


def test_invalid_arguments_are_rejected_before_touching_the_device():
    from phylogenetics_b200.engine import LikelihoodContext
    for bad in [dict(nstates=1, nsites=10, ntaxa=3), dict(nstates=65, nsites=10, ntaxa=3),
                dict(nstates=4, nsites=0, ntaxa=3), dict(nstates=4, nsites=10, ntaxa=0),
                dict(nstates=4, nsites=10, ntaxa=3, ncat=0)]:
        with pytest.raises(_lib.PhyloB200InvalidArgument):
            LikelihoodContext(**bad)


def test_host_eigensystem_helper_reconstructs_q():
    """pb_host_reversible_eigensystem: Q = U diag(lambda) U^-1 for GTR, WAG-sized and codon-sized Q."""
    import numpy as np
    from phylogenetics_b200.engine import host_reversible_eigensystem
    from oracle import models_ref as M
    rng = np.random.default_rng(2)
    for n in (4, 20, 61):
        pi = rng.dirichlet(np.full(n, 5.0))
        ex = M.make_symetric(n, lambda i, j: rng.gamma(1.0, 1.0))
        Q = M.gtr(pi, ex)
        U, lam, Ui = host_reversible_eigensystem(Q, pi)
        assert np.abs(U @ np.diag(lam) @ Ui - Q).max() < 1e-13
        assert np.abs(U @ Ui - np.eye(n)).max() < 1e-13
        assert abs(lam.max()) < 1e-13 and (lam <= 1e-13).all()
    # a non-reversible matrix is refused
    Qn = np.array([[-1., 1., 0., 0.], [0., -1., 1., 0.], [0., 0., -1., 1.], [1., 0., 0., -1.]])
    with pytest.raises(_lib.PhyloB200Error):
        host_reversible_eigensystem(Qn, np.full(4, 0.25))


def test_host_chunk_schedule_invariants():
    """pb_debug_host_schedule (host only): the column chunks of the pipelined host evaluation cover [0, spad) exactly
    once in order, on 1024-site bounds, never more than 64 of them; with a round size they are whole rounds of the
    persistent blocks laid from the end (a ramp of short last chunks), without one they are equal."""
    import ctypes
    import numpy as np

    def schedule(spad, chunks, tail, rnd):
        b = np.zeros(80, np.int64)
        n = ctypes.c_int(0)
        rc = _lib.lib.pb_debug_host_schedule(spad, chunks, tail, rnd, b.ctypes.data, len(b), ctypes.byref(n))
        assert rc == 0, (spad, chunks, tail, rnd)
        return b[:n.value]

    rng = np.random.default_rng(0)
    for _ in range(400):
        spad = int(rng.integers(1, 4000)) * 1024
        chunks = int(rng.integers(1, 65))
        rnd = int(rng.choice([0, 512, 3 * 512, 111 * 512, 444 * 512, 148 * 1024, 7 * 4096]))
        tail = int(rng.choice([-2, -1, 0, 1000, 60000, 250000, 10**7]))
        b = schedule(spad, chunks, tail, rnd)
        assert b[0] == 0 and b[-1] == spad and np.all(np.diff(b) > 0) and np.all(b % 1024 == 0)
        assert 2 <= len(b) <= 65
        if rnd == 0:
            w = np.diff(b)
            assert len(w) <= chunks and np.all(w[:-1] == w[0]) and w[-1] <= w[0]
    # the bench shape: 1M sites, 111 blocks of 512 sites per round (4 categories per launch), 6 chunks
    b = schedule(1000448, 6, 0, 111 * 512)
    w = np.diff(b)
    per_round = 111 * 512
    assert list(w[-3:] // 1024 * 1024) == [2 * per_round // 1024 * 1024, per_round // 1024 * 1024, per_round // 1024 * 1024]
    assert np.all(w[1:-3] == 3 * per_round // 1024 * 1024) and w[0] <= 3 * per_round + per_round // 2
    assert list(np.diff(schedule(1000448, 6, -1, 111 * 512))[1:]) == [3 * per_round // 1024 * 1024] * 5
    # the dense alphabets' ramp-up: 1, 2, 4, ... rounds from the start (c3: 296 CTAs x 11 warps x 8 sites per round)
    w = np.diff(schedule(200704, 6, -2, 296 * 11 * 8))
    assert list(w[:2]) == [25600, 51200] and w.sum() == 200704 and len(w) == 3
    for bad in ((1000, 4, 0, 0), (2048, 0, 0, 0), (2048, 4, 0, -5)):
        n = ctypes.c_int(0)
        out = np.zeros(80, np.int64)
        assert _lib.lib.pb_debug_host_schedule(*bad, out.ctypes.data, 80, ctypes.byref(n)) != 0
