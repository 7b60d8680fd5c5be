"""Thin object wrapper over the C ABI context (include/phylo_b200.h).

All arithmetic happens in libphylo_b200.so on the GPU.  numpy arrays are only
the host buffers that cross the ABI (the role Bigarrays play for OCaml).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import lib, check


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _colmajor(m):
    """n x n matrix indexed [row, col] -> the ABI's column-major buffer."""
    m = np.asarray(m, dtype=np.float64)
    return np.ascontiguousarray(m.T)


class LikelihoodContext:
    """One pb_ctx: a (tree, alignment shard, model) evaluator on one GPU."""

    def __init__(self, nstates, nsites, ntaxa, ncat=1, device=0, keep_clv=False, level_kernels=False):
        flags = (_lib.PB_FLAG_KEEP_CLV if keep_clv else 0) | (_lib.PB_FLAG_LEVEL_KERNELS if level_kernels else 0)
        self.nstates, self.nsites, self.ntaxa, self.ncat = int(nstates), int(nsites), int(ntaxa), int(ncat)
        self._ctx = lib.pb_create(int(device), self.nstates, self.ncat, self.nsites, self.ntaxa, flags)
        if not self._ctx:
            msg = lib.pb_last_error(None).decode()
            if "no CPU fallback" in msg or "sm_100" in msg or "CUDA" in msg:
                raise _lib.PhyloB200Error(_lib.PB_ERR_CUDA, msg)
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, msg)
        self.nnodes = 0
        self.root = -1

    # -- life cycle
    def close(self):
        if getattr(self, "_ctx", None):
            lib.pb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        check(rc, self._ctx)

    # -- configuration
    def set_stream(self, cuda_stream_ptr):
        """Enqueue on this cudaStream_t handle (0 = the CUDA default stream)."""
        self._ck(lib.pb_set_stream(self._ctx, ctypes.c_void_p(int(cuda_stream_ptr))))

    def use_private_stream(self):
        self._ck(lib.pb_use_private_stream(self._ctx))

    def set_mode(self, mode):
        self._ck(lib.pb_set_mode(self._ctx, mode))

    def set_rescaling(self, threshold, shift_at_root):
        self._ck(lib.pb_set_rescaling(self._ctx, float(threshold), int(bool(shift_at_root))))

    def set_tree(self, flat, branch_lengths=None):
        """flat: tree.FlatTree (or any object with nnodes/root/child_ptr/child_idx/tip_row)."""
        cp = np.ascontiguousarray(flat.child_ptr, dtype=np.int32)
        ci = np.ascontiguousarray(flat.child_idx, dtype=np.int32)
        tr = np.ascontiguousarray(flat.tip_row, dtype=np.int32)
        bl = None if branch_lengths is None else _f64(branch_lengths)
        if len(cp) != flat.nnodes + 1 or len(tr) != flat.nnodes or len(ci) != flat.nnodes - 1:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "tree arrays have the wrong length")
        if bl is not None and len(bl) != flat.nnodes:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "branch_lengths must have nnodes entries")
        self._ck(lib.pb_set_tree(self._ctx, int(flat.nnodes), int(flat.root), cp.ctypes.data, ci.ctypes.data,
                                 tr.ctypes.data, None if bl is None else bl.ctypes.data))
        self.nnodes, self.root = int(flat.nnodes), int(flat.root)

    def set_branch_lengths(self, branch_lengths):
        bl = _f64(branch_lengths)
        if len(bl) != self.nnodes:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "branch_lengths must have nnodes entries")
        self._ck(lib.pb_set_branch_lengths(self._ctx, bl.ctypes.data))

    def update_branch_length(self, node, t):
        """Change one branch length; with keep_clv the next evaluation only redoes the path to the root."""
        self._ck(lib.pb_update_branch_length(self._ctx, int(node), float(t)))

    @property
    def nodes_recomputed(self):
        return int(lib.pb_nodes_recomputed(self._ctx))

    def set_tips(self, states):
        """states: uint8 [ntaxa][>=nsites] host array (numpy, or a pinned torch tensor's numpy view)."""
        states = np.asarray(states)
        if states.dtype != np.uint8 or states.ndim != 2 or states.shape[0] != self.ntaxa \
                or states.shape[1] < self.nsites:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "tips must be uint8 [ntaxa][>=nsites]")
        ntaxa, width = states.shape
        if ntaxa == 1 or width == 1:
            # numpy leaves the strides of length-1 axes arbitrary: use a dense copy and ld = width
            states = np.ascontiguousarray(states)
            ld = width
        else:
            if states.strides[1] != 1 or states.strides[0] < width:
                states = np.ascontiguousarray(states)
            ld = states.strides[0]          # row-padded views are passed through as they are
        self._ck(lib.pb_set_tips(self._ctx, states.ctypes.data, ld))

    def set_tips_ptr(self, host_ptr, ld):
        self._ck(lib.pb_set_tips(self._ctx, ctypes.c_void_p(host_ptr), int(ld)))

    def set_tips_device(self, device_ptr, ld):
        self._ck(lib.pb_set_tips_device(self._ctx, ctypes.c_void_p(device_ptr), int(ld)))

    def set_tips_text(self, rows, alphabet):
        """rows: one str/bytes per taxon (all the same length); encoded to states on the device."""
        rows = [r.encode() if isinstance(r, str) else bytes(r) for r in rows]
        width = 3 if alphabet == _lib.PB_ALPHABET_CODON else 1
        if len(rows) != self.ntaxa or any(len(r) != self.nsites * width for r in rows):
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "need ntaxa rows of nsites symbols")
        self._ck(lib.pb_set_tips_text(self._ctx, b"".join(rows), self.nsites * width, int(alphabet)))

    def set_site_weights(self, w):
        if w is None:
            self._ck(lib.pb_set_site_weights(self._ctx, None))
            return
        w = _f64(w)
        if len(w) != self.nsites:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "need one weight per site")
        self._ck(lib.pb_set_site_weights(self._ctx, w.ctypes.data))

    def set_tip_masks(self, masks):
        masks = np.ascontiguousarray(masks, dtype=np.uint64)
        if masks.ndim != 2 or masks.shape[0] != self.ntaxa or masks.shape[1] < self.nsites:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "masks must be uint64 [ntaxa][>=nsites]")
        self._ck(lib.pb_set_tip_masks(self._ctx, masks.ctypes.data, masks.shape[1]))

    def set_root_frequencies(self, pi):
        pi = _f64(pi)
        if len(pi) != self.nstates:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "pi must have nstates entries")
        self._ck(lib.pb_set_root_frequencies(self._ctx, pi.ctypes.data))

    def set_categories(self, rates, weights):
        r, w = _f64(rates), _f64(weights)
        if len(r) != self.ncat or len(w) != self.ncat:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "need ncat rates and weights")
        self._ck(lib.pb_set_categories(self._ctx, r.ctypes.data, w.ctypes.data))

    def set_eigensystem(self, U, lam, Uinv):
        u, ui, l = _colmajor(U), _colmajor(Uinv), _f64(lam)
        self._ck(lib.pb_set_eigensystem(self._ctx, u.ctypes.data, l.ctypes.data, ui.ctypes.data))

    def set_rate_matrix(self, Q, pi):
        q, p = _colmajor(Q), _f64(pi)
        self._ck(lib.pb_set_rate_matrix(self._ctx, q.ctypes.data, p.ctypes.data))

    def set_rate_matrix_general(self, Q):
        q = _colmajor(Q)
        self._ck(lib.pb_set_rate_matrix_general(self._ctx, q.ctypes.data))

    def set_pmatrices(self, P):
        """P: [ncat][nnodes][n][n] (or [nnodes][n][n] when ncat == 1), indexed [row, col]."""
        P = np.asarray(P, dtype=np.float64).reshape(self.ncat, self.nnodes, self.nstates, self.nstates)
        Pc = np.ascontiguousarray(np.swapaxes(P, 2, 3))
        self._ck(lib.pb_set_pmatrices(self._ctx, Pc.ctypes.data))

    # -- evaluation
    def loglik(self, per_site=False):
        total = ctypes.c_double()
        ps = np.empty(self.nsites) if per_site else None
        self._ck(lib.pb_loglik(self._ctx, ctypes.byref(total), None if ps is None else ps.ctypes.data))
        return (total.value, ps) if per_site else total.value

    def loglik_host(self, states, per_site=False):
        """pb_loglik_host: evaluate straight from a host tip matrix (uint8 [ntaxa][>=nsites], ideally
        pinned), overlapping the host-to-device copy with the pruning."""
        states = np.asarray(states)
        if states.dtype != np.uint8 or states.ndim != 2 or states.shape[0] != self.ntaxa \
                or states.shape[1] < self.nsites or (states.shape[1] > 1 and states.strides[1] != 1):
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "tips must be uint8 [ntaxa][>=nsites], site-contiguous")
        ld = states.strides[0] if min(states.shape) > 1 else states.shape[1]
        return self.loglik_host_ptr(states.ctypes.data, ld, per_site)

    def loglik_host_ptr(self, host_ptr, ld, per_site=False):
        total = ctypes.c_double()
        ps = np.empty(self.nsites) if per_site else None
        self._ck(lib.pb_loglik_host(self._ctx, ctypes.c_void_p(host_ptr), int(ld), ctypes.byref(total),
                                    None if ps is None else ps.ctypes.data))
        return (total.value, ps) if per_site else total.value

    def loglik_host_packed2_ptr(self, host_ptr, ld, per_site=False):
        """pb_loglik_host_packed2: the same from a 2-bit packed host alignment (uint8 [ntaxa][>= ceil(nsites/4)],
        see pack_tips2)."""
        total = ctypes.c_double()
        ps = np.empty(self.nsites) if per_site else None
        self._ck(lib.pb_loglik_host_packed2(self._ctx, ctypes.c_void_p(host_ptr), int(ld), ctypes.byref(total),
                                            None if ps is None else ps.ctypes.data))
        return (total.value, ps) if per_site else total.value

    def loglik_host_packed2(self, packed, per_site=False):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        if packed.ndim != 2 or packed.shape[0] != self.ntaxa or packed.shape[1] < (self.nsites + 3) // 4:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "packed tips must be uint8 [ntaxa][>= ceil(nsites/4)]")
        return self.loglik_host_packed2_ptr(packed.ctypes.data, packed.shape[1], per_site)

    def set_tips_packed2(self, packed):
        """pb_set_tips_packed2: upload a 2-bit packed nucleotide alignment (4 sites per byte)."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        if packed.ndim != 2 or packed.shape[0] != self.ntaxa or packed.shape[1] < (self.nsites + 3) // 4:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "packed tips must be uint8 [ntaxa][>= ceil(nsites/4)]")
        self._ck(lib.pb_set_tips_packed2(self._ctx, packed.ctypes.data, packed.shape[1]))

    def enqueue(self):
        self._ck(lib.pb_loglik_enqueue(self._ctx))

    def host_enqueue_ptr(self, host_ptr, ld, packed2=False):
        """pb_loglik_host_enqueue: the non-blocking half of loglik_host[_packed2]_ptr; finish with fetch().  The
        host buffer must stay valid and unchanged until then."""
        self._ck(lib.pb_loglik_host_enqueue(self._ctx, ctypes.c_void_p(host_ptr), int(ld), int(bool(packed2))))

    def fetch(self, per_site=False):
        total = ctypes.c_double()
        ps = np.empty(self.nsites) if per_site else None
        self._ck(lib.pb_loglik_fetch(self._ctx, ctypes.byref(total), None if ps is None else ps.ctypes.data))
        return (total.value, ps) if per_site else total.value

    def result_device_ptrs(self):
        t, p = ctypes.c_void_p(), ctypes.c_void_p()
        self._ck(lib.pb_result_device_ptrs(self._ctx, ctypes.byref(t), ctypes.byref(p)))
        return t.value, p.value

    def get_clv(self, node, cat=0):
        v = np.empty((self.nsites, self.nstates))
        c = np.empty(self.nsites)
        self._ck(lib.pb_get_clv(self._ctx, int(node), int(cat), v.ctypes.data, c.ctypes.data))
        return v, c

    def conditional_simulation(self, seed):
        """One joint draw of every node's state per site given the leaves (pb_conditional_simulation):
        uint8 [nnodes][nsites].  Needs keep_clv=True and a prior evaluation."""
        out = np.empty((self.nnodes, self.nsites), dtype=np.uint8)
        self._ck(lib.pb_conditional_simulation(self._ctx, int(seed) & 0xFFFFFFFFFFFFFFFF, out.ctypes.data, self.nsites))
        return out

    def get_pmatrices(self):
        Pc = np.empty((self.ncat, self.nnodes, self.nstates, self.nstates))
        self._ck(lib.pb_get_pmatrices(self._ctx, Pc.ctypes.data))
        return np.ascontiguousarray(np.swapaxes(Pc, 2, 3))

    def set_option(self, name, value):
        self._ck(lib.pb_set_option(self._ctx, name.encode(), float(value)))

    def set_timing(self, on=True):
        """on = 2: also time the dominant kernel alone (get_kernel_timing)."""
        self._ck(lib.pb_set_timing(self._ctx, 2 if on == 2 else int(bool(on))))

    def get_timing(self, back=0):
        """(pmat_ms, pruning_ms, root_ms) of the last evaluation (or the one `back` evaluations before
        it, up to 63), from CUDA events on the stream."""
        ms = np.empty(3)
        self._ck(lib.pb_get_timing_at(self._ctx, int(back), ms.ctypes.data))
        return tuple(float(x) for x in ms)

    def get_kernel_timing(self):
        """pb_get_kernel_timing: (ms, launches) of the dominant pruning kernel alone in the last evaluation."""
        ms, n = ctypes.c_double(), ctypes.c_int()
        self._ck(lib.pb_get_kernel_timing(self._ctx, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def total_as_cuda_array(self):
        """The device-resident total (1 double) as an object torch.as_tensor can wrap without a
        copy (so a collective can reduce it in place)."""
        ptr, _ = self.result_device_ptrs()

        class _Total:
            __cuda_array_interface__ = {"shape": (1,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        return _Total()

    @property
    def kernel_launches(self):
        return int(lib.pb_kernel_launches(self._ctx))

    def host_trace(self):
        """pb_debug_host_trace: timeline of the last pipelined host evaluation (option "host_trace"): a list of
        (first site, sites, copy-done ms, compute-done ms) per column chunk and the ms at which the call's device work ended."""
        out = np.zeros(2 + 4 * 64)
        self._ck(lib.pb_debug_host_trace(self._ctx, out.ctypes.data, len(out)))
        k = int(out[0])
        return [tuple(float(x) for x in out[1 + 4 * i:5 + 4 * i]) for i in range(k)], float(out[1 + 4 * k])

    @property
    def graph_replays(self):
        """Evaluations that were replays of the captured CUDA graph (option "graph")."""
        return int(lib.pb_graph_replays(self._ctx))

    @property
    def path_name(self):
        return lib.pb_path_name(self._ctx).decode()


class ShardedContext:
    """pb_sharded: ONE host thread drives several GPUs of one box; the alignment's columns are sharded over
    them, tree and model replicated, the total combined by an NCCL all-reduce inside the library
    (include/phylo_b200.h, "multi-GPU").  The torchrun / one-process-per-GPU form is phylogenetics_b200.sharding."""

    def __init__(self, devices, nstates, nsites, ntaxa, ncat=1, keep_clv=False, level_kernels=False):
        flags = (_lib.PB_FLAG_KEEP_CLV if keep_clv else 0) | (_lib.PB_FLAG_LEVEL_KERNELS if level_kernels else 0)
        self.devices = [int(d) for d in devices]
        self.nstates, self.nsites, self.ntaxa, self.ncat = int(nstates), int(nsites), int(ntaxa), int(ncat)
        dev = np.ascontiguousarray(self.devices, dtype=np.int32)
        self._s = lib.pb_create_sharded(len(self.devices), dev.ctypes.data, self.nstates, self.ncat, self.nsites,
                                        self.ntaxa, flags)
        if not self._s:
            msg = lib.pb_sharded_last_error(None).decode()
            code = _lib.PB_ERR_NCCL if ("NCCL" in msg or "nccl" in msg) else \
                (_lib.PB_ERR_INVALID_ARG if "pb_create_sharded" in msg or "must be" in msg else _lib.PB_ERR_CUDA)
            cls = _lib.PhyloB200InvalidArgument if code == _lib.PB_ERR_INVALID_ARG else _lib.PhyloB200Error
            raise cls(code, msg)
        self.nnodes = 0

    def close(self):
        if getattr(self, "_s", None):
            lib.pb_sharded_destroy(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        if rc != _lib.PB_OK:
            msg = lib.pb_sharded_last_error(self._s).decode()
            cls = _lib.PhyloB200InvalidArgument if rc == _lib.PB_ERR_INVALID_ARG else _lib.PhyloB200Error
            raise cls(rc, msg)

    @property
    def ndev(self):
        return int(lib.pb_sharded_ndev(self._s))

    def shard_bounds(self, d):
        lo, hi = ctypes.c_int64(), ctypes.c_int64()
        self._ck(lib.pb_sharded_shard_bounds(self._s, int(d), ctypes.byref(lo), ctypes.byref(hi)))
        return lo.value, hi.value

    def shard_option(self, d, name, value):
        check(lib.pb_set_option(ctypes.c_void_p(lib.pb_sharded_ctx(self._s, int(d))), name.encode(), float(value)))

    def shard_set_tips_ptr(self, d, host_ptr, ld):
        """pb_set_tips on shard d's own context (its columns only: [ntaxa][>= hi - lo])."""
        check(lib.pb_set_tips(ctypes.c_void_p(lib.pb_sharded_ctx(self._s, int(d))), ctypes.c_void_p(host_ptr), int(ld)))

    def shard_path_name(self, d):
        return lib.pb_path_name(ctypes.c_void_p(lib.pb_sharded_ctx(self._s, int(d)))).decode()

    def set_reduce(self, how):
        self._ck(lib.pb_sharded_set_reduce(self._s, int(how)))

    def set_mode(self, mode):
        self._ck(lib.pb_sharded_set_mode(self._s, int(mode)))

    def set_option(self, name, value):
        self._ck(lib.pb_sharded_set_option(self._s, name.encode(), float(value)))

    def set_tree(self, flat, branch_lengths=None):
        cp = np.ascontiguousarray(flat.child_ptr, dtype=np.int32)
        ci = np.ascontiguousarray(flat.child_idx, dtype=np.int32)
        tr = np.ascontiguousarray(flat.tip_row, dtype=np.int32)
        bl = None if branch_lengths is None else _f64(branch_lengths)
        if len(cp) != flat.nnodes + 1 or len(tr) != flat.nnodes or len(ci) != flat.nnodes - 1:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "tree arrays have the wrong length")
        self._ck(lib.pb_sharded_set_tree(self._s, int(flat.nnodes), int(flat.root), cp.ctypes.data, ci.ctypes.data,
                                         tr.ctypes.data, None if bl is None else bl.ctypes.data))
        self.nnodes = int(flat.nnodes)

    def set_branch_lengths(self, branch_lengths):
        bl = _f64(branch_lengths)
        if len(bl) != self.nnodes:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "branch_lengths must have nnodes entries")
        self._ck(lib.pb_sharded_set_branch_lengths(self._s, bl.ctypes.data))

    def set_root_frequencies(self, pi):
        pi = _f64(pi)
        if len(pi) != self.nstates:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "pi must have nstates entries")
        self._ck(lib.pb_sharded_set_root_frequencies(self._s, pi.ctypes.data))

    def set_categories(self, rates, weights):
        r, w = _f64(rates), _f64(weights)
        if len(r) != self.ncat or len(w) != self.ncat:
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "need ncat rates and weights")
        self._ck(lib.pb_sharded_set_categories(self._s, r.ctypes.data, w.ctypes.data))

    def set_rate_matrix(self, Q, pi):
        q, p = _colmajor(Q), _f64(pi)
        self._ck(lib.pb_sharded_set_rate_matrix(self._s, q.ctypes.data, p.ctypes.data))

    def set_rate_matrix_general(self, Q):
        q = _colmajor(Q)
        self._ck(lib.pb_sharded_set_rate_matrix_general(self._s, q.ctypes.data))

    def set_pmatrices(self, P):
        P = np.asarray(P, dtype=np.float64).reshape(self.ncat, self.nnodes, self.nstates, self.nstates)
        Pc = np.ascontiguousarray(np.swapaxes(P, 2, 3))
        self._ck(lib.pb_sharded_set_pmatrices(self._s, Pc.ctypes.data))

    def set_tips(self, states):
        """states: uint8 [ntaxa][nsites], the FULL alignment; every device keeps its own columns."""
        states = np.ascontiguousarray(states, dtype=np.uint8)
        if states.ndim != 2 or states.shape != (self.ntaxa, self.nsites):
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "tips must be uint8 [ntaxa][nsites]")
        self._ck(lib.pb_sharded_set_tips(self._s, states.ctypes.data, states.shape[1]))

    def set_tips_ptr(self, host_ptr, ld):
        self._ck(lib.pb_sharded_set_tips(self._s, ctypes.c_void_p(host_ptr), int(ld)))

    def set_tip_masks(self, masks):
        masks = np.ascontiguousarray(masks, dtype=np.uint64)
        if masks.ndim != 2 or masks.shape != (self.ntaxa, self.nsites):
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "masks must be uint64 [ntaxa][nsites]")
        self._ck(lib.pb_sharded_set_tip_masks(self._s, masks.ctypes.data, masks.shape[1]))

    def loglik(self, per_site=False):
        total = ctypes.c_double()
        ps = np.empty(self.nsites) if per_site else None
        self._ck(lib.pb_sharded_loglik(self._s, ctypes.byref(total), None if ps is None else ps.ctypes.data))
        return (total.value, ps) if per_site else total.value

    def loglik_host_ptr(self, host_ptr, ld, packed2=False, per_site=False):
        total = ctypes.c_double()
        ps = np.empty(self.nsites) if per_site else None
        self._ck(lib.pb_sharded_loglik_host(self._s, ctypes.c_void_p(host_ptr), int(ld), int(bool(packed2)),
                                            ctypes.byref(total), None if ps is None else ps.ctypes.data))
        return (total.value, ps) if per_site else total.value


def transition_matrices_eigen(U, lam, Uinv, ts, device=0):
    """P(t) = U diag(exp(t lam)) Uinv for every t, one kernel launch (K1)."""
    u, ui, l, t = _colmajor(U), _colmajor(Uinv), _f64(lam), _f64(np.atleast_1d(ts))
    n = len(l)
    out = np.empty((len(t), n, n))
    check(lib.pb_transition_matrices_eigen(int(device), n, len(t), u.ctypes.data, l.ctypes.data,
                                           ui.ctypes.data, t.ctypes.data, out.ctypes.data))
    return np.ascontiguousarray(np.swapaxes(out, 1, 2))


def transition_matrices_expm(Q, ts, device=0):
    """P(t) = expm(t Q) for every t by the batched Pade kernel."""
    q, t = _colmajor(Q), _f64(np.atleast_1d(ts))
    n = q.shape[0]
    out = np.empty((len(t), n, n))
    check(lib.pb_transition_matrices_expm(int(device), n, len(t), q.ctypes.data, t.ctypes.data,
                                          out.ctypes.data))
    return np.ascontiguousarray(np.swapaxes(out, 1, 2))


def host_reversible_eigensystem(Q, pi):
    """(U, lambda, Uinv) with Q = U diag(lambda) Uinv — the host step of pb_set_rate_matrix."""
    q, p = _colmajor(Q), _f64(pi)
    n = len(p)
    U, Ui, lam = np.empty((n, n)), np.empty((n, n)), np.empty(n)
    rc = lib.pb_host_reversible_eigensystem(n, q.ctypes.data, p.ctypes.data, U.ctypes.data,
                                            lam.ctypes.data, Ui.ctypes.data)
    if rc != 0:
        raise _lib.PhyloB200Error(rc, "rate matrix is not reversible with respect to pi")
    return np.ascontiguousarray(U.T), lam, np.ascontiguousarray(Ui.T)


def pack_tips2(tips):
    """uint8 states 0..3 [ntaxa][nsites] -> 2-bit packed [ntaxa][ceil(nsites/4)]: site 4b+j in bits 2j..2j+1 of byte b
    (the host layout of pb_set_tips_packed2 / pb_loglik_host_packed2)."""
    tips = np.asarray(tips, dtype=np.uint8)
    if tips.ndim != 2 or (tips.size and tips.max() > 3):
        raise ValueError("pack_tips2: states must be 0..3, shape [ntaxa][nsites]")
    n = tips.shape[1]
    pad = (-n) % 4
    t = np.pad(tips, ((0, 0), (0, pad))) if pad else tips
    t = t.reshape(t.shape[0], -1, 4)
    return (t[:, :, 0] | (t[:, :, 1] << 2) | (t[:, :, 2] << 4) | (t[:, :, 3] << 6)).astype(np.uint8)


def measure_fp64_peak(device=0):
    """pb_measure_fp64_peak: (DFMA TFLOP/s, DMMA TFLOP/s) of the device, measured now."""
    out = []
    for kind in (0, 1):
        v = ctypes.c_double()
        rc = lib.pb_measure_fp64_peak(int(device), kind, ctypes.byref(v))
        if rc != 0:
            raise _lib.PhyloB200Error(rc, "pb_measure_fp64_peak failed")
        out.append(v.value)
    return tuple(out)


def device_local_cpus(device=0):
    """CPUs of the NUMA node the GPU `device` hangs off (sysfs local_cpulist of its PCI function), or None when the
    topology cannot be read.  Host buffers the copy engine reads should live on that node: on a two-socket box a pinned
    buffer first touched from the other socket crosses the socket interconnect on every copy (measured on this pool:
    28 instead of 51 GB/s, run to run, depending on where the scheduler put the process)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return cpus or None
    except Exception:
        return None


def pinned_near_device(array, device=0):
    """A page-locked torch tensor holding `array`, allocated while the calling thread is bound to the GPU's NUMA node
    (first touch places the pages there); the thread's affinity is restored afterwards.  Falls back to a plain pinned
    allocation when the topology or the affinity call is unavailable."""
    import os
    import torch
    src = torch.from_numpy(np.ascontiguousarray(array))
    saved = None
    try:
        local = device_local_cpus(device)
        if local and hasattr(os, "sched_getaffinity"):
            saved = os.sched_getaffinity(0)
            if saved & local:
                os.sched_setaffinity(0, saved & local)
            else:
                saved = None
    except OSError:
        saved = None
    try:
        out = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
        out.copy_(src)
    finally:
        if saved is not None:
            os.sched_setaffinity(0, saved)
    return out


def compress_site_patterns(tips):
    """f4: distinct alignment columns, their multiplicities, and the map back to the original sites
    (tips[:, s] == patterns[:, inverse[s]])."""
    tips = np.asarray(tips)
    patterns, inverse, counts = np.unique(tips, axis=1, return_inverse=True, return_counts=True)
    return np.ascontiguousarray(patterns), counts.astype(np.float64), np.asarray(inverse).reshape(-1)
