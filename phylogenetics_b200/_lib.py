"""ctypes binding of the C ABI (include/phylo_b200.h).

The shared library is built in-tree (``python -m phylogenetics_b200.build``) and
loaded from this directory.  There is no fallback: if the library is missing the
import fails, and every compute entry point fails when no B200 is present.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PHYLO_B200_LIB names another build of the same library (A/B measurements of a kernel variant, tools/ab_lib.py).
LIB_PATH = os.environ.get("PHYLO_B200_LIB") or os.path.join(_HERE, "libphylo_b200.so")

PB_OK = 0
PB_ERR_INVALID_ARG = -1
PB_ERR_CUDA = -2
PB_ERR_STATE = -3
PB_ERR_NOMEM = -4
PB_ERR_NCCL = -5
PB_ERR_NUMERIC = -6

PB_FLAG_KEEP_CLV = 0x1
PB_FLAG_LEVEL_KERNELS = 0x2
PB_REDUCE_NCCL, PB_REDUCE_ORDERED = 0, 1
PB_MODE_PHYLO_CTMC = 0
PB_MODE_FELSENSTEIN = 1
PB_ALPHABET_DNA, PB_ALPHABET_AMINO_ACID, PB_ALPHABET_CODON = 0, 1, 2

# Every symbol include/phylo_b200.h declares (tests check the .so exports them all).
EXPORTS = [
    "pb_create", "pb_destroy", "pb_last_error", "pb_last_status", "pb_get_dims", "pb_set_stream", "pb_use_private_stream", "pb_set_mode", "pb_set_rescaling",
    "pb_set_tree", "pb_set_branch_lengths", "pb_update_branch_length", "pb_nodes_recomputed", "pb_set_tips", "pb_set_tips_device", "pb_set_tips_packed2", "pb_loglik_host_packed2", "pb_loglik_host_enqueue", "pb_pin_host_memory", "pb_unpin_host_memory", "pb_set_tip_masks", "pb_set_tips_text", "pb_set_site_weights",
    "pb_set_root_frequencies", "pb_set_categories", "pb_set_eigensystem", "pb_set_rate_matrix",
    "pb_set_rate_matrix_general", "pb_set_pmatrices", "pb_loglik", "pb_loglik_host", "pb_loglik_enqueue",
    "pb_loglik_fetch", "pb_result_device_ptrs", "pb_get_clv", "pb_conditional_simulation", "pb_get_pmatrices",
    "pb_set_option", "pb_set_timing", "pb_get_timing", "pb_get_timing_at", "pb_get_kernel_timing", "pb_kernel_launches", "pb_graph_replays", "pb_path_name", "pb_transition_matrices_eigen",
    "pb_transition_matrices_expm", "pb_host_reversible_eigensystem", "pb_debug_fused_program", "pb_debug_reduced_program", "pb_debug_dense_reduced_program", "pb_debug_host_trace", "pb_debug_host_schedule", "pb_measure_fp64_peak", "pb_version",
    "pb_create_sharded", "pb_sharded_destroy", "pb_sharded_last_error", "pb_sharded_ndev", "pb_sharded_ctx",
    "pb_sharded_shard_bounds", "pb_sharded_set_reduce", "pb_sharded_set_mode", "pb_sharded_set_option", "pb_sharded_set_tree",
    "pb_sharded_set_branch_lengths", "pb_sharded_update_branch_length", "pb_sharded_set_root_frequencies",
    "pb_sharded_set_categories", "pb_sharded_set_eigensystem", "pb_sharded_set_rate_matrix",
    "pb_sharded_set_rate_matrix_general", "pb_sharded_set_pmatrices", "pb_sharded_set_tips", "pb_sharded_set_tips_packed2",
    "pb_sharded_set_tip_masks", "pb_sharded_loglik", "pb_sharded_loglik_host",
]


class PhyloB200Error(RuntimeError):
    """Failure reported by the native library (maps to OCaml Failure)."""

    def __init__(self, code, message):
        super().__init__("phylo_b200 error %d: %s" % (code, message))
        self.code = code


class PhyloB200InvalidArgument(PhyloB200Error, ValueError):
    """PB_ERR_INVALID_ARG (maps to OCaml Invalid_argument)."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "phylogenetics_b200: %s is missing. Build it with `python -m phylogenetics_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    sig = {
        "pb_create": (vp, [ci, ci, ci, i64, ci, ctypes.c_uint]),
        "pb_destroy": (None, [vp]),
        "pb_last_error": (ctypes.c_char_p, [vp]),
        "pb_last_status": (ci, [vp]),
        "pb_get_dims": (ci, [vp, vp]),
        "pb_set_stream": (ci, [vp, vp]),
        "pb_use_private_stream": (ci, [vp]),
        "pb_set_mode": (ci, [vp, ci]),
        "pb_set_rescaling": (ci, [vp, dbl, ci]),
        "pb_set_tree": (ci, [vp, ci, ci, vp, vp, vp, vp]),
        "pb_set_branch_lengths": (ci, [vp, vp]),
        "pb_update_branch_length": (ci, [vp, ci, dbl]),
        "pb_nodes_recomputed": (ci, [vp]),
        "pb_set_tips": (ci, [vp, vp, i64]),
        "pb_set_tips_device": (ci, [vp, vp, i64]),
        "pb_set_tips_packed2": (ci, [vp, vp, i64]),
        "pb_set_tip_masks": (ci, [vp, vp, i64]),
        "pb_set_tips_text": (ci, [vp, ctypes.c_char_p, i64, ci]),
        "pb_set_site_weights": (ci, [vp, vp]),
        "pb_set_root_frequencies": (ci, [vp, vp]),
        "pb_set_categories": (ci, [vp, vp, vp]),
        "pb_set_eigensystem": (ci, [vp, vp, vp, vp]),
        "pb_set_rate_matrix": (ci, [vp, vp, vp]),
        "pb_set_rate_matrix_general": (ci, [vp, vp]),
        "pb_set_pmatrices": (ci, [vp, vp]),
        "pb_loglik": (ci, [vp, vp, vp]),
        "pb_loglik_host": (ci, [vp, vp, i64, vp, vp]),
        "pb_loglik_host_packed2": (ci, [vp, vp, i64, vp, vp]),
        "pb_loglik_host_enqueue": (ci, [vp, vp, i64, ci]),
        "pb_pin_host_memory": (ci, [vp, ctypes.c_size_t]),
        "pb_unpin_host_memory": (ci, [vp]),
        "pb_loglik_enqueue": (ci, [vp]),
        "pb_loglik_fetch": (ci, [vp, vp, vp]),
        "pb_result_device_ptrs": (ci, [vp, vp, vp]),
        "pb_get_clv": (ci, [vp, ci, ci, vp, vp]),
        "pb_conditional_simulation": (ci, [vp, ctypes.c_uint64, vp, ctypes.c_int64]),
        "pb_get_pmatrices": (ci, [vp, vp]),
        "pb_set_option": (ci, [vp, ctypes.c_char_p, dbl]),
        "pb_set_timing": (ci, [vp, ci]),
        "pb_get_timing": (ci, [vp, vp]),
        "pb_get_timing_at": (ci, [vp, ci, vp]),
        "pb_get_kernel_timing": (ci, [vp, vp, vp]),
        "pb_kernel_launches": (i64, [vp]),
        "pb_graph_replays": (i64, [vp]),
        "pb_path_name": (ctypes.c_char_p, [vp]),
        "pb_transition_matrices_eigen": (ci, [ci, ci, ci, vp, vp, vp, vp, vp]),
        "pb_transition_matrices_expm": (ci, [ci, ci, ci, vp, vp, vp]),
        "pb_host_reversible_eigensystem": (ci, [ci, vp, vp, vp, vp, vp]),
        "pb_debug_fused_program": (ci, [ci, ci, vp, vp, vp, ci, ci, vp, vp, vp, ci, vp, vp, vp, vp, ci]),
        "pb_create_sharded": (vp, [ci, vp, ci, ci, i64, ci, ctypes.c_uint]),
        "pb_sharded_destroy": (None, [vp]),
        "pb_sharded_last_error": (ctypes.c_char_p, [vp]),
        "pb_sharded_ndev": (ci, [vp]),
        "pb_sharded_ctx": (vp, [vp, ci]),
        "pb_sharded_shard_bounds": (ci, [vp, ci, vp, vp]),
        "pb_sharded_set_reduce": (ci, [vp, ci]),
        "pb_sharded_set_mode": (ci, [vp, ci]),
        "pb_sharded_set_option": (ci, [vp, ctypes.c_char_p, dbl]),
        "pb_sharded_set_tree": (ci, [vp, ci, ci, vp, vp, vp, vp]),
        "pb_sharded_set_branch_lengths": (ci, [vp, vp]),
        "pb_sharded_update_branch_length": (ci, [vp, ci, dbl]),
        "pb_sharded_set_root_frequencies": (ci, [vp, vp]),
        "pb_sharded_set_categories": (ci, [vp, vp, vp]),
        "pb_sharded_set_eigensystem": (ci, [vp, vp, vp, vp]),
        "pb_sharded_set_rate_matrix": (ci, [vp, vp, vp]),
        "pb_sharded_set_rate_matrix_general": (ci, [vp, vp]),
        "pb_sharded_set_pmatrices": (ci, [vp, vp]),
        "pb_sharded_set_tips": (ci, [vp, vp, i64]),
        "pb_sharded_set_tips_packed2": (ci, [vp, vp, i64]),
        "pb_sharded_set_tip_masks": (ci, [vp, vp, i64]),
        "pb_sharded_loglik": (ci, [vp, vp, vp]),
        "pb_sharded_loglik_host": (ci, [vp, vp, i64, ci, vp, vp]),
        "pb_debug_reduced_program": (ci, [ci, ci, vp, vp, vp, ci, ci, vp, vp, ci, vp, ci, vp, vp, ci, vp, vp, vp, ci]),
        "pb_debug_dense_reduced_program": (ci, [ci, ci, vp, vp, vp, ci, ci, vp, vp, ci, vp, ci, vp, vp, ci, vp, vp]),
        "pb_debug_host_trace": (ci, [vp, vp, ci]),
        "pb_measure_fp64_peak": (ci, [ci, ci, vp]),
        "pb_debug_host_schedule": (ci, [i64, ci, i64, i64, vp, ci, vp]),
        "pb_version": (ctypes.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    return L


lib = _load()


def check(rc, ctx=None):
    """Raise on a negative status, with the library's message."""
    if rc == PB_OK:
        return
    msg = lib.pb_last_error(ctx)
    msg = msg.decode() if msg else ""
    if rc == PB_ERR_INVALID_ARG:
        raise PhyloB200InvalidArgument(rc, msg)
    raise PhyloB200Error(rc, msg)
