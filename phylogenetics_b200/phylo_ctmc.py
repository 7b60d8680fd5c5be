"""Mirror of the reference's Phylo_ctmc likelihood interface (lib/phylo_ctmc.mli:55-73)
on top of the GPU path.

The reference API prunes ONE site per call with two closures.  Closures cannot
cross the C ABI, so the wrapper evaluates them eagerly: the tree is flattened
once, `transition_probabilities` is called once per branch and `leaf_state` once
per (leaf, site) into a uint8 matrix.  `pruning` keeps the per-site
signature (S = 1: correct, slow); `pruning_sites` is the batched entry beside it.
"""
import numpy as np

from . import _lib
from . import tree as T
from .engine import LikelihoodContext


def identity(dim):
    return np.eye(dim)


def matrix_decomposition_reduce(dim, decomp):
    """lib/phylo_ctmc.ml:13-24 — multiply a [`Mat|`Transpose|`Diag] list out (host, once per branch)."""
    if not decomp:
        return identity(dim)
    kind, x = decomp[0]
    x = np.asarray(x, dtype=np.float64)
    head = x if kind == "Mat" else (x.T if kind == "Transpose" else np.diag(x))
    if len(decomp) == 1:
        return head
    return head @ matrix_decomposition_reduce(dim, decomp[1:])


def _dense(tp, nstates):
    if isinstance(tp, (list, tuple)) and (len(tp) == 0 or isinstance(tp[0], tuple)):
        return matrix_decomposition_reduce(nstates, list(tp))
    return np.asarray(tp, dtype=np.float64)


def _prepare(t, nstates, transition_probabilities, device, keep_clv, nsites, level_kernels=False):
    flat = T.flatten(t)
    P = np.zeros((1, flat.nnodes, nstates, nstates))
    for v in range(flat.nnodes):
        if v != flat.root:
            m = _dense(transition_probabilities(flat.branch_data[v]), nstates)
            if m.shape != (nstates, nstates):
                raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "transition matrix has the wrong shape")
            P[0, v] = m
    ctx = LikelihoodContext(nstates, nsites, flat.nleaves, ncat=1, device=device, keep_clv=keep_clv,
                            level_kernels=level_kernels)
    ctx.set_tree(flat)
    ctx.set_pmatrices(P)
    return flat, P, ctx


def _leaf_matrix(flat, leaf_state, nsites, nstates):
    """leaf_state(leaf_payload, site) -> int, or an array [nleaves][nsites] already."""
    if not callable(leaf_state):
        tips = np.ascontiguousarray(leaf_state, dtype=np.uint8)
        if tips.shape != (flat.nleaves, nsites):
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "tip matrix must be [nleaves][nsites]")
        return tips
    tips = np.empty((flat.nleaves, nsites), dtype=np.uint8)
    for row, data in enumerate(flat.leaf_data):
        for s in range(nsites):
            st = leaf_state(data, s)
            if not 0 <= st < nstates:
                raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "leaf state out of range")
            tips[row, s] = st
    return tips


def pruning_sites(t, nstates, transition_probabilities, leaf_state, nsites, root_frequencies,
                  device=0, level_kernels=False):
    """Batched Phylo_ctmc.pruning: per-site log-likelihoods [nsites] of one tree.

    leaf_state(leaf_payload, site) -> state, or a uint8 matrix [nleaves][nsites]
    (rows in left-to-right leaf order).
    """
    flat, _, ctx = _prepare(t, nstates, transition_probabilities, device, False, nsites, level_kernels)
    with ctx:
        ctx.set_mode(_lib.PB_MODE_PHYLO_CTMC)
        tips = _leaf_matrix(flat, leaf_state, nsites, nstates)
        if nstates == 4:
            from .engine import pack_tips2
            ctx.set_tips_packed2(pack_tips2(tips))        # nucleotides cross PCIe at 4 sites per byte
        else:
            ctx.set_tips(tips)
        ctx.set_root_frequencies(root_frequencies)
        _, per_site = ctx.loglik(per_site=True)
    return per_site


def pruning(t, nstates, transition_probabilities, leaf_state, root_frequencies, device=0):
    """Phylo_ctmc.pruning (lib/phylo_ctmc.mli:55-61, .ml:83-98): log-likelihood of one site.

    transition_probabilities: branch payload -> matrix_decomposition (list of ("Mat", m) /
    ("Transpose", m) / ("Diag", v)) or a dense matrix; leaf_state: leaf payload -> int.
    """
    return float(pruning_sites(t, nstates, transition_probabilities, lambda l, _s: leaf_state(l), 1,
                               root_frequencies, device=device)[0])


def conditional_likelihoods_sites(t, nstates, leaf_state, transition_probabilities, nsites, device=0):
    """Every node's shifted vector for every site: (flat, clv [nnodes][nsites][n], carry [nnodes][nsites], P)."""
    flat, P, ctx = _prepare(t, nstates, transition_probabilities, device, True, nsites)
    with ctx:
        ctx.set_mode(_lib.PB_MODE_PHYLO_CTMC)
        ctx.set_tips(_leaf_matrix(flat, leaf_state, nsites, nstates))
        ctx.set_root_frequencies(np.full(nstates, 1.0 / nstates))   # the root product is not part of the CLVs
        ctx.loglik()
        clv = np.empty((flat.nnodes, nsites, nstates))
        carry = np.empty((flat.nnodes, nsites))
        for v in range(flat.nnodes):
            clv[v], carry[v] = ctx.get_clv(v, 0)
    return flat, clv, carry, P[0]


def conditional_likelihoods(t, nstates, leaf_state, transition_probabilities, device=0):
    """Phylo_ctmc.conditional_likelihoods (lib/phylo_ctmc.mli:68-73, .ml:100-120) for one site:
    a tree of the same shape with node payload SV = (vector, carry), leaf payload = state and
    branch payload = (branch data, P)."""
    flat, clv, carry, P = conditional_likelihoods_sites(
        t, nstates, lambda l, _s: leaf_state(l), transition_probabilities, 1, device=device)
    states = [int(np.argmax(clv[v, 0])) if flat.is_leaf(v) else None for v in range(flat.nnodes)]
    row_state = {int(flat.tip_row[v]): states[v] for v in range(flat.nnodes) if flat.is_leaf(v)}
    return T.unflatten(flat,
                       node=lambda v, _d: (clv[v, 0].copy(), float(carry[v, 0])),
                       leaf=lambda row, _d: row_state[row],
                       branch=lambda v, d: (d, P[v]))


def conditional_simulation_sites(seed, t, nstates, leaf_state, transition_probabilities, nsites, root_frequencies, device=0):
    """Phylo_ctmc.conditional_simulation (lib/phylo_ctmc.mli:80-90, .ml:122-143) for every site at once, on
    the device: conditional likelihoods (kept in HBM) followed by one top-down joint draw of all node states.
    Returns (flat, states uint8 [nnodes][nsites]); observed leaves keep their states.  `seed` keys a
    counter-based generator (the reference threads a Gsl.Rng.t: same distribution, another stream)."""
    flat, P, ctx = _prepare(t, nstates, transition_probabilities, device, True, nsites)
    with ctx:
        ctx.set_mode(_lib.PB_MODE_PHYLO_CTMC)
        ctx.set_tips(_leaf_matrix(flat, leaf_state, nsites, nstates))
        ctx.set_root_frequencies(np.asarray(root_frequencies, dtype=np.float64))
        ctx.loglik()
        return flat, ctx.conditional_simulation(seed)


def _mask_matrix(flat, nsites, mask_of):
    """uint64 [nleaves][nsites]: bit k set = state k compatible with the observation; 0 = missing."""
    masks = np.zeros((flat.nleaves, nsites), dtype=np.uint64)
    for row, data in enumerate(flat.leaf_data):
        for s in range(nsites):
            masks[row, s] = mask_of(data, s)
    return masks


def _pruning_sites_masks(t, nstates, transition_probabilities, mask_of, nsites, root_frequencies, device):
    flat, _, ctx = _prepare(t, nstates, transition_probabilities, device, False, nsites)
    with ctx:
        ctx.set_mode(_lib.PB_MODE_PHYLO_CTMC)
        ctx.set_tip_masks(_mask_matrix(flat, nsites, mask_of))
        ctx.set_root_frequencies(root_frequencies)
        _, per_site = ctx.loglik(per_site=True)
    return per_site


class Missing_values:
    """Phylo_ctmc.Missing_values (lib/phylo_ctmc.mli:91-128, .ml:145-172): leaves may be unobserved.
    `leaf_state` returns a state or None; a missing child contributes nothing (no product, no shift) and a site
    without any observation has log-likelihood 0."""

    @staticmethod
    def pruning_sites(t, nstates, transition_probabilities, leaf_state, nsites, root_frequencies, device=0):
        """All sites at once: leaf_state(leaf payload, site) -> int | None."""
        def mask_of(data, s):
            st = leaf_state(data, s)
            if st is None:
                return 0
            if not 0 <= st < nstates:
                raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "leaf state out of range")
            return 1 << st
        return _pruning_sites_masks(t, nstates, transition_probabilities, mask_of, nsites, root_frequencies, device)

    @staticmethod
    def pruning(t, nstates, transition_probabilities, leaf_state, root_frequencies, device=0):
        """One site, the reference's signature: leaf_state(leaf payload) -> int | None."""
        return float(Missing_values.pruning_sites(t, nstates, transition_probabilities, lambda l, _s: leaf_state(l), 1,
                                                  root_frequencies, device=device)[0])


class Ambiguous:
    """Phylo_ctmc.Ambiguous (lib/phylo_ctmc.mli:130-160, .ml:232-302): `leaf_state l i` tells whether state i is
    compatible with the observation at leaf l; the tip vector is the 0/1 indicator of that set, an empty set is a
    missing observation."""

    @staticmethod
    def pruning_sites(t, nstates, transition_probabilities, leaf_state, nsites, root_frequencies, device=0):
        """All sites at once: leaf_state(leaf payload, site, state) -> bool."""
        def mask_of(data, s):
            m = 0
            for i in range(nstates):
                if leaf_state(data, s, i):
                    m |= 1 << i
            return m
        return _pruning_sites_masks(t, nstates, transition_probabilities, mask_of, nsites, root_frequencies, device)

    @staticmethod
    def pruning(t, nstates, transition_probabilities, leaf_state, root_frequencies, device=0):
        """One site, the reference's signature: leaf_state(leaf payload, state) -> bool."""
        return float(Ambiguous.pruning_sites(t, nstates, transition_probabilities, lambda l, _s, i: leaf_state(l, i), 1,
                                             root_frequencies, device=device)[0])
