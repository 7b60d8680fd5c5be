"""Mirror of the reference's (deprecated) Felsenstein functor, lib/felsenstein.mli:11-48,
on top of the GPU path.

Felsenstein.Make(A)(Align)(E) becomes `Make(nstates, align_ops, model)`:
  align_ops.get_base(align, seq=<leaf index string>, pos=<site>) -> state int
  align_ops.length(align) -> number of sites
  model.transition_probability_matrix(param) -> (t -> matrix),  model.stationary_distribution(param),
  optionally model.rate_matrix_reduction(param) -> (U, lambda, U^-1) (Make_diag models): then the
  P(t) of every branch is rebuilt on the device from the eigensystem in one launch.
Rescaling: shift_normal with threshold 1e-7, no shift after the product with pi
(lib/felsenstein.ml:47-48,57-67).
"""
import math

import numpy as np

from . import _lib
from . import tree as T
from .engine import LikelihoodContext


class StringAlignment:
    """Alignment.Make(Seq)-like container: dict name -> string over an alphabet
    (lib/alignment.ml:167-204); get_base looks the leaf up by name like the Hashtbl."""

    def __init__(self, alphabet="ACGT"):
        self.alphabet = alphabet

    def of_string_list(self, seqs):
        """Alignment.of_string_list: sequences named T0, T1, ... in order."""
        return {"T%d" % i: s for i, s in enumerate(seqs)}

    def get_base(self, align, seq, pos):
        return self.alphabet.index(align[seq][pos].upper())

    def length(self, align):
        return len(next(iter(align.values()))) if align else 0

    def encode(self, align, names):
        """uint8 [len(names)][length] state matrix (the eager evaluation of get_base)."""
        lut = np.full(256, 255, dtype=np.uint8)
        for i, ch in enumerate(self.alphabet):
            lut[ord(ch)] = i
            lut[ord(ch.lower())] = i
        out = np.stack([lut[np.frombuffer(align[nm].encode(), dtype=np.uint8)] for nm in names])
        if (out == 255).any():
            raise _lib.PhyloB200InvalidArgument(_lib.PB_ERR_INVALID_ARG, "character outside the alphabet")
        return out


class Make:
    def __init__(self, nstates, align_ops, model, device=0):
        self.nstates, self.align, self.model, self.device = nstates, align_ops, model, device

    def _evaluate(self, param, tree, align, threshold, sites=None):
        flat = T.flatten(T.to_tree(tree))
        names = [d[1] for d in flat.leaf_data]
        if hasattr(self.align, "encode"):
            tips = self.align.encode(align, names)
        else:
            L = self.align.length(align)
            tips = np.array([[self.align.get_base(align, seq=nm, pos=p) for p in range(L)] for nm in names],
                            dtype=np.uint8)
        if sites is not None:
            tips = np.ascontiguousarray(tips[:, sites])
        nsites = tips.shape[1]
        if nsites == 0:
            return 0.0, np.zeros(0)
        bl = flat.branch_lengths(float)
        with LikelihoodContext(self.nstates, nsites, flat.nleaves, device=self.device) as ctx:
            ctx.set_tree(flat, bl)
            if hasattr(self.model, "rate_matrix_reduction"):
                ctx.set_eigensystem(*self.model.rate_matrix_reduction(param))
            else:
                tpm = self.model.transition_probability_matrix(param)
                P = np.zeros((1, flat.nnodes, self.nstates, self.nstates))
                for v in range(flat.nnodes):
                    if v != flat.root:
                        P[0, v] = tpm(bl[v])
                ctx.set_pmatrices(P)
            ctx.set_rescaling(threshold, False)
            ctx.set_tips(tips)
            ctx.set_root_frequencies(self.model.stationary_distribution(param))
            return ctx.loglik(per_site=True)

    def felsenstein_single(self, param, site, tree, align):
        """felsenstein_single with the default shift: never rescales (lib/felsenstein.ml:22)."""
        return float(self._evaluate(param, tree, align, -math.inf, sites=[site])[1][0])

    def felsenstein_single_shift(self, param, site, tree, align, threshold=0.0000001):
        return float(self._evaluate(param, tree, align, threshold, sites=[site])[1][0])

    def felsenstein_noshift(self, param, tree, align):
        """lib/felsenstein.ml:77"""
        return self._evaluate(param, tree, align, -math.inf)[0]

    def felsenstein(self, param, tree, align):
        """lib/felsenstein.ml:78 — sum over sites of the shifted single-site log-likelihood."""
        return self._evaluate(param, tree, align, 0.0000001)[0]
