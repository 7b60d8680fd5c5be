"""phylogenetics_b200 — B200-native (sm_100a) phylogenetic likelihood: Felsenstein
pruning with the reference's rescaling and per-branch P(t) = exp(Qt), behind the
C ABI of include/phylo_b200.h.  Host layer mirroring biocaml/phylogenetics' API
for that path; everything numeric runs in libphylo_b200.so.

Submodules are imported on first use so that `python -m phylogenetics_b200.build`
works before the library exists.
"""
import importlib

_SUBMODULES = ("_lib", "engine", "tree", "models", "phylo_ctmc", "felsenstein", "synthetic",
               "sharding", "nelder_mead")

__all__ = list(_SUBMODULES)


def __getattr__(name):
    if name in _SUBMODULES:
        return importlib.import_module("." + name, __name__)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
