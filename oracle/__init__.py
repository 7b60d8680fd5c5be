"""CPU oracle for the phylogenetic-likelihood hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU, the algorithm of
the reference (biocaml/phylogenetics v0.3.0, OCaml) for Felsenstein pruning and
the per-branch transition matrices.  It exists to check the CUDA path; nothing
in ``phylogenetics_b200/`` may import it.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py``.

Parity status: the reference cannot be built in this environment (no OCaml
toolchain, no GSL/LAPACK) and its own tests pin this path to 1e-5 relative at
best.  The oracle is therefore pinned by
  * the reference tests' RNG-free properties (tests/test_oracle_reference_properties.py),
  * closed-form likelihoods, and
  * a 50-digit mpmath restatement (oracle/mp_certify.py),
not by outputs of the reference binary: **parity unpinned by reference outputs**.

Every function cites the reference file:line it follows (paths relative to the
reference checkout).
"""
