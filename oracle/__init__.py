"""CPU oracle for the phylogenetic-likelihood hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU, the algorithm of
the reference (biocaml/phylogenetics v0.3.0, OCaml) for Felsenstein pruning and
the per-branch transition matrices.  It exists to check the CUDA path; nothing
in ``phylogenetics_b200/`` may import it.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py``.

Parity status: PINNED against the reference's own golden vector.  The reference cannot be
built in this environment (no OCaml toolchain, no GSL/LAPACK), but its dune expect test
tests/expect/phylo_ctmc_substitution_mappings.expected:5 prints the pruning log-likelihood
-117.326089 of a WAG-simulated site on a 100-taxon tree, both drawn from GSL's default random
stream.  oracle/gsl_replay.py replays that stream from the published GSL algorithms (checked on
the way against the two histograms the same test prints and against the Newick string of
tests/expect/birth_death_simulator.expected), and the oracle prints -117.326089 for the replayed
tree and site (tests/test_oracle_reference_golden.py; fixture
tests/golden/reference_substitution_mappings.json).  Beyond the 6 printed decimals the oracle is
certified by
  * the reference tests' RNG-free properties (tests/test_oracle_reference_properties.py),
  * closed-form likelihoods, and
  * a 60-digit mpmath restatement (oracle/make_golden.py -> tests/golden/known_answers.json).
The +Gamma mixture is an extension the reference does not have: unpinned by construction.

Every function cites the reference file:line it follows (paths relative to the
reference checkout).
"""
