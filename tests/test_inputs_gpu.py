"""f4: alignment text encoded to tip states on the device, and site-pattern
compression with weights — the per-site output is re-expanded, the total counts original sites."""
import numpy as np
import pytest

import cases
from oracle import c_oracle as C
from phylogenetics_b200 import _lib, models
from phylogenetics_b200.engine import LikelihoodContext, compress_site_patterns

pytestmark = pytest.mark.gpu


def ctx_for(case, P):
    ctx = LikelihoodContext(case.n, case.nsites, case.flat.nleaves, ncat=case.ncat)
    ctx.set_tree(case.flat, case.bl)
    ctx.set_categories(case.rates, case.weights)
    ctx.set_pmatrices(P)
    ctx.set_root_frequencies(case.pi)
    return ctx


@pytest.mark.parametrize("model,alphabet,symbols", [
    ("gtr", _lib.PB_ALPHABET_DNA, list("ACGT")),
    ("wag", _lib.PB_ALPHABET_AMINO_ACID, list(models.AMINO_ACIDS)),
    ("mg94", _lib.PB_ALPHABET_CODON, models.Codon.ns_triplets),
])
def test_text_alignment_encoded_on_device(model, alphabet, symbols):
    case = cases.make_case("txt" + model, model, 9, 700, seed=4)
    P = case.host_P()
    rows = ["".join(symbols[s] for s in row) for row in case.tips]
    if alphabet == _lib.PB_ALPHABET_DNA:
        rows[2] = rows[2].lower()                        # Nucleotide.of_char_exn accepts either case
    with ctx_for(case, P) as a, ctx_for(case, P) as b:
        a.set_tips(case.tips)
        b.set_tips_text(rows, alphabet)
        ta, pa = a.loglik(per_site=True)
        tb, pb_ = b.loglik(per_site=True)
        assert ta == tb and np.array_equal(pa, pb_)
        bad = list(rows)
        junk = {_lib.PB_ALPHABET_DNA: "N", _lib.PB_ALPHABET_AMINO_ACID: "B", _lib.PB_ALPHABET_CODON: "TAA"}[alphabet]
        bad[4] = junk + bad[4][len(junk):]                   # outside the alphabet / a stop codon
        with pytest.raises(_lib.PhyloB200InvalidArgument):
            b.set_tips_text(bad, alphabet)
    with pytest.raises(_lib.PhyloB200InvalidArgument):
        with ctx_for(case, P) as cwrong:
            cwrong.set_tips_text(rows, (alphabet + 1) % 3)    # alphabet does not match nstates


def test_site_pattern_compression():
    case = cases.make_case("pat", "k80", 6, 5000, seed=2)      # few taxa: many repeated columns
    P = case.host_P()
    f = case.flat
    patterns, counts, inverse = compress_site_patterns(case.tips)
    assert patterns.shape[1] < case.nsites / 3 and counts.sum() == case.nsites
    assert np.array_equal(patterns[:, inverse], case.tips)
    expect = C.pruning(4, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi, tips=case.tips)
    with LikelihoodContext(4, patterns.shape[1], f.nleaves) as ctx:
        ctx.set_tree(f, case.bl)
        ctx.set_pmatrices(P)
        ctx.set_root_frequencies(case.pi)
        ctx.set_tips(patterns)
        ctx.set_site_weights(counts)
        total, per_pattern = ctx.loglik(per_site=True)
        per_site = per_pattern[inverse]                                   # re-expanded to original sites
        assert (np.abs(per_site - expect) / np.abs(expect)).max() <= 1e-9
        assert abs(total - C.sum_exact(expect)) <= 1e-10 * abs(total)
        ctx.set_site_weights(None)
        unweighted = ctx.loglik()
        assert abs(unweighted - per_pattern.sum()) <= 1e-10 * abs(unweighted)
        with pytest.raises(_lib.PhyloB200InvalidArgument):
            ctx.set_site_weights(-counts)


@pytest.mark.parametrize("nsites,chunks", [(20001, 4), (50000, 3), (1003, 8), (7, 1)])
def test_two_bit_packed_host_alignment(nsites, chunks):
    """pb_set_tips_packed2 / pb_loglik_host_packed2: a nucleotide alignment that crosses PCIe at 4 sites per byte
    gives bit-identical results to the byte-per-site calls, on ragged site counts, through the chunked
    pipeline and its small-input fall-back, and for the level kernels (which read the expanded matrix)."""
    import cases
    from phylogenetics_b200.engine import pack_tips2
    case = cases.make_case("pk2", "gtr", 37, nsites, ncat=2, seed=6)
    packed = pack_tips2(case.tips)
    assert packed.shape == (37, (nsites + 3) // 4)
    for level in (False, True):
        with LikelihoodContext(4, nsites, 37, ncat=2, level_kernels=level) as ctx:
            ctx.set_tree(case.flat, case.bl)
            ctx.set_categories(case.rates, case.weights)
            ctx.set_rate_matrix(case.Q, case.pi)
            ctx.set_root_frequencies(case.pi)
            ctx.set_option("host_chunks", chunks)
            ctx.set_tips(case.tips)
            total, per_site = ctx.loglik(per_site=True)
            ctx.set_tips_packed2(packed)
            total2, per_site2 = ctx.loglik(per_site=True)
            assert total2 == total and np.array_equal(per_site2, per_site)
            total3, per_site3 = ctx.loglik_host_packed2(packed, per_site=True)
            assert np.array_equal(per_site3, per_site)
            assert abs(total3 - total) <= 1e-12 * abs(total)
            total4 = ctx.loglik()                        # the expanded matrix stays resident
            assert abs(total4 - total) <= 1e-12 * abs(total)
            ctx.set_tree(case.flat, case.bl)             # ... also the byte-per-site copy expanded on the side stream:
            total5, per_site5 = ctx.loglik(per_site=True)    # a new tree re-packs the tips from it
            assert total5 == total and np.array_equal(per_site5, per_site)
    with LikelihoodContext(20, 10, 5) as ctx:            # 4 states only
        with pytest.raises(_lib.PhyloB200Error):
            ctx.set_tips_packed2(np.zeros((5, 3), np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [{}, {"host_side_streams": 0}, {"host_chunks": 12}, {"host_tail_sites": -1},
                                  {"host_tail_sites": 200000, "host_chunks": 3}, {"host_chunk_align": 0, "host_chunks": 5},
                                  {"host_lead_cats": 1}, {"fused_const_keep": 0}, {"fused_const_cats": 1},
                                  {"host_lanes": 3}, {"host_lanes": 4, "host_chunks": 10}])
def test_pipelined_schedule_variants(opts):
    """The column-chunk schedule of the pipelined host evaluation (whole rounds of the persistent blocks, laid from the
    end with a short last chunk; chunks alternating between two lanes; the 2-bit transpose kernel): every schedule gives
    the per-site values of the resident evaluation to the bit, on enough sites for several chunks, and the timeline
    (pb_debug_host_trace) covers the columns exactly once."""
    import cases
    from phylogenetics_b200.engine import pack_tips2
    nsites = 700_001
    case = cases.make_case("sched", "gtr", 24, nsites, ncat=2, seed=11)
    packed = pack_tips2(case.tips)
    with LikelihoodContext(4, nsites, 24, ncat=2) as ctx:
        ctx.set_tree(case.flat, case.bl)
        ctx.set_categories(case.rates, case.weights)
        ctx.set_rate_matrix(case.Q, case.pi)
        ctx.set_root_frequencies(case.pi)
        ctx.set_tips(case.tips)
        total, per_site = ctx.loglik(per_site=True)
        for name, value in opts.items():
            ctx.set_option(name, value)
        ctx.set_option("host_trace", 1)
        for _ in range(2):                                   # (the second call reuses streams, events and the constant bank)
            total2, per_site2 = ctx.loglik_host_packed2(packed, per_site=True)
            assert np.array_equal(per_site2, per_site) and abs(total2 - total) <= 1e-13 * abs(total)
        chunks, end_ms = ctx.host_trace()
        assert len(chunks) > 1 or opts.get("host_chunks") == 1
        pos = 0
        for first, width, copy_ms, done_ms in chunks:
            assert first == pos and width > 0 and first % 1024 == 0 and 0 < copy_ms < done_ms <= end_ms
            pos += width
        assert pos == (nsites + 1023) // 1024 * 1024
        total3, per_site3 = ctx.loglik_host(case.tips, per_site=True)      # byte per site through the same lanes
        assert np.array_equal(per_site3, per_site) and abs(total3 - total) <= 1e-13 * abs(total)
        total4 = ctx.loglik()
        assert abs(total4 - total) <= 1e-13 * abs(total)


@pytest.mark.gpu
@pytest.mark.parametrize("model,ntaxa,nsites,opts", [("wag", 14, 120_001, {}), ("wag", 14, 120_001, {"host_chunks": 3, "host_tail_sites": -1}),
                                                      ("wag", 9, 30_000, {"host_chunk_align": 0, "host_chunks": 4}),
                                                      ("mg94", 10, 60_001, {}), ("wag", 6, 5000, {})])
def test_pipelined_host_evaluation_dense(model, ntaxa, nsites, opts):
    """pb_loglik_host on the tensor-core path (20 / 61 states): the alignment crosses in column chunks that are pruned
    (fused_dmma3 launched per chunk) and reduced behind the copy — the per-site values of the two-step pb_set_tips +
    pb_loglik to the bit, tips resident afterwards, a state outside the alphabet still an error."""
    import cases
    case = cases.make_case("hostd", model, ntaxa, nsites, seed=13)
    f = case.flat
    with LikelihoodContext(case.n, nsites, ntaxa) as ctx:
        ctx.set_tree(f, case.bl)
        ctx.set_rate_matrix(case.Q, case.pi)
        ctx.set_root_frequencies(case.pi)
        ctx.set_tips(case.tips)
        assert ctx.path_name == "fused_dmma"
        total, per_site = ctx.loglik(per_site=True)
        P = ctx.get_pmatrices()
        chk = 3000
        expect = C.pruning(case.n, f.root, f.child_ptr, f.child_idx, f.tip_row, P, case.pi,
                           tips=np.ascontiguousarray(case.tips[:, :chk]))
        assert np.max(np.abs(per_site[:chk] - expect) / np.abs(expect)) < 1e-9
        for name, value in opts.items():
            ctx.set_option(name, value)
        ctx.set_option("host_trace", 1)
        for _ in range(2):
            total2, per_site2 = ctx.loglik_host(case.tips, per_site=True)
            assert np.array_equal(per_site2, per_site) and abs(total2 - total) <= 1e-13 * abs(total)
        if nsites >= 8192:
            chunks, end_ms = ctx.host_trace()
            assert sum(w for _, w, _, _ in chunks) == (nsites + 1023) // 1024 * 1024
            assert len(chunks) > 1
        total3, per_site3 = ctx.loglik(per_site=True)                 # the tips stay resident
        assert np.array_equal(per_site3, per_site)
        bad = case.tips.copy()
        bad[2, nsites - 1] = case.n
        with pytest.raises(_lib.PhyloB200InvalidArgument):
            ctx.loglik_host(bad)
        total4, per_site4 = ctx.loglik_host(case.tips, per_site=True)
        assert np.array_equal(per_site4, per_site)
