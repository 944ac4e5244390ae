"""Build-container logic tests: the product's pipeline source compiled against the host-simulation
backend (tests/hostsim) must reproduce the reference fixtures and the oracle on seeded inputs.
The same assertions run on the real CUDA library in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import compare
from tests import util, golden_util as G


@pytest.mark.parametrize("name", G.case_names())
def test_files_match_reference(hostsim, name):
    c = G.load_case(name)
    kw = G.args_to_kw(c["meta"]["args"])
    got, res, _ = util.product_outputs(hostsim, c["vcf"], c["sams"], mapq=c["meta"]["mapq"], paired_end=c["meta"]["paired_end"], **kw)
    bad = compare.diff_outputs(c["ref"], got)
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("name", G.case_names())
def test_mapper_tuples_match_oracle(hostsim, name):
    c = G.load_case(name)
    kw = G.args_to_kw(c["meta"]["args"])
    vt, st, batches, col, fd = util.load_inputs(c["vcf"], c["sams"], mapq=c["meta"]["mapq"], paired_end=c["meta"]["paired_end"],
                                                remove_dups=kw.get("remove_dups", 1), pass_only=kw.get("pass_only", 1),
                                                id_separator=kw.get("id_separator", "_"), gw_phase_method=kw.get("gw_phase_method", 0),
                                                include_indels=kw.get("include_indels", 0))
    for batch in batches:
        got, exp = util.compare_tuples(hostsim, vt, batch)
        assert got == exp


@pytest.mark.parametrize("seed,n_bams,switch,mbs", [(31, 1, 0.002, 15), (32, 2, 0.01, 15), (33, 1, 0.03, 4),
                                                     (34, 2, 0.05, 3), (35, 1, 0.03, 5)])
def test_seeded_cases_match_oracle(hostsim, tmp_path, seed, n_bams, switch, mbs):
    vcf, sams = util.make_case(tmp_path, seed, 250, 2500, n_bams=n_bams, switch_per_base=switch)
    got, res, _ = util.product_outputs(hostsim, vcf, sams, max_block_size=mbs)
    exp, ores = util.oracle_outputs(vcf, sams, max_block_size=mbs)
    bad = compare.diff_outputs(exp, got)
    assert not bad, "\n".join(bad)
    assert res.counters["n_tuples"] == ores.total_tuples
    assert abs(res.noise_e - ores.noise_e) == 0.0


def test_totals_above_the_precomputed_range_take_the_side_list(hostsim, tmp_path, monkeypatch):
    """Critical values of c_total values above pipeline.PRECOMPUTED_TOTALS come from the device's side list
    (big_tot); with the range shrunk to 3 nearly every tested edge goes that way and nothing may change."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 36, 250, 2500, n_bams=2, switch_per_base=0.03)
    exp, _ = util.oracle_outputs(vcf, sams, max_block_size=5)
    monkeypatch.setattr(pipeline, "PRECOMPUTED_TOTALS", 3)
    got, res, _ = util.product_outputs(hostsim, vcf, sams, max_block_size=5)
    bad = compare.diff_outputs(exp, got)
    assert not bad, "\n".join(bad)
    assert res.counters["edges"] > 0


def test_windowed_k1_logic_equals_generic(hostsim):
    """Slab selection + in-slab range search (shared with the CUDA kernel) against plain global search."""
    import numpy as np
    from phaser_b200 import synth
    g = synth.make_genome(61, 6000, exonic_frac=0.3, n_genes=400, contigs=synth.GRCH38[19:22])
    vt = synth.to_variant_table_arrays(g)
    rb = synth.to_read_batch(synth.make_reads(g, 6100, 40000), len(vt.contigs), "b0")
    d = hostsim.upload_reads(rb)
    out = {}
    try:
        for mode in (0, 1, 2, 3, 4):          # 4: the tile path with every candidate re-derived by the emission
            hostsim.set_option("k1_mode", min(mode, 3)); hostsim.set_option("k1_staged_emit", 0 if mode == 4 else 1)
            hostsim.set_variants(vt)
            n = hostsim.map_reads(d, 10, 0.0)
            out[mode] = (n, hostsim.download("t_rec"), hostsim.download("t_var"), hostsim.download("t_misc"))
    finally:
        hostsim.set_option("k1_mode", 3); hostsim.set_option("k1_staged_emit", 1)
    assert out[0][0] == out[1][0] == out[2][0] == out[3][0] == out[4][0] and out[0][0] > 5000
    for m in (1, 2, 3, 4):
        for a, b in zip(out[0][1:], out[m][1:]):
            assert np.array_equal(a, b), m


def test_wgs_plus_rna_joint_matches_oracle(hostsim, tmp_path):
    from phaser_b200 import synth
    contigs = [("21", 200000), ("22", 150000)]
    g = synth.make_genome(71, 500, contigs=contigs, n_genes=40)
    vcf = synth.write_vcf(g, str(tmp_path / "j.vcf.gz"))
    wgs = synth.make_wgs_reads(g, 7100, 6000, mapq=60, lowmapq_frac=0.05, dup_frac=0.03)
    rna = synth.make_reads(g, 7101, 3000, dup_frac=0.05)
    sams = [synth.write_sam(wgs, g, str(tmp_path / "wgs.bam"), "w"), synth.write_sam(rna, g, str(tmp_path / "rna.bam"), "r")]
    kw = dict(mapq="20,255", paired_end="1,1", exclude=[0])
    got, res, _ = util.product_outputs(hostsim, vcf, sams, **kw)
    exp, ores = util.oracle_outputs(vcf, sams, **kw)
    bad = compare.diff_outputs(exp, got)
    assert not bad, "\n".join(bad)
    assert res.counters["n_tuples"] == ores.total_tuples > 3000


@pytest.mark.parametrize("n_quals,expect_bits", [(2, 1), (3, 2), (4, 2), (11, 4), (16, 4), (40, 8)])
def test_packed_transport_is_lossless(hostsim, tmp_path, n_quals, expect_bits):
    """Every index width of the quality table, bases outside A/C/G/T (N, IUPAC, '=') through the exception list."""
    vcf, sams = util.make_case(tmp_path, 41, 200, 1500, n_bams=1)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    b = batches[0]
    rng = np.random.default_rng(n_quals)
    b.qual = rng.choice(np.arange(2, 2 + 2 * n_quals, 2), size=b.qual.shape[0]).astype(np.uint8)     # around baseq = 10
    seq = b.seq.copy()
    hit = rng.random(seq.shape[0]) < 0.02
    seq[hit] = rng.integers(0, 256, size=int(hit.sum()), dtype=np.uint8)                           # any pair of 4-bit codes
    b.seq = seq
    p = util.packed_vs_plain(hostsim, vt, b, len(vt.contigs))
    assert p.qual_bits == expect_bits and p.n_exceptions > 0
    plain_bytes = sum(a.nbytes for a in (b.pos, b.tlen, b.aln_score, b.frag, b.cigar_off, b.cigar, b.seq_off, b.seq, b.qual))
    assert p.nbytes < plain_bytes * (0.5 if expect_bits == 1 else 1.0)      # 4 % exceptions here cost 9 bytes each


def test_packed_transport_refuses_oversized_records(hostsim):
    from phaser_b200 import engine as eng
    from phaser_b200.layout import ReadBatch
    n = 70000
    b = ReadBatch(1, np.array([0, 1], np.int64), np.array([5], np.int32), np.zeros(1, np.int32), np.zeros(1, np.int16),
                  np.zeros(1, np.uint32), np.array([0, 1], np.uint32), np.array([n << 4], np.uint32), np.array([0, n], np.uint64),
                  np.full((n + 1) // 2, 0x11, np.uint8), np.full(n, 30, np.uint8), None)
    with pytest.raises(eng.PhzError, match="not packable"):
        eng.pack_reads(b, 1, lib=hostsim.lib)


def test_result_buffer_views_equal_private_copies(hostsim, tmp_path):
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 43, 200, 1500, n_bams=1)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams()
    a = pipeline.run_path(hostsim, vt, [hostsim.upload_reads(b) for b in batches], P, n_fragments=len(fd.names))
    b = pipeline.run_path(hostsim, vt, [hostsim.upload_reads(b) for b in batches], P, n_fragments=len(fd.names),
                          reuse_result_buffer=True)
    assert set(a.arrays) == set(b.arrays)
    for k in a.arrays:
        assert a.arrays[k].dtype == b.arrays[k].dtype and np.array_equal(a.arrays[k], b.arrays[k]), k


def test_fragment_sort_fixup_and_fallback_agree(hostsim, tmp_path):
    """The graph stage sorts tuples on the fragment bits only and repairs interleaved runs in place; with
    frag_run_limit = 1 every multi-tuple fragment takes the full-key sort instead.  Same arrays either way, and the
    case (short inserts: overlapping mates) must exercise the in-place repair."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 44, 400, 8000, n_bams=2, switch_per_base=0.02, insert_lo=60, insert_hi=200)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams()
    hostsim.set_option("graph_mode", 0)           # the sort-based stage (A/B switch and fallback of the fragment-table stage)
    try:
        a = pipeline.run_path(hostsim, vt, [hostsim.upload_reads(b) for b in batches], P, n_fragments=len(fd.names))
        assert a.counters["frag_runs_resorted"] > 0 and a.counters["full_sort_fallback"] == 0
        hostsim.set_option("frag_run_limit", 1)
        b = pipeline.run_path(hostsim, vt, [hostsim.upload_reads(b) for b in batches], P, n_fragments=len(fd.names))
    finally:
        hostsim.set_option("frag_run_limit", 1024); hostsim.set_option("graph_mode", 1)
    assert b.counters["full_sort_fallback"] == 1
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k
    exp, _ = util.oracle_outputs(vcf, sams)
    got, _, _ = util.product_outputs(hostsim, vcf, sams)
    assert not compare.diff_outputs(exp, got)


def test_prefetched_packed_sample_is_the_one_mapped(hostsim, tmp_path):
    """phz_prefetch_packed stages a sample in the free transport slot; phz_map_reads_packed with the same buffers
    consumes the oldest staged copy; two slots, a third outstanding prefetch is refused."""
    from phaser_b200 import engine as eng
    vcf, sams = util.make_case(tmp_path, 45, 200, 1500, n_bams=2)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    hostsim.set_variants(vt)
    packs = [eng.pack_reads(b, len(vt.contigs), lib=hostsim.lib) for b in batches]
    want = []
    for b in batches:
        hostsim.map_reads(hostsim.upload_reads(b), 10, 0.0)
        want.append([hostsim.download(k).copy() for k in ("t_rec", "t_var", "t_misc")])
    hostsim.prefetch_packed(packs[0]); hostsim.prefetch_packed(packs[1])
    with pytest.raises(eng.PhzError, match="both transport slots"):
        hostsim.prefetch_packed(packs[0])
    for order in ((1, 0), (0, 1)):
        for i in order:
            hostsim.map_reads_packed(packs[i], 10, 0.0)
            got = [hostsim.download(k).copy() for k in ("t_rec", "t_var", "t_misc")]
            assert all(np.array_equal(a, b) for a, b in zip(want[i], got))
        hostsim.prefetch_packed(packs[0]); hostsim.prefetch_packed(packs[1])
    hostsim.map_reads_packed(packs[0], 10, 0.0); hostsim.map_reads_packed(packs[1], 10, 0.0)       # drain
    # steady-state loop: prefetch the next while mapping the current
    hostsim.prefetch_packed(packs[0])
    for i in (0, 1, 0, 1):
        hostsim.prefetch_packed(packs[1 - i])
        hostsim.map_reads_packed(packs[i], 10, 0.0)
        got = [hostsim.download(k).copy() for k in ("t_rec", "t_var", "t_misc")]
        assert all(np.array_equal(a, b) for a, b in zip(want[i], got))
    hostsim.map_reads_packed(packs[0], 10, 0.0)


def _odd_batch(n_contigs, seed=9):
    """Hand-made records that push every field of the transport form off its narrow coding: > 65536 distinct CIGAR
    words, records with > 255 operations, unequal read lengths, > 256 distinct alignment scores, TLENs beyond 16 bits,
    position gaps beyond 16 bits and a contig restart."""
    from phaser_b200.layout import ReadBatch
    rng = np.random.default_rng(seed)
    nrec = 320
    cig = []; coff = [0]; soff = [0]; counter = 0
    for k in range(nrec):
        first = k % 7 + 1
        words = [(first << 4) | 0]
        for _ in range(300 if k % 40 == 0 else 240):
            counter += 1
            words.append(((counter // 2 + 1) << 4) | (2 if counter % 2 else 3))      # D / N, every (op, length) pair once
        words.append((1 << 4) | 0)
        cig += words; coff.append(len(cig)); soff.append(soff[-1] + first + 1)
    nb = soff[-1]
    pos = np.cumsum(rng.choice([0, 3, 200, 70000, 900000], size=nrec)).astype(np.int64) + 1
    half = nrec // 2
    pos[half:] -= pos[half] - 5                                                       # second contig starts over
    off = np.zeros(n_contigs + 1, np.int64); off[1] = half; off[2:] = nrec
    tlen = rng.choice([0, 250, -300, 40000, -1200000, 32767, -32768, 32768], size=nrec).astype(np.int32)
    return ReadBatch(n_contigs, off, pos.astype(np.int32), tlen, rng.integers(-400, 400, size=nrec).astype(np.int16),
                     np.arange(nrec, dtype=np.uint32), np.asarray(coff, np.uint32), np.asarray(cig, np.uint32),
                     np.asarray(soff, np.uint64), rng.integers(0, 256, size=(nb + 1) // 2, dtype=np.uint8),
                     rng.integers(0, 60, size=nb).astype(np.uint8), None)


def test_packed_transport_wide_codings_round_trip(hostsim, tmp_path):
    vcf, sams = util.make_case(tmp_path, 41, 200, 300, n_bams=1)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    b = _odd_batch(len(vt.contigs))
    p = util.packed_vs_plain(hostsim, vt, b, len(vt.contigs))
    c = p.coding
    assert c["cigar_bits"] == 32 and c["n_cigar_bits"] == 16 and c["l_seq_const"] == -1 and c["as_bits"] == 16
    assert c["pos_exceptions"] > 2 and c["tlen_exceptions"] > 2 and c["qual_bits"] == 8
    # and the narrow side of the same fields on ordinary data
    c = util.packed_vs_plain(hostsim, vt, batches[0], len(vt.contigs)).coding
    assert c["cigar_bits"] == 16 and c["n_cigar_bits"] == 8 and c["l_seq_const"] == 76 and c["as_bits"] == 8


@pytest.mark.parametrize("ids", ["first_appearance", "offset", "shuffled", "far_back", "gaps"])
def test_packed_fragment_ids_round_trip(hostsim, tmp_path, ids):
    """Fragment ids numbered by first appearance travel as a bitmap + 16-bit back references (phz.h: frag_first /
    frag_back); anything else must come back bit for bit too -- through the exception list or as plain 32-bit ids."""
    vcf, sams = util.make_case(tmp_path, 43, 200, 2500, n_bams=1)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    b = batches[0]
    first = np.unique(b.frag, return_index=True)[1]
    rank = np.empty(int(b.frag.max()) + 1, np.int64); rank[b.frag[np.sort(first)]] = np.arange(first.shape[0])
    dense = rank[b.frag]                                         # ids in order of first appearance
    rng = np.random.default_rng(7)
    if ids == "first_appearance":
        f = dense; expect = 16
    elif ids == "offset":                                        # a second BAM: new names continue after the first BAM's
        f = dense + 1_000_000; expect = 16
    elif ids == "shuffled":                                      # arbitrary numbering: plain ids
        f = rng.permutation(first.shape[0])[dense] * 100000; expect = 32
    elif ids == "far_back":                                      # a few mates further back than 16 bits can say: exceptions
        f = dense.copy(); f[-5:] = 0; f[-1] = f.max() + 70000; expect = 16
    else:                                                        # ids with holes (a shard's view of a global numbering)
        f = dense * 3; expect = None        # small here: either form may win, the round trip is what counts
    b.frag = f.astype(np.uint32)
    p = util.packed_vs_plain(hostsim, vt, b, len(vt.contigs))
    assert expect is None or p.coding["frag_bits"] == expect, p.coding
    if ids == "far_back":
        assert p.coding["frag_exceptions"] >= 1


def test_compact_pair_keys_equal_wide_ones(hostsim, tmp_path):
    """Pair table sorted on (va << dbits | vb - va) (32-bit when it fits) == sorted on 64-bit keys."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 49, 500, 9000, n_bams=2, switch_per_base=0.02)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams()
    out = []
    for wide in (0, 1):
        hostsim.set_option("wide_pair_keys", wide); hostsim.set_option("graph_mode", 0)
        try:
            out.append(pipeline.run_path(hostsim, vt, [hostsim.upload_reads(b) for b in batches], P, n_fragments=len(fd.names)))
        finally:
            hostsim.set_option("wide_pair_keys", 0); hostsim.set_option("graph_mode", 1)
    a, b = out
    assert a.counters == b.counters and a.counters["edges"] > 50
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k


def test_single_sort_read_lists_equal_two_pass(hostsim, tmp_path):
    """Read lists sorted once on (block, BAM, haplotype, rank of the variant in its block) == variant sort + row sort."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 50, 500, 9000, n_bams=3, switch_per_base=0.02)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams(haplo_count_bam_exclude=[2], max_block_size=6)
    out = []
    for two in (0, 1):
        hostsim.set_option("two_pass_read_lists", two)
        try:
            out.append(pipeline.run_path(hostsim, vt, [hostsim.upload_reads(b) for b in batches], P, n_fragments=len(fd.names)))
        finally:
            hostsim.set_option("two_pass_read_lists", 0)
    a, b = out
    assert a.counters == b.counters and a.counters["read_list_entries"] > 100
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k


def test_fragment_table_graph_equals_the_sort_based_graph(hostsim, tmp_path):
    """graph_mode 1 (fragment table: rank, scan, scatter, one thread per fragment, pair hash) against graph_mode 0 (global
    tuple sort, entry / group arrays, pair sort): identical result arrays and counters -- with the pair table started
    tiny so that it overflows and is grown, and with three BAMs sharing read names."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 52, 500, 9000, n_bams=3, switch_per_base=0.02, insert_lo=60, insert_hi=220)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    for b in batches[1:]:                    # the same fragment ids in every BAM: shared QNAMEs (Q9)
        b.frag = (b.frag % np.uint32(len(batches[0].qnames))).astype(np.uint32)
    P = pipeline.PhaseParams(haplo_count_bam_exclude=[1])
    out = []
    for mode, slots in ((1, 16), (1, 1 << 20), (0, 1 << 20)):
        hostsim.set_option("graph_mode", mode); hostsim.set_option("pair_table_slots", slots)
        try:
            out.append(pipeline.run_path(hostsim, vt, [hostsim.upload_reads(b) for b in batches], P, n_fragments=len(fd.names)))
            # the three commits ranked the tuples inside their fragments; the graph stage did not have to
            assert hostsim.get_option("graph_ranked_in_commit") == (1 if mode == 1 else 0)
        finally:
            hostsim.set_option("graph_mode", 1); hostsim.set_option("pair_table_slots", 1 << 20)
    a, b, c = out
    assert a.counters["edges"] > 100
    for x in (b, c):
        for k in ("n_tuples", "entries", "groups", "pairs", "distinct_pairs", "edges", "dropped", "members", "final_blocks",
                  "read_list_entries"):
            assert a.counters[k] == x.counters[k], k
        for k in a.arrays:
            assert np.array_equal(a.arrays[k], x.arrays[k]), k


def test_ranks_taken_by_the_commits_equal_ranks_taken_by_the_graph_stage(hostsim, tmp_path):
    """With the fragment count announced (option "n_fragments") every commit ranks its tuples inside their fragments while it
    writes them, and the graph stage skips its ranking pass; without it the graph stage ranks.  Same graph either way, also
    when the announced count is not the one the graph stage is then given (the stage ranks itself)."""
    vcf, sams = util.make_case(tmp_path, 54, 400, 6000, n_bams=2, switch_per_base=0.02)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    nf = len(fd.names)
    out = []
    for announced, expect in ((nf, 1), (0, 0), (nf // 2, 0)):
        hostsim.set_variants(vt); hostsim.set_option("n_fragments", announced)
        for bi, rb in enumerate(batches):
            hostsim.map_reads(hostsim.upload_reads(rb), 10, 0.0); hostsim.commit_bam(bi, None)
        hostsim.variant_stats()
        hostsim.build_graph(nf, 0)
        assert hostsim.get_option("graph_ranked_in_commit") == expect
        out.append((hostsim.counters(), {k: hostsim.download(k).copy() for k in ("ed_a", "ed_b", "ed_sup", "ed_tot", "ed_cfg", "setsize", "vb_cnt")}))
        # asking again on the same commits ranks afresh (the fragment kernel reused the count array)
        hostsim.build_graph(nf, 0)
        assert hostsim.get_option("graph_ranked_in_commit") == 0
        again = {k: hostsim.download(k).copy() for k in out[-1][1]}
        for k in again:
            assert np.array_equal(again[k], out[-1][1][k]), k
    assert out[0][0]["edges"] > 50 and out[0][0]["full_sort_fallback"] == 0
    for c, a in out[1:]:
        for k in ("n_tuples", "entries", "groups", "pairs", "distinct_pairs", "edges"):
            assert c[k] == out[0][0][k], k
        for k in a:
            assert np.array_equal(a[k], out[0][1][k]), k


def test_a_huge_fragment_takes_the_sort_based_graph(hostsim, tmp_path):
    """More than 65535 tuples under ONE read name do not fit the 16-bit in-fragment rank: the stage falls back to the
    sort-based graph (full-key sort), loudly in the counters, with the same arrays as asking for that stage outright."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 53, 600, 100000, n_bams=1, switch_per_base=0.0)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    batches[0].frag = np.zeros_like(batches[0].frag)
    P = pipeline.PhaseParams(max_block_size=0)
    out = []
    for mode, announced in ((1, 1), (1, 0), (0, 0)):      # ranks taken by the commit / by the graph stage / not at all
        hostsim.set_option("graph_mode", mode)
        try:
            hostsim.set_variants(vt); hostsim.set_option("n_fragments", announced)
            for bi, rb in enumerate(batches):
                hostsim.map_reads(hostsim.upload_reads(rb), 10, 0.0); hostsim.commit_bam(bi, None)
            hostsim.variant_stats()
            hostsim.build_graph(1, 0)
            out.append((hostsim.counters(), {k: hostsim.download(k).copy() for k in ("ed_a", "ed_b", "ed_sup", "ed_tot", "ed_cfg", "setsize", "vb_cnt")}))
        finally:
            hostsim.set_option("graph_mode", 1)
    (ca, a), (cc, c), (cb, b) = out
    assert ca["n_tuples"] > 65535 and ca["full_sort_fallback"] == 1 and cb["full_sort_fallback"] == 1
    assert ca == cb and cc == cb
    for k in a:
        assert np.array_equal(a[k], b[k]) and np.array_equal(c[k], b[k]), k


def test_variant_stats_in_halves_returns_the_same_sums(hostsim, tmp_path):
    """phz_variant_stats == phz_variant_stats_async + phz_noise_wait == phz_variant_stats_device + phz_noise_publish +
    phz_noise_wait (the forms run_path uses on one rank / on several ranks), and a wait without a pending half fails."""
    import torch
    from phaser_b200 import engine as eng
    vcf, sams = util.make_case(tmp_path, 44, 200, 2000, n_bams=1)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    hostsim.set_variants(vt)
    hostsim.map_reads(hostsim.upload_reads(batches[0]), 10, 0.0); hostsim.commit_bam(0, None)
    whole = hostsim.variant_stats()
    assert whole[0] > 0
    hostsim.variant_stats_async()
    assert hostsim.noise_wait() == whole
    t = torch.zeros(2, dtype=torch.int64)
    hostsim.variant_stats_device(t)
    assert tuple(int(x) for x in t.tolist()) == whole
    t += 5                                   # what an all-reduce over the ranks would leave behind
    hostsim.noise_publish(t)
    assert hostsim.noise_wait() == (whole[0] + 5, whole[1] + 5)
    with pytest.raises(eng.PhzError):
        hostsim.noise_wait()
