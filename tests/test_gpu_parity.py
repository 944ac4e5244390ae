"""Parity tests proper: the CUDA library (phaser_b200/_phz.so, sm_100a) through the C ABI against the
reference fixtures and the oracle.  Run on the B200 box:  python -m pytest tests -m gpu"""
import numpy as np
import pytest

from oracle import compare, port
from tests import util, golden_util as G

pytestmark = pytest.mark.gpu


def test_backend_is_the_cuda_library(gpu):
    assert gpu.backend == "cuda-sm_100a"


@pytest.mark.parametrize("name", G.case_names())
def test_files_match_reference(gpu, name):
    c = G.load_case(name)
    kw = G.args_to_kw(c["meta"]["args"])
    got, res, _ = util.product_outputs(gpu, c["vcf"], c["sams"], mapq=c["meta"]["mapq"], paired_end=c["meta"]["paired_end"], **kw)
    bad = compare.diff_outputs(c["ref"], got)
    assert not bad, "\n".join(bad)
    own, lib = gpu.launch_counts()
    assert own > 0


@pytest.mark.parametrize("name", G.case_names())
def test_mapper_tuples_match_oracle(gpu, name):
    c = G.load_case(name)
    kw = G.args_to_kw(c["meta"]["args"])
    vt, st, batches, col, fd = util.load_inputs(c["vcf"], c["sams"], mapq=c["meta"]["mapq"], paired_end=c["meta"]["paired_end"],
                                                remove_dups=kw.get("remove_dups", 1), pass_only=kw.get("pass_only", 1),
                                                id_separator=kw.get("id_separator", "_"), gw_phase_method=kw.get("gw_phase_method", 0),
                                                include_indels=kw.get("include_indels", 0))
    for batch in batches:
        got, exp = util.compare_tuples(gpu, vt, batch)
        assert got == exp


@pytest.mark.parametrize("seed,n_bams,switch,mbs", [(31, 1, 0.002, 15), (32, 2, 0.01, 15), (33, 1, 0.03, 4),
                                                     (34, 2, 0.05, 3), (35, 1, 0.03, 5), (36, 3, 0.01, 15)])
def test_seeded_cases_match_oracle(gpu, tmp_path, seed, n_bams, switch, mbs):
    vcf, sams = util.make_case(tmp_path, seed, 250, 2500, n_bams=n_bams, switch_per_base=switch)
    got, res, _ = util.product_outputs(gpu, vcf, sams, max_block_size=mbs)
    exp, ores = util.oracle_outputs(vcf, sams, max_block_size=mbs)
    bad = compare.diff_outputs(exp, got)
    assert not bad, "\n".join(bad)
    assert res.counters["n_tuples"] == ores.total_tuples
    assert res.noise_e == ores.noise_e


def test_larger_case_matches_oracle(gpu, tmp_path):
    """~40k records over 3 contigs: long enough for multi-block grids, short enough for the Python oracle."""
    vcf, sams = util.make_case(tmp_path, 41, 2500, 20000, contigs=[("20", 900000), ("21", 700000), ("22", 500000)],
                               switch_per_base=0.004)
    got, res, _ = util.product_outputs(gpu, vcf, sams)
    exp, ores = util.oracle_outputs(vcf, sams)
    bad = compare.diff_outputs(exp, got)
    assert not bad, "\n".join(bad)


def test_full_path_is_deterministic_and_consistent(gpu, tmp_path):
    """Size-independent properties on a case too big for the oracle: two runs give identical arrays;
    list lengths add up to the tuple count; every final block has >= 2 variants on one contig, sorted;
    per-BAM haplotype counts never exceed the all-BAM counts."""
    from phaser_b200 import synth, pipeline
    g = synth.make_genome(51, 40000, contigs=synth.GRCH38[18:22], n_genes=5000)
    vt = synth.to_variant_table(g)
    recs = [synth.filter_raw(synth.make_reads(g, 5100 + b, 300000), True, True, 0) for b in range(2)]
    P = pipeline.PhaseParams()
    outs = []
    for rep in range(2):
        batches = []
        nfrag = 0
        for b, rec in enumerate(recs):
            rb = synth.to_read_batch(rec, len(vt.contigs), "b%d" % b)
            rb.frag = (rb.frag + nfrag).astype(np.uint32); nfrag += len(rb.qnames)
            batches.append(gpu.upload_reads(rb))
        outs.append(pipeline.run_path(gpu, vt, batches, P, n_fragments=nfrag))
    a, b = outs
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k
    assert int(a.ncls.sum()) == a.counters["n_tuples"]
    assert a.counters["final_blocks"] > 100
    contig_of = np.searchsorted(vt.contig_var_off, np.arange(vt.n_variants), side="right") - 1
    for f in range(a.fb_first.shape[0]):
        m = a.members[a.fb_first[f]:a.fb_first[f] + a.fb_len[f]]
        assert m.shape[0] >= 2 and (np.diff(m.astype(np.int64)) > 0).all() and len(set(contig_of[m].tolist())) == 1
    fc = a.fb_cnt.reshape(-1, 2); fbc = a.fb_bcnt.reshape(-1, 2, 2)
    assert (fbc.sum(1) >= fc).all() and (fbc.max(1) <= fc).all()


def test_windowed_k1_equals_generic_k1(gpu):
    """The shared-memory/TMA K1 and the plain one-thread-per-record K1 must emit identical tuples
    (whole-genome contig layout, ~1.2 M records, spliced + indels + clips)."""
    from phaser_b200 import synth
    g = synth.make_genome(61, 60000, exonic_frac=0.3, n_genes=4000)
    vt = synth.to_variant_table_arrays(g)
    rec = synth.make_reads(g, 6100, 600000)
    rb = synth.to_read_batch(rec, len(vt.contigs), "b0")
    d = gpu.upload_reads(rb)
    out = {}
    try:
        for mode in (0, 1, 2, 3, 4):          # 4: the tile kernel with every candidate re-derived by the dense emission
            gpu.set_option("k1_mode", min(mode, 3)); gpu.set_option("k1_staged_emit", 0 if mode == 4 else 1)
            gpu.set_variants(vt)
            n = gpu.map_reads(d, 10, 0.0)
            out[mode] = (n, gpu.download("t_rec"), gpu.download("t_var"), gpu.download("t_misc"))
    finally:
        gpu.set_option("k1_mode", 3); gpu.set_option("k1_staged_emit", 1)
    assert out[0][0] == out[1][0] == out[2][0] == out[3][0] == out[4][0] and out[0][0] > 100000
    for m in (1, 2, 3, 4):
        for a, b in zip(out[0][1:], out[m][1:]):
            assert np.array_equal(a, b), m


def test_wgs_plus_rna_joint_matches_oracle(gpu, tmp_path):
    """configs[3] shape at oracle scale: a WGS BAM (2x150, unspliced, MAPQ filter 20) + an RNA BAM, per-BAM
    mapq / paired_end lists, haplotypic counts only for the RNA BAM."""
    from phaser_b200 import synth
    contigs = [("21", 200000), ("22", 150000)]
    g = synth.make_genome(71, 500, contigs=contigs, n_genes=40)
    vcf = synth.write_vcf(g, str(tmp_path / "j.vcf.gz"))
    wgs = synth.make_wgs_reads(g, 7100, 6000, mapq=60, lowmapq_frac=0.05, dup_frac=0.03)
    rna = synth.make_reads(g, 7101, 3000, dup_frac=0.05)
    sams = [synth.write_sam(wgs, g, str(tmp_path / "wgs.bam"), "w"), synth.write_sam(rna, g, str(tmp_path / "rna.bam"), "r")]
    kw = dict(mapq="20,255", paired_end="1,1", exclude=[0])
    got, res, _ = util.product_outputs(gpu, vcf, sams, **kw)
    exp, ores = util.oracle_outputs(vcf, sams, **kw)
    bad = compare.diff_outputs(exp, got)
    assert not bad, "\n".join(bad)
    assert res.counters["n_tuples"] == ores.total_tuples > 3000


def test_output_does_not_depend_on_the_sharding_gpu(gpu, tmp_path):
    """Shard-merge invariance at a size the oracle cannot reach: 1 shard vs 4 logical shards (4 contexts on this
    GPU, threads, exact reductions in between) over 24 contigs -> identical result arrays after the merge."""
    from phaser_b200 import synth, pipeline, shard, engine as eng
    g = synth.make_genome(91, 60000, exonic_frac=0.4, n_genes=5000)
    vt = synth.to_variant_table_arrays(g)
    batches = []
    nfrag = 0
    for b in range(2):
        rb = synth.to_read_batch(synth.make_reads(g, 9100 + b, 300000), len(vt.contigs), "b%d" % b)
        rb.frag = (rb.frag + nfrag).astype(np.uint32); nfrag += len(rb.qnames)
        batches.append(rb)
    P = pipeline.PhaseParams()
    one = pipeline.run_path(gpu, vt, [gpu.upload_reads(b) for b in batches], P, n_fragments=nfrag)
    many = shard.run_logical_shards(lambda: eng.Engine(device="cuda:0"), vt, batches, P, nfrag, 4)
    merged_one = shard.merge_results([(one, np.arange(vt.n_variants, dtype=np.int64), list(range(len(vt.contigs))))], vt, 2)
    for k in ("ncls", "setsize", "vb_cnt", "v_final", "v_hap", "members", "fb_first", "fb_len", "fb_sup", "fb_tot", "fb_cnt",
              "fb_bcnt", "rl_row", "rl_var", "rl_frag"):
        assert np.array_equal(merged_one.arrays[k], many.arrays[k]), k
    assert np.array_equal(np.argsort(merged_one.vfirst, kind="stable"), np.argsort(many.vfirst, kind="stable"))
    assert many.counters["final_blocks"] == one.counters["final_blocks"] > 500


def test_config1_shape_matches_reference_outputs(gpu, tmp_path):
    """BASELINE.json configs[0] shape (chr22, ~10k het SNVs, 300k read pairs = 600k SAM records) against the
    outputs of the UNMODIFIED reference (tests/golden/config1, made by tests/golden/make_golden.py).  The inputs
    are regenerated from the seed and checked by hash; ~80 blocks go through the exhaustive phase_v3 path."""
    import gzip, json, os, sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_golden
    meta = json.load(open(os.path.join(here, "config1", "case.json")))
    vcf, sam, digest = make_golden.config1_inputs(str(tmp_path))
    # a drifted generator must not look green: the fixture pins the reference's outputs for exactly these inputs
    assert digest == meta["sha256_inputs"], "seeded generator produced different inputs on this box (torch CPU generator drift): " \
                                            "regenerate tests/golden/config1 with tests/golden/make_golden.py"
    ref = {k: gzip.open(os.path.join(here, "config1", "ref." + k + (".txt.gz" if k != "vcf" else ".gz")), "rt").read()
           for k in ("allelic_counts", "allele_config", "haplotypes", "haplotypic_counts", "variant_connections", "vcf")}
    got, res, _ = util.product_outputs(gpu, vcf, [sam])
    bad = compare.diff_outputs(ref, got)
    assert not bad, "\n".join(bad)
    assert res.counters["hard_blocks"] > 20


def test_totals_above_the_precomputed_range_take_the_side_list(gpu, tmp_path, monkeypatch):
    """Critical values of c_total values above pipeline.PRECOMPUTED_TOTALS come from the device's side list
    (big_tot); with the range shrunk to 3 nearly every tested edge goes that way and nothing may change."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 36, 250, 2500, n_bams=2, switch_per_base=0.03)
    exp, _ = util.oracle_outputs(vcf, sams, max_block_size=5)
    monkeypatch.setattr(pipeline, "PRECOMPUTED_TOTALS", 3)
    got, res, _ = util.product_outputs(gpu, vcf, sams, max_block_size=5)
    bad = compare.diff_outputs(exp, got)
    assert not bad, "\n".join(bad)
    assert res.counters["edges"] > 0


@pytest.mark.parametrize("n_quals,expect_bits", [(2, 1), (4, 2), (11, 4), (40, 8)])
def test_packed_transport_is_lossless(gpu, tmp_path, n_quals, expect_bits):
    """phz_map_reads_packed (H2D of the packed form + expansion kernels + K1) == phz_map_reads on the plain arrays."""
    vcf, sams = util.make_case(tmp_path, 41, 300, 20000, n_bams=1)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    b = batches[0]
    rng = np.random.default_rng(n_quals)
    b.qual = rng.choice(np.arange(2, 2 + 2 * n_quals, 2), size=b.qual.shape[0]).astype(np.uint8)
    seq = b.seq.copy()
    hit = rng.random(seq.shape[0]) < 0.02
    seq[hit] = rng.integers(0, 256, size=int(hit.sum()), dtype=np.uint8)
    b.seq = seq
    p = util.packed_vs_plain(gpu, vt, b, len(vt.contigs))
    assert p.qual_bits == expect_bits and p.n_exceptions > 0


def test_packed_host_form_through_the_whole_path(gpu, tmp_path):
    """run_path fed with PackedReads (what the command line does) == run_path on uploaded arrays."""
    from phaser_b200 import pipeline, engine as eng
    vcf, sams = util.make_case(tmp_path, 42, 300, 6000, n_bams=2, switch_per_base=0.02)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams()
    a = pipeline.run_path(gpu, vt, [gpu.upload_reads(b) for b in batches], P, n_fragments=len(fd.names))
    b = pipeline.run_path(gpu, vt, [eng.pack_reads(x, len(vt.contigs), lib=gpu.lib) for x in batches], P, n_fragments=len(fd.names))
    assert a.counters == b.counters
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k


def test_result_buffer_views_equal_private_copies(gpu, tmp_path):
    """download_many (page-locked buffer, async copies, one wait) returns what the per-array downloads return."""
    from phaser_b200 import pipeline, engine as eng
    vcf, sams = util.make_case(tmp_path, 43, 300, 6000, n_bams=1)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams()
    a = pipeline.run_path(gpu, vt, [gpu.upload_reads(b) for b in batches], P, n_fragments=len(fd.names))
    b = pipeline.run_path(gpu, vt, [eng.pack_reads(x, len(vt.contigs), lib=gpu.lib) for x in batches], P,
                          n_fragments=len(fd.names), reuse_result_buffer=True)
    assert set(a.arrays) == set(b.arrays)
    for k in a.arrays:
        assert a.arrays[k].dtype == b.arrays[k].dtype and np.array_equal(a.arrays[k], b.arrays[k]), k


def test_fragment_sort_fixup_and_fallback_agree(gpu, tmp_path):
    """The graph stage sorts tuples on the fragment bits only and repairs interleaved runs in place; with
    frag_run_limit = 1 every multi-tuple fragment takes the full-key sort instead.  Same arrays either way, and the
    case (short inserts: overlapping mates) must exercise the in-place repair."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 44, 400, 8000, n_bams=2, switch_per_base=0.02, insert_lo=60, insert_hi=200)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams()
    gpu.set_option("graph_mode", 0)           # the sort-based stage (A/B switch and fallback of the fragment-table stage)
    try:
        a = pipeline.run_path(gpu, vt, [gpu.upload_reads(b) for b in batches], P, n_fragments=len(fd.names))
        assert a.counters["frag_runs_resorted"] > 0 and a.counters["full_sort_fallback"] == 0
        gpu.set_option("frag_run_limit", 1)
        b = pipeline.run_path(gpu, vt, [gpu.upload_reads(b) for b in batches], P, n_fragments=len(fd.names))
    finally:
        gpu.set_option("frag_run_limit", 1024); gpu.set_option("graph_mode", 1)
    assert b.counters["full_sort_fallback"] == 1
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k
    exp, _ = util.oracle_outputs(vcf, sams)
    got, _, _ = util.product_outputs(gpu, vcf, sams)
    assert not compare.diff_outputs(exp, got)


def test_prefetched_samples_give_the_same_results(gpu, tmp_path):
    """Loop of samples with the next copy prefetched on the copy stream == each sample run on its own."""
    from phaser_b200 import pipeline, engine as eng
    P = pipeline.PhaseParams()
    samples = []
    for seed in (46, 47):
        vcf, sams = util.make_case(tmp_path, seed, 300, 20000, n_bams=1, switch_per_base=0.02)
        vt, st, batches, col, fd = util.load_inputs(vcf, sams)
        packed = eng.pack_reads(batches[0], len(vt.contigs), lib=gpu.lib, page_locked=True)
        alone = pipeline.run_path(gpu, vt, [gpu.upload_reads(batches[0])], P, n_fragments=len(fd.names))
        samples.append((vt, packed, len(fd.names), alone))
    gpu.prefetch_packed(samples[0][1])
    for it in range(6):
        vt, packed, nf, alone = samples[it % 2]
        gpu.prefetch_packed(samples[(it + 1) % 2][1])
        got = pipeline.run_path(gpu, vt, [packed], P, n_fragments=nf)
        assert got.counters == alone.counters
        for k in alone.arrays:
            assert np.array_equal(alone.arrays[k], got.arrays[k]), (it, k)
    vt, packed, nf, alone = samples[0]
    pipeline.run_path(gpu, vt, [packed], P, n_fragments=nf)          # drain


def test_packed_transport_wide_codings_round_trip(gpu, tmp_path):
    """Every field of the transport form on its wide coding and on its narrow one: the device expansion returns the
    original arrays bit for bit (st_* arrays) and K1 emits the same tuples."""
    from tests.test_hostsim_parity import _odd_batch
    vcf, sams = util.make_case(tmp_path, 41, 200, 300, n_bams=1)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    c = util.packed_vs_plain(gpu, vt, _odd_batch(len(vt.contigs)), len(vt.contigs)).coding
    assert c["cigar_bits"] == 32 and c["n_cigar_bits"] == 16 and c["l_seq_const"] == -1 and c["as_bits"] == 16
    c = util.packed_vs_plain(gpu, vt, batches[0], len(vt.contigs)).coding
    assert c["cigar_bits"] == 16 and c["n_cigar_bits"] == 8 and c["l_seq_const"] == 76 and c["as_bits"] == 8


@pytest.mark.parametrize("ids", ["first_appearance", "offset", "shuffled", "far_back", "gaps"])
def test_packed_fragment_ids_round_trip(gpu, tmp_path, ids):
    """The implicit fragment-id coding of the transport form (bitmap + 16-bit back references + exceptions) expanded by the
    CUDA library: same cases as on the host-simulation backend."""
    from tests import test_hostsim_parity as H
    H.test_packed_fragment_ids_round_trip(gpu, tmp_path, ids)


@pytest.mark.parametrize("n_bams", [1, 3, 5])
def test_window_aggregated_counters_equal_plain_atomics(gpu, tmp_path, n_bams):
    """The shared-memory window kernels for the per-variant counters (vfirst / ncls, set sizes, per-BAM allele
    counts; <= 4 BAMs windowed, more go to global memory) against the warp-aggregated global atomics."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 48, 600, 12000, n_bams=n_bams, switch_per_base=0.02)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams(haplo_count_bam_exclude=[1] if n_bams > 1 else [])
    out = []
    for mode, graph in ((1, 0), (0, 0), (1, 1)):          # (1, 1): the fragment-table stage's own CTA windows
        gpu.set_option("window_agg", mode); gpu.set_option("graph_mode", graph)
        try:
            out.append(pipeline.run_path(gpu, vt, [gpu.upload_reads(b) for b in batches], P, n_fragments=len(fd.names)))
        finally:
            gpu.set_option("window_agg", 1); gpu.set_option("graph_mode", 1)
    a, b, c = out
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], c.arrays[k]), k
    assert a.counters == b.counters
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k


def test_compact_pair_keys_equal_wide_ones(gpu, tmp_path):
    """Pair table sorted on (va << dbits | vb - va) (32-bit when it fits) == sorted on 64-bit keys."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 49, 500, 9000, n_bams=2, switch_per_base=0.02)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams()
    out = []
    for wide in (0, 1):
        gpu.set_option("wide_pair_keys", wide); gpu.set_option("graph_mode", 0)
        try:
            out.append(pipeline.run_path(gpu, vt, [gpu.upload_reads(b) for b in batches], P, n_fragments=len(fd.names)))
        finally:
            gpu.set_option("wide_pair_keys", 0); gpu.set_option("graph_mode", 1)
    a, b = out
    assert a.counters == b.counters and a.counters["edges"] > 50
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k


def test_single_sort_read_lists_equal_two_pass(gpu, tmp_path):
    """Read lists sorted once on (block, BAM, haplotype, rank of the variant in its block) == variant sort + row sort."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 50, 500, 9000, n_bams=3, switch_per_base=0.02)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams(haplo_count_bam_exclude=[2], max_block_size=6)
    out = []
    for two in (0, 1):
        gpu.set_option("two_pass_read_lists", two)
        try:
            out.append(pipeline.run_path(gpu, vt, [gpu.upload_reads(b) for b in batches], P, n_fragments=len(fd.names)))
        finally:
            gpu.set_option("two_pass_read_lists", 0)
    a, b = out
    assert a.counters == b.counters and a.counters["read_list_entries"] > 100
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k


def test_tile_order_commit_equals_canonical_commit(gpu, tmp_path):
    """The commit that compacts each tile of the K1 output to its canonical place (no permute of the candidates)
    against permute + scan + scatter; and the canonical candidate arrays materialised on demand are the same."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 52, 600, 40000, n_bams=2, switch_per_base=0.02)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams()
    out = []; tup = []
    for lazy in (1, 0):
        gpu.set_option("lazy_canonical", lazy)
        try:
            gpu.set_variants(vt)
            gpu.map_reads(gpu.upload_reads(batches[0]), 10, 0.0)
            tup.append([gpu.download(k).copy() for k in ("t_rec", "t_var", "t_misc")])
            out.append(pipeline.run_path(gpu, vt, [gpu.upload_reads(b) for b in batches], P, n_fragments=len(fd.names)))
        finally:
            gpu.set_option("lazy_canonical", 1)
    assert tup[0][0].shape[0] > 10000 and all(np.array_equal(a, b) for a, b in zip(*tup))
    a, b = out
    assert a.counters == b.counters
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k


def test_fragment_table_graph_equals_the_sort_based_graph(gpu, tmp_path):
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
        gpu.set_option("graph_mode", mode); gpu.set_option("pair_table_slots", slots)
        try:
            out.append(pipeline.run_path(gpu, vt, [gpu.upload_reads(b) for b in batches], P, n_fragments=len(fd.names)))
            # the three commits ranked the tuples inside their fragments; the graph stage did not have to
            assert gpu.get_option("graph_ranked_in_commit") == (1 if mode == 1 else 0)
        finally:
            gpu.set_option("graph_mode", 1); gpu.set_option("pair_table_slots", 1 << 20)
    a, b, c = out
    assert a.counters["edges"] > 100
    for x in (b, c):
        for k in ("n_tuples", "entries", "groups", "pairs", "distinct_pairs", "edges", "dropped", "members", "final_blocks",
                  "read_list_entries"):
            assert a.counters[k] == x.counters[k], k
        for k in a.arrays:
            assert np.array_equal(a.arrays[k], x.arrays[k]), k


@pytest.mark.parametrize("n_bams,fold", [(1, 0), (3, 0), (1, 11), (2, 5)])
def test_slot_chunk_fragment_kernel_equals_the_range_form(gpu, tmp_path, n_bams, fold):
    """frag_stage 1 (a CTA owns a chunk of tuple slots staged in shared memory, fragments found from the head marks)
    against frag_stage 0 (a CTA owns a range of fragment ids, one thread per fragment on global memory) and the sort-based
    stage: identical result arrays and counters over many chunks.  With the read names folded onto a handful of
    fragments every fragment holds thousands of tuples: their tails are not staged and the second pass takes them."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 58 + n_bams, 700, 30000, n_bams=n_bams, switch_per_base=0.02, insert_lo=60, insert_hi=220)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    nf = len(fd.names)
    if fold:
        for b in batches:
            b.frag = (b.frag % np.uint32(fold)).astype(np.uint32)
        nf = fold
    P = pipeline.PhaseParams(haplo_count_bam_exclude=[1] if n_bams > 1 else [])
    out = []
    for graph, stage in ((1, 1), (1, 0), (0, 0)):
        gpu.set_option("graph_mode", graph); gpu.set_option("frag_stage", stage)
        try:
            out.append(pipeline.run_path(gpu, vt, [gpu.upload_reads(b) for b in batches], P, n_fragments=nf))
            if graph == 1 and stage == 1:
                assert out[-1].counters["full_sort_fallback"] == 0
                assert (gpu.get_option("fragments_deferred") > 0) == bool(fold)
        finally:
            gpu.set_option("graph_mode", 1); gpu.set_option("frag_stage", 1)
    a, b, c = out
    assert a.counters["n_tuples"] > 20000 and a.counters["edges"] > 100
    assert a.counters == b.counters
    for x in (b, c):
        for k in ("n_tuples", "entries", "groups", "pairs", "distinct_pairs", "edges", "dropped", "members", "final_blocks",
                  "read_list_entries"):
            assert a.counters[k] == x.counters[k], k
        for k in a.arrays:
            assert np.array_equal(a.arrays[k], x.arrays[k]), k


def test_ranks_taken_by_the_commits_equal_ranks_taken_by_the_graph_stage(gpu, tmp_path):
    """With the fragment count announced (option "n_fragments") every commit ranks its tuples inside their fragments while it
    writes them, and the graph stage skips its ranking pass; without it the graph stage ranks.  Same graph either way, also
    when the announced count is not the one the graph stage is then given (the stage ranks itself)."""
    vcf, sams = util.make_case(tmp_path, 54, 400, 6000, n_bams=2, switch_per_base=0.02)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    nf = len(fd.names)
    out = []
    for announced, expect in ((nf, 1), (0, 0), (nf // 2, 0)):
        gpu.set_variants(vt); gpu.set_option("n_fragments", announced)
        for bi, rb in enumerate(batches):
            gpu.map_reads(gpu.upload_reads(rb), 10, 0.0); gpu.commit_bam(bi, None)
        gpu.variant_stats()
        gpu.build_graph(nf, 0)
        assert gpu.get_option("graph_ranked_in_commit") == expect
        out.append((gpu.counters(), {k: gpu.download(k).copy() for k in ("ed_a", "ed_b", "ed_sup", "ed_tot", "ed_cfg", "setsize", "vb_cnt")}))
        # asking again on the same commits ranks afresh (the fragment kernel reused the count array)
        gpu.build_graph(nf, 0)
        assert gpu.get_option("graph_ranked_in_commit") == 0
        again = {k: gpu.download(k).copy() for k in out[-1][1]}
        for k in again:
            assert np.array_equal(again[k], out[-1][1][k]), k
    assert out[0][0]["edges"] > 50 and out[0][0]["full_sort_fallback"] == 0
    for c, a in out[1:]:
        for k in ("n_tuples", "entries", "groups", "pairs", "distinct_pairs", "edges"):
            assert c[k] == out[0][0][k], k
        for k in a:
            assert np.array_equal(a[k], out[0][1][k]), k


def test_a_huge_fragment_takes_the_sort_based_graph(gpu, tmp_path):
    """More than 65535 tuples under ONE read name do not fit the 16-bit in-fragment rank: the stage falls back to the
    sort-based graph (full-key sort), loudly in the counters, with the same arrays as asking for that stage outright."""
    from phaser_b200 import pipeline
    vcf, sams = util.make_case(tmp_path, 53, 600, 100000, n_bams=1, switch_per_base=0.0)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    batches[0].frag = np.zeros_like(batches[0].frag)
    P = pipeline.PhaseParams(max_block_size=0)
    out = []
    for mode, announced in ((1, 1), (1, 0), (0, 0)):      # ranks taken by the commit / by the graph stage / not at all
        gpu.set_option("graph_mode", mode)
        try:
            gpu.set_variants(vt); gpu.set_option("n_fragments", announced)
            for bi, rb in enumerate(batches):
                gpu.map_reads(gpu.upload_reads(rb), 10, 0.0); gpu.commit_bam(bi, None)
            gpu.variant_stats()
            gpu.build_graph(1, 0)
            out.append((gpu.counters(), {k: gpu.download(k).copy() for k in ("ed_a", "ed_b", "ed_sup", "ed_tot", "ed_cfg", "setsize", "vb_cnt")}))
        finally:
            gpu.set_option("graph_mode", 1)
    (ca, a), (cc, c), (cb, b) = out
    assert ca["n_tuples"] > 65535 and ca["full_sort_fallback"] == 1 and cb["full_sort_fallback"] == 1
    assert ca == cb and cc == cb
    for k in a:
        assert np.array_equal(a[k], b[k]) and np.array_equal(c[k], b[k]), k


def test_staged_upload_equals_pageable_copy(gpu, tmp_path):
    """phz_upload (page-locked staging ring filled by host threads, what the command line uses for large BAMs) puts the
    same bytes on the device as plain pageable copies -- incl. arrays of several staging chunks and odd sizes."""
    import torch
    vcf, sams = util.make_case(tmp_path, 57, 300, 4000, n_bams=1)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    a = gpu.upload_reads(batches[0], staged=False); b = gpu.upload_reads(batches[0], staged=True)
    for k in a:
        if k != "contig_rec_off":
            assert torch.equal(a[k], b[k]), k
    big = np.random.default_rng(1).integers(0, 255, size=(3 * (32 << 20) + 12345,), dtype=np.uint8)
    assert torch.equal(gpu._upload(big).cpu(), torch.from_numpy(big))
    got, res, _ = util.product_outputs(gpu, vcf, sams)
    exp, _ = util.oracle_outputs(vcf, sams)
    assert not compare.diff_outputs(exp, got)
