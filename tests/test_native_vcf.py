"""Native host VCF ingest / output (include/phz.h: phz_vcf_*) against the Python implementations they replace in the
command line -- which the parity tests pin against the reference -- on every golden input, and against the reference's
own output VCF where a fixture holds one."""
import ctypes
import gzip

import numpy as np
import pytest

from phaser_b200 import engine as eng, pipeline, vcfio, writer
from tests import util, golden_util as G


@pytest.fixture(scope="module")
def lib():
    return eng._declare(ctypes.CDLL(util.build_hostsim()))


def _kw(c):
    kw = G.args_to_kw(c["meta"]["args"])
    return kw, dict(pass_only=kw.get("pass_only", 1), id_separator=kw.get("id_separator", "_"), include_indels=kw.get("include_indels", 0),
                    gw_phase_method=kw.get("gw_phase_method", 0), chrom_of_interest=kw.get("chr", ""))


@pytest.mark.parametrize("name", G.case_names())
def test_native_parser_equals_python_parser(lib, name):
    c = G.load_case(name)
    _, k = _kw(c)
    col = vcfio.sample_column_map(c["vcf"])["S1"]
    a, sa = vcfio.parse_vcf(c["vcf"], col, **k)
    nv = vcfio.NativeVcf(c["vcf"], lib, threads=3)
    assert nv.sample_column_map() == vcfio.sample_column_map(c["vcf"])
    b, sb = vcfio.parse_vcf_native(nv, col, **k)
    assert sa == sb and a.contigs == b.contigs
    for f in ("contig_var_off", "pos", "a0", "a1", "ref_len"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for f in ("ids", "rsids", "all_alleles", "gt", "maf"):
        assert list(getattr(a, f)) == list(getattr(b, f)), f


@pytest.mark.parametrize("name", G.case_names())
def test_native_vcf_writer_equals_python_writer_and_reference(lib, hostsim, name):
    c = G.load_case(name)
    kw, k = _kw(c)
    if kw.get("blacklist") or kw.get("haplo_count_blacklist"):
        pytest.skip("blacklists go through the Python parser (the command line does the same)")
    col = vcfio.sample_column_map(c["vcf"])["S1"]
    nv = vcfio.NativeVcf(c["vcf"], lib, threads=3)
    vt, st = vcfio.parse_vcf_native(nv, col, **k)
    from phaser_b200 import samio
    fd = samio.FragmentDictionary()
    mq = [int(x) for x in c["meta"]["mapq"].split(",")] * len(c["sams"]); pe = [int(x) for x in c["meta"]["paired_end"].split(",")] * len(c["sams"])
    batches = [samio.parse_sam(s, vt.contigs, fd, bool(kw.get("remove_dups", 1)), bool(pe[i]), mq[i]) for i, s in enumerate(c["sams"])]
    P = pipeline.PhaseParams(isize=kw.get("isize", [0.0]), as_q_cutoff=kw.get("as_q_cutoff", 0.05), cc_threshold=kw.get("cc_threshold", 0.01),
                             max_block_size=kw.get("max_block_size", 15), haplo_count_bam_exclude=kw.get("exclude", []))
    res = pipeline.run_path(hostsim, vt, [hostsim.upload_reads(b) for b in batches], P, n_fragments=len(fd.names))
    for mode in sorted({kw.get("gw_phase_vcf", 0), 0, 1, 2}):
        for conf in (kw.get("gw_phase_vcf_min_confidence", 0.90), 0.6):
            o = writer.Outputs(res, vt, util.bam_display_names(c["sams"]), P, unphased_vars=kw.get("unphased_vars", 1),
                               gw_phase_method=kw.get("gw_phase_method", 0), unique_ids=kw.get("unique_ids", 0))
            o.allelic_counts(); o.block_tables()
            with gzip.open(c["vcf"], "rt") as f:
                exp, up, pc = o.vcf_text(f.readlines(), col, id_separator=k["id_separator"], gw_phase_vcf=mode, min_conf=conf,
                                         chrom_of_interest=k["chrom_of_interest"])
            got, up2, pc2, rec = o.vcf_native(nv, gw_phase_vcf=mode, min_conf=conf, chrom_of_interest=k["chrom_of_interest"],
                                              id_separator=k["id_separator"])
            assert bytes(got).decode() == exp, (mode, conf)
            assert (up, pc) == (up2, pc2)
            ch, beg, end = o.vcf_records
            assert [rec[1][i] for i in rec[0].tolist()] == ch and rec[2].tolist() == beg and rec[3].tolist() == end
            if mode == kw.get("gw_phase_vcf", 0) and conf == kw.get("gw_phase_vcf_min_confidence", 0.90):
                assert bytes(got).decode() == c["ref"]["vcf"]          # the unmodified reference's own output VCF


@pytest.mark.parametrize("name", ["rna_two_bams", "quirks", "fuzz_indels"])
def test_native_bgzip_and_index_equal_the_python_writers(lib, hostsim, tmp_path, name):
    """phz_vcf_save (BGZF blocks deflated in parallel, .tbi / .csi) writes byte for byte what bgzf.py + tabix.py write --
    which tests/test_cli_and_io.py checks by region queries through the index."""
    from phaser_b200 import tabix, samio
    c = G.load_case(name)
    kw, k = _kw(c)
    col = vcfio.sample_column_map(c["vcf"])["S1"]
    nv = vcfio.NativeVcf(c["vcf"], lib, threads=3)
    vt, st = vcfio.parse_vcf_native(nv, col, **k)
    fd = samio.FragmentDictionary()
    batches = [samio.parse_sam(s, vt.contigs, fd, True, True, int(c["meta"]["mapq"].split(",")[0])) for s in c["sams"]]
    P = pipeline.PhaseParams(as_q_cutoff=kw.get("as_q_cutoff", 0.05), haplo_count_bam_exclude=kw.get("exclude", []))
    res = pipeline.run_path(hostsim, vt, [hostsim.upload_reads(b) for b in batches], P, n_fragments=len(fd.names))
    o = writer.Outputs(res, vt, util.bam_display_names(c["sams"]), P)
    o.allelic_counts(); o.block_tables()
    text, _, _, rec = o.vcf_native(nv)
    text = bytes(text)
    for csi in (False, True):
        a = str(tmp_path / ("n%d.vcf.gz" % csi)); b = str(tmp_path / ("p%d.vcf.gz" % csi))
        o.vcf_save_native(nv, a, csi=csi)
        tabix.write_vcf_with_index(b, text, csi=csi)
        ext = ".csi" if csi else ".tbi"
        assert open(a, "rb").read() == open(b, "rb").read()
        assert open(a + ext, "rb").read() == open(b + ext, "rb").read()


@pytest.mark.parametrize("name", G.case_names())
def test_site_text_gives_the_same_variant_records(lib, name):
    """phz_vcf_site_text (one native call for every site the tables name) -> writer.VariantMeta.from_site_text must be the
    record writer.VariantMeta builds from the Python reading of the same VCF line, on every site of every golden input
    (multi-allelic sites, unphased / phased genotypes, missing IDs, other id separators)."""
    c = G.load_case(name)
    kw, k = _kw(c)
    if k["gw_phase_method"] == 1:
        pytest.skip("allele frequencies are read by the Python path (prefetch_sites leaves those runs alone)")
    col = vcfio.sample_column_map(c["vcf"])["S1"]
    nv = vcfio.NativeVcf(c["vcf"], lib, threads=3)
    vt, st = vcfio.parse_vcf_native(nv, col, **k)
    rows = vt.ids.o
    rows.prefetch_sites(np.arange(vt.n_variants)[::-1])
    rows.prefetch_sites(np.arange(vt.n_variants))               # a second request for known sites asks for nothing
    assert len(rows.site_rows) == vt.n_variants
    contig_of = np.repeat(np.arange(len(vt.contigs)), np.diff(vt.contig_var_off))
    n_native = 0
    for v in range(vt.n_variants):
        chrom = vt.contigs[int(contig_of[v])]
        want = writer.VariantMeta(vt, v, chrom)
        got = writer.VariantMeta.from_site_text(rows.site_rows[v], chrom, rows.sep, int(vt.pos[v]))
        if got is None:
            continue
        n_native += 1
        for f in writer.VariantMeta.__slots__:
            assert getattr(got, f) == getattr(want, f), (v, f, rows.site_rows[v])
    assert n_native > 0 or vt.n_variants == 0


def test_site_text_on_hand_written_sites(lib, tmp_path):
    """Missing ID, a FORMAT with GT in second place, a multi-allelic unphased site and a ninth ALT allele."""
    alts = ",".join("ACGT"[i % 4] * (i + 2) for i in range(9))
    lines = ["##fileformat=VCFv4.2\n", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\n",
             "1\t100\t.\tA\tG\t.\tPASS\t.\tGT\t0|1\n",
             "1\t200\trs2\tA\t%s\t.\tPASS\t.\tGT:DP\t9|0:5\n" % alts,
             "1\t300\trs3\tA\tC,G\t.\tPASS\t.\tDP:GT\t7:2/1\n"]
    p = str(tmp_path / "odd.vcf.gz")
    with gzip.open(p, "wt") as f:
        f.writelines(lines)
    nv = vcfio.NativeVcf(p, lib, threads=2)
    vt, st = vcfio.parse_vcf_native(nv, 9, include_indels=1)
    rows = vt.ids.o
    rows.prefetch_sites(np.arange(vt.n_variants))
    got = {int(vt.pos[v]): rows.site_rows[v] for v in range(vt.n_variants)}
    assert got[100] == "100\t.\tA\tG\tA\tG\tA\tG"
    last = alts.split(",")[-1]
    assert got[200] == "200\trs2\tA\t%s\tA\t%s\t%s\tA" % (alts, last, last)
    assert got[300] == "300\trs3\tA\tC,G\tC\tG\t-\t-"
    for pos in (100, 200, 300):
        m = writer.VariantMeta.from_site_text(got[pos], "1", "_", pos)
        w = writer.VariantMeta(vt, [int(x) for x in vt.pos].index(pos), "1")
        assert all(getattr(m, f) == getattr(w, f) for f in writer.VariantMeta.__slots__), pos
