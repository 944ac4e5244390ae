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
