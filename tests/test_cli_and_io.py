"""CLI drop-in behaviour and BAM/BGZF ingest.  Tests taking `engine` run on the host-simulation double in the build
container and on the CUDA library on the GPU box (tests/conftest.py)."""
import gzip
import os

import numpy as np
import pytest

from oracle import compare
from phaser_b200 import phaser as cli, bamio, bgzf, samio
from tests import util, golden_util as G


def _run_cli(engine, case, tmp_path, extra=()):
    c = G.load_case(case)
    o = str(tmp_path / "out")
    args = [os.path.join(os.path.dirname(G.CASES), a) if a.startswith("beds/") else a for a in c["meta"]["args"]]
    argv = ["--vcf", c["vcf"], "--bam", ",".join(c["sams"]), "--sample", "S1", "--mapq", c["meta"]["mapq"], "--baseq", "10",
            "--paired_end", c["meta"]["paired_end"], "--o", o] + args + list(extra)
    cli.run(cli.build_parser().parse_args(argv), engine=engine)
    got = {k: open(o + "." + k + ".txt").read() for k in ("allelic_counts", "allele_config", "haplotypes", "haplotypic_counts",
                                                          "variant_connections")}
    got["vcf"] = gzip.open(o + ".vcf.gz", "rt").read()
    for k in ("network_links", "network_nodes"):
        fn = o + "." + k.replace("_", ".") + ".txt"
        if os.path.exists(fn):
            got[k] = open(fn).read()
    return c, got


@pytest.mark.parametrize("case", ["quirks", "rna_two_bams", "opt_blacklists", "opt_maf_gwvcf2", "opt_nounphased_uid", "opt_filters",
                                  "indels", "fuzz_indels", "opt_read_ids", "opt_network", "q9_shared_qnames",
                                  "q26_same_basename", "opt_chr", "q16_tie_glue", "q22_unphased_blocks"])
def test_cli_writes_reference_identical_files(engine, tmp_path, case):
    c, got = _run_cli(engine, case, tmp_path)
    bad = compare.diff_outputs(c["ref"], got)
    assert not bad, "\n".join(bad)


def test_cli_fatal_errors_exit_1(engine, tmp_path, capsys):
    c = G.load_case("quirks")
    base = ["--vcf", c["vcf"], "--bam", c["sams"][0], "--mapq", "255", "--baseq", "10", "--paired_end", "1", "--o", str(tmp_path / "x")]
    for extra, msg in ((["--sample", "NOPE"], "Sample 'NOPE' not found"),
                       (["--sample", "S1", "--id_separator", ":"], "ID separator must not be"),
                       (["--sample", "S1", "--mapq", "1,2"], "Number of mapq values"),
                       (["--sample", "S1", "--blacklist", "x.bed"], "File: x.bed not found"),
                       (["--sample", "S1", "--process_slow", "1"], "not supported")):
        with pytest.raises(SystemExit) as e:
            cli.run(cli.build_parser().parse_args(base + extra), engine=engine)
        assert e.value.code == 1
        assert "FATAL ERROR: " in capsys.readouterr().out


def test_bam_reader_equals_sam_reader(tmp_path):
    """Write the golden SAM text as BAM with our writer, read it back: identical packed arrays."""
    c = G.load_case("rna_small")
    sam = c["sams"][0]
    refs = []; recs = []
    for ln in open(sam):
        f = ln.rstrip("\n").split("\t")
        if ln.startswith("@SQ"):
            refs.append((f[1][3:], int(f[2][3:])))
        elif not ln.startswith("@"):
            cig = []; n = ""
            for ch in f[5]:
                if ch.isdigit():
                    n += ch
                else:
                    cig.append((int(n), ch)); n = ""
            a = [int(t[5:]) for t in f[11:] if t.startswith("AS:i:")]
            recs.append((f[0], int(f[1]), [r[0] for r in refs].index(f[2]), int(f[3]), int(f[4]), cig, f[9],
                         bytes(ord(q) - 33 for q in f[10]), int(f[8]), a[0] if a else None))
    bam = str(tmp_path / "x.bam")
    bamio.write_bam(bam, refs, recs)
    contigs = [r[0] for r in refs]
    a = samio.read_alignments(sam, contigs, samio.FragmentDictionary(), True, True, 255)
    b = samio.read_alignments(bam, contigs, samio.FragmentDictionary(), True, True, 255)
    for k in ("contig_rec_off", "pos", "tlen", "aln_score", "frag", "cigar_off", "cigar", "seq_off", "seq", "qual"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert a.qnames == b.qnames


def test_bgzf_roundtrip_and_gzip_compat(tmp_path):
    p = str(tmp_path / "t.gz")
    payload = (b"line\tof\ttext\n" * 20000)
    with bgzf.BGZFWriter(p) as w:
        w.write(payload[:100000]); w.write(payload[100000:])
    assert gzip.open(p).read() == payload            # plain gzip readers accept BGZF
    assert bgzf.read_all(p) == payload
    raw = open(p, "rb").read()
    assert raw.endswith(bgzf.EOF_BLOCK) and raw[12:14] == b"BC"


@pytest.mark.parametrize("case", ["quirks", "rna_small", "indels", "fuzz_indels"])
def test_read_variant_map_seam_writes_the_reference_tsv(engine, tmp_path, monkeypatch, case):
    """Seam S1: do_read_variant_map(variant_table, baseq, o, splice, isize_cutoff) with SAM text on stdin
    produces byte for byte the TSV of the reference mapper (committed golden)."""
    import io
    from phaser_b200 import read_variant_map as rvm, vcfio
    c = G.load_case(case)
    col = vcfio.sample_column_map(c["vcf"])["S1"]
    vt, _ = vcfio.parse_vcf(c["vcf"], col, include_indels=G.args_to_kw(c["meta"]["args"]).get("include_indels", 0))
    table = tmp_path / "table.tsv"
    with open(table, "w") as f:
        for v in range(vt.n_variants):
            ci = int(np.searchsorted(vt.contig_var_off, v, side="right") - 1)
            f.write("\t".join([vt.contigs[ci], str(int(vt.pos[v])), vt.ids[v], vt.rsids[v], ",".join(vt.all_alleles[v]),
                               str(int(vt.ref_len[v])), vt.gt[v], vt.maf[v]]) + "\n")
    lines = []
    for ln in open(c["sams"][0]):          # what the two samtools stages would let through
        if ln[0] != "@":
            f_ = ln.split("\t")
            if int(f_[1]) & 0x400 or not int(f_[1]) & 2 or int(f_[4]) < 255:
                continue
        lines.append(ln)
    monkeypatch.setattr("sys.stdin", io.StringIO("".join(lines)))
    rvm.set_engine(engine)
    out = tmp_path / "out.tsv"
    rvm.do_read_variant_map(str(table), 10, str(out), 1, 0.0)
    assert open(out).read() == c["mapper"][c["meta"]["bams"][0]]


def test_native_reader_equals_python_readers(tmp_path):
    """phz_read_alignments (C++, parallel BGZF inflate) == the Python SAM / BAM readers, array for array,
    on SAM text, gzipped SAM and BAM, with and without filters."""
    import ctypes
    from phaser_b200 import engine as eng
    lib = eng._declare(ctypes.CDLL(util.build_hostsim()))
    c = G.load_case("rna_two_bams")
    contigs = ["21", "22"]
    refs = [("21", 120000), ("22", 80000)]
    for sam in c["sams"]:
        recs = []
        for ln in open(sam):
            f = ln.rstrip("\n").split("\t")
            if not ln.startswith("@"):
                cig = []; n = ""
                for ch in f[5]:
                    if ch.isdigit():
                        n += ch
                    else:
                        cig.append((int(n), ch)); n = ""
                a = [int(t[5:]) for t in f[11:] if t.startswith("AS:i:")]
                recs.append((f[0], int(f[1]), contigs.index(f[2]), int(f[3]), int(f[4]), cig, f[9],
                             bytes(ord(q) - 33 for q in f[10]), int(f[8]), a[0] if a else None))
        bam = str(tmp_path / (os.path.basename(sam) + ".real.bam"))
        bamio.write_bam(bam, refs, recs)
        gz = str(tmp_path / (os.path.basename(sam) + ".sam.gz"))
        with gzip.open(gz, "wt") as f:
            f.write(open(sam).read())
        for (rd, pp, mq) in ((True, True, 255), (False, False, 0)):
            ref_batch = samio.parse_sam(sam, contigs, samio.FragmentDictionary(), rd, pp, mq)
            for path in (sam, gz, bam):
                fd = eng.NativeFragmentDictionary(lib)
                b = eng.read_alignments_native(path, contigs, fd, rd, pp, mq, threads=3, lib=lib)
                for k in ("contig_rec_off", "pos", "tlen", "aln_score", "frag", "cigar_off", "cigar", "seq_off", "seq", "qual"):
                    assert np.array_equal(getattr(ref_batch, k), getattr(b, k)), (path, k)
                assert fd.names == ref_batch.qnames


def test_cli_empty_bam_is_the_reference_fatal_error(engine, tmp_path, capsys):
    c = G.load_case("rna_small")
    empty = str(tmp_path / "empty.bam")
    open(empty, "w").writelines(l for l in open(c["sams"][0]) if l[0] == "@")
    base = ["--vcf", c["vcf"], "--sample", "S1", "--mapq", "255", "--baseq", "10", "--paired_end", "1", "--o", str(tmp_path / "x")]
    with pytest.raises(SystemExit) as e:
        cli.run(cli.build_parser().parse_args(base + ["--bam", empty]), engine=engine)
    assert e.value.code == 1 and "No reads could be matched to variants" in capsys.readouterr().out
    cli.run(cli.build_parser().parse_args(base + ["--bam", empty + "," + c["sams"][0]]), engine=engine)   # an empty BAM beside a real one


@pytest.mark.parametrize("csi", [False, True])
def test_tabix_index_finds_exactly_the_overlapping_records(tmp_path, csi):
    """Region queries answered through the written index (bins + chunks + linear index, virtual offsets into the
    BGZF file) == brute-force filtering of the text; spans several BGZF blocks, contigs and bin levels."""
    import random
    from phaser_b200 import tabix
    rnd = random.Random(5)
    lines = ["##fileformat=VCFv4.2\n", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\n"]
    recs = []
    for chrom, L in (("1", 248_000_000), ("2", 3_000_000), ("X", 40_000)):
        pos = sorted(rnd.randrange(1, L) for _ in range(4000))
        for p in pos:
            ref = "A" * rnd.choice([1, 1, 1, 2, 30])
            info = "END=%d" % (p + 70000) if rnd.random() < 0.01 else "AF=0.1"
            lines.append("%s\t%d\t.\t%s\tC\t50\tPASS\t%s\tGT\t0|1:%s\n" % (chrom, p, ref, info, "x" * rnd.randrange(0, 60)))
            recs.append((chrom,) + tabix.record_span(lines[-1].split("\t", 8)) + (lines[-1],))
    path = str(tmp_path / "t.vcf.gz")
    idx_path = tabix.write_vcf_with_index(path, "".join(lines), csi=csi)
    assert idx_path.endswith(".csi" if csi else ".tbi")
    assert bgzf.read_all(path).decode() == "".join(lines)
    idx = tabix.read_index(idx_path)
    assert idx["names"] == ["1", "2", "X"] and idx["format"][:5] == (2, 1, 2, 0, ord("#"))
    assert len("".join(lines)) > 8 * bgzf.MAX_BLOCK         # several BGZF blocks
    for _ in range(300):
        chrom, L = rnd.choice([("1", 248_000_000), ("2", 3_000_000), ("X", 40_000), ("Y", 1000)])
        b = rnd.randrange(0, L); e = b + rnd.choice([1, 100, 20_000, 300_000, 50_000_000])
        exp = [r[3] for r in recs if r[0] == chrom and r[1] < e and r[2] > b]
        assert tabix.query(path, idx, chrom, b, e) == exp


def test_cli_writes_a_tabix_index_for_its_vcf(engine, tmp_path):
    from phaser_b200 import tabix
    c, got = _run_cli(engine, "rna_two_bams", tmp_path)
    o = str(tmp_path / "out")
    idx = tabix.read_index(o + ".vcf.gz.tbi")
    body = [l for l in got["vcf"].splitlines(keepends=True) if not l.startswith("#")]
    assert sum(r["bins"][tabix.META_BIN][1][0] for r in idx["refs"]) == len(body)
    first = body[0].split("\t")
    assert tabix.query(o + ".vcf.gz", idx, first[0], int(first[1]) - 1, int(first[1])) == [body[0]]
    # the spans handed over by the VCF writer give the same index as splitting the written lines again
    o2 = str(tmp_path / "again.vcf.gz")
    tabix.write_vcf_with_index(o2, got["vcf"])
    assert open(o2 + ".tbi", "rb").read() == open(o + ".vcf.gz.tbi", "rb").read()
    assert open(o2, "rb").read() == open(o + ".vcf.gz", "rb").read()


def test_vectorised_index_builder_equals_record_by_record():
    """IndexBuilder.add_many (array operations per reference) writes byte-identical .tbi / .csi to add() per record."""
    import random
    from phaser_b200 import tabix
    rnd = random.Random(3)
    for trial in range(4):
        A = tabix.IndexBuilder(); B = tabix.IndexBuilder()
        v = 4000                                          # data lines never start at virtual offset 0 (the header comes first)
        for chrom, L in (("1", 248_000_000), ("2", 3_000_000), ("X", 40_000)):
            pos = sorted(rnd.randrange(0, L) for _ in range(2500))
            beg = np.asarray(pos, np.int64)
            end = beg + np.asarray([rnd.choice([1, 1, 2, 30, 70000, 0]) for _ in pos], np.int64)
            sz = np.asarray([rnd.randrange(20, 200) for _ in pos], np.int64)
            u0 = v + np.concatenate([[0], np.cumsum(sz)[:-1]]); u1 = u0 + sz; v = int(u1[-1])
            f = lambda u: (((u // 0xff00) * 777) << 16) | (u % 0xff00)
            v0 = np.asarray([f(int(u)) for u in u0], np.int64); v1 = np.asarray([f(int(u)) for u in u1], np.int64)
            for k in range(len(pos)):
                A.add(chrom, int(beg[k]), int(end[k]), int(v0[k]), int(v1[k]))
            B.add_many(chrom, beg, end, v0, v1)
        assert A.tbi_bytes() == B.tbi_bytes() and A.csi_bytes() == B.csi_bytes()


def test_cli_with_the_packed_transport_form(engine, tmp_path, monkeypatch):
    """PHZ_PACK=1: the command line sends the BAMs through pack_reads / phz_map_reads_packed; same files."""
    monkeypatch.setenv("PHZ_PACK", "1")
    c, got = _run_cli(engine, "rna_two_bams", tmp_path)
    bad = compare.diff_outputs(c["ref"], got)
    assert not bad, "\n".join(bad)


def test_parallel_bgzf_blocks_equal_the_streaming_writer(tmp_path):
    import random
    rnd = random.Random(1)
    data = "".join("line %d %s\n" % (i, "x" * rnd.randrange(0, 90)) for i in range(40000)).encode()
    p = str(tmp_path / "a.gz")
    with bgzf.BGZFWriter(p) as w:
        w.write(data)
    assert b"".join(bgzf.compress_all(data, threads=4)) + bgzf.EOF_BLOCK == open(p, "rb").read()
    assert bgzf.read_all(p) == data


@pytest.mark.parametrize("case", ["rna_two_bams", "opt_read_ids", "rna_small"])
def test_soa_cache_gives_the_same_files(engine, tmp_path, case):
    """--soa_cache: the second run maps the first BAM's packed arrays back in instead of parsing (and refills the QNAME
    dictionary when a second BAM or --output_read_ids needs the names); same files as the reference either way."""
    cache = str(tmp_path / "cache")
    c, got1 = _run_cli(engine, case, tmp_path, extra=["--soa_cache", cache])
    assert not compare.diff_outputs(c["ref"], got1)
    assert any(os.path.isfile(os.path.join(cache, d, "done")) for d in os.listdir(cache))
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        c, got2 = _run_cli(engine, case, tmp_path, extra=["--soa_cache", cache])
    assert "from the SoA cache" in buf.getvalue()
    assert not compare.diff_outputs(c["ref"], got2)
    assert got1 == got2
