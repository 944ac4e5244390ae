"""Generate the golden fixtures by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Needs /root/reference (read-only) and the harness in oracle/harness.  For every case it writes the
inputs (VCF + SAM text named *.bam, as the harness serves them) and the reference's own outputs
under tests/golden/cases/<case>/ : ref.<file> for the six outputs (VCF decompressed) plus, for the
mapper-level cases, the reference mapper's TSV.  The reference was run with PYTHONHASHSEED=0.
"""
import gzip
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.harness import run_reference as rr       # noqa: E402
from phaser_b200 import synth                        # noqa: E402

CASES = os.path.join(HERE, "cases")

VCF_HEAD = ("##fileformat=VCFv4.2\n##contig=<ID=1,length=100000>\n##contig=<ID=2,length=100000>\n"
            '##INFO=<ID=AF,Number=A,Type=Float,Description="Allele frequency">\n'
            '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
            '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Depth">\n'
            "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tOTHER\tS1\n")
SAM_HEAD = "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:1\tLN:100000\n@SQ\tSN:2\tLN:100000\n"


def vline(chrom, pos, vid, ref, alt, gt, filt="PASS", af="0.2", fmt="GT:DP", other="0|0:5", extra=":7"):
    # note: the OTHER sample's 0|0 is cut away before the grep (phaser.py:220-224 cuts first)
    return "%s\t%d\t%s\t%s\t%s\t50\t%s\tAF=%s\t%s\t%s\t%s%s\n" % (chrom, pos, vid, ref, alt, filt, af, fmt, other, gt, extra)


def sline(q, flag, chrom, pos, mapq, cigar, seq, qual=None, AS=140, tlen=0, tags=True):
    qual = qual if qual is not None else "F" * len(seq)
    t = "\tNH:i:1\tAS:i:%d" % AS if tags and AS is not None else ("\tNH:i:1" if tags else "")
    return "%s\t%d\t%s\t%d\t%d\t%s\t=\t%d\t%d\t%s\t%s%s\n" % (q, flag, chrom, pos, mapq, cigar, pos, tlen, seq, qual, t)


def quirk_case():
    """Hand-written records, one per mapper quirk of SURVEY.md section 8a (Q3, Q4, Q5, Q27, clips, =/X,
    D followed by I, IUPAC D, multi-allelic sites, duplicate positions, '.' ids) + a few clean pairs so
    that phasing has something to do."""
    V = []
    V.append(vline("1", 105, "rs1", "A", "G", "0|1"))
    V.append(vline("1", 112, "rs2", "C", "T", "1|0"))
    V.append(vline("1", 120, ".", "G", "A", "0/1"))
    V.append(vline("1", 139, "rs4", "T", "C", "0|1"))
    V.append(vline("1", 144, "rs5", "A", "C", "0|1"))
    V.append(vline("1", 150, "rs6", "A", "C,G", "1|2"))          # multi-allelic, sample carries both ALTs
    V.append(vline("1", 150, "rs6b", "A", "T", "0|1"))           # same position again
    V.append(vline("1", 160, "rs7", "G", "T", "1|1"))            # hom: removed by the grep
    V.append(vline("1", 161, "rs8", "G", "T", "0|1", filt="q10"))  # not PASS
    V.append(vline("1", 162, "rs9", "GA", "G", "0|1"))           # indel: excluded
    V.append(vline("1", 170, "rs10", "C", "A", "0|1", fmt="GT", extra=""))
    V.append(vline("1", 171, "rs11", "C", "A", "./."))
    V.append(vline("1", 180, "rs:12", "T", "G", "0|1"))
    V.append(vline("1", 190, "rs13", "T", "G", "1/0"))
    V.append(vline("2", 50, "rs20", "A", "T", "0|1"))
    V.append(vline("2", 60, "rs21", "C", "G", "0|1"))
    V.append(vline("2", 300, "rs22", "C", "G", "0|1"))
    S = []

    def rd(name, flag, chrom, pos, cigar, edits=None, qual_edits=None, **kw):
        n = 0; num = ""
        for ch in cigar:
            if ch.isdigit():
                num += ch
            else:
                if ch in "MIS=X":
                    n += int(num)
                num = ""
        seq = ["A"] * n
        q = ["F"] * n
        for k, b in (edits or {}).items():
            seq[k] = b
        for k, b in (qual_edits or {}).items():
            q[k] = b
        S.append((chrom, pos, sline(name, flag, chrom, pos, 255, cigar, "".join(seq), "".join(q), **kw)))

    # contig 1
    rd("q3", 99, "1", 95, "2M3N40M1I10M", {41: "C", 42: "G", 46: "C"})      # Q3: insertion lands on 144 not 139
    rd("q4", 99, "1", 100, "6M2I22M", {5: "G", 6: "C", 7: "C", 14: "T"})     # Q4: G + CC at 105 -> other
    rd("del", 99, "1", 100, "10M4D20M", {5: "G"})                            # deletion over 112 -> nothing
    rd("lowq", 99, "1", 100, "30M", {5: "G", 12: "T"}, {5: "#"})             # low quality at 105 -> nothing
    rd("di", 99, "1", 100, "12M1D1I10M", {5: "A", 12: "T"})                  # D then I: key 12 = the D itself -> "T"
    rd("clipS", 99, "1", 102, "3S20M", {6: "G", 13: "T"})                    # soft clip: query offset shifts
    rd("clipH", 99, "1", 102, "2H20M3H", {3: "G", 10: "T"})
    rd("eqx", 99, "1", 100, "5=1X24=", {5: "G", 12: "C", 20: "A"})
    rd("iupacD", 99, "1", 100, "30M", {5: "D", 12: "T"})                     # IUPAC D is stripped like a deletion
    rd("nbase", 99, "1", 100, "30M", {5: "N", 12: "T"})
    rd("multi", 99, "1", 140, "30M", {4: "C", 10: "C"})                      # 144:C alt ; 150:C = allele 0 of 1|2, other for rs6b
    rd("multi2", 147, "1", 140, "30M", {4: "A", 10: "G"})
    rd("multi3", 99, "1", 140, "30M", {4: "A", 10: "T"})
    for i in range(6):                                                        # clean cis support 105-112-120
        rd("c%d" % i, 99, "1", 100, "30M", {5: "G", 12: "C", 20: "A"} if i % 2 else {5: "A", 12: "T", 20: "G"})
    rd("m1", 99, "1", 100, "15M", {5: "G", 12: "C"})                          # Q27: mates disagree at 112
    rd("m1", 147, "1", 106, "20M", {6: "T", 14: "A"})
    rd("far", 99, "1", 165, "30M", {5: "A", 15: "G", 25: "G"}, tlen=900)      # 170 alt, 180 alt, 190 alt
    rd("far2", 99, "1", 165, "30M", {5: "C", 15: "T", 25: "T"}, tlen=-100)
    rd("far3", 99, "1", 165, "30M", {5: "A", 15: "T", 25: "G"})               # conflicting with far/far2 on 170-180
    rd("far4", 99, "1", 165, "30M", {5: "A", 15: "G", 25: "G"})
    rd("dup", 99 | 0x400, "1", 100, "30M", {5: "G"})                          # duplicate: filtered
    rd("unpaired", 65, "1", 100, "30M", {5: "G"})                             # not proper pair: filtered
    # contig 2: spliced pair joining 50/60 with 300
    rd("s1", 99, "2", 45, "20M230N20M", {5: "T", 15: "G", 25: "G"})
    rd("s2", 99, "2", 45, "20M230N20M", {5: "A", 15: "C", 25: "C"})
    rd("s3", 99, "2", 45, "20M230N20M", {5: "T", 15: "G", 25: "G"})
    rd("s4", 99, "2", 55, "10M230N20M", {5: "C", 15: "C"})
    S.sort(key=lambda t: (t[0], t[1]))
    return VCF_HEAD + "".join(V), SAM_HEAD + "".join(x[2] for x in S)


def indel_case():
    """--include_indels 1 (phaser.py:1398-1408, read_variant_map.py:236-258 with ref_length > 1): deletions, insertions,
    an MNP, a multi-allelic indel site, reads carrying each allele through D / I CIGAR ops, a deletion that ends at a
    segment edge, a site cut by a splice junction, low-quality and N bases inside a multi-base call."""
    V = []
    V.append(vline("1", 105, "rs1", "A", "G", "0|1"))               # plain SNVs around the indels
    V.append(vline("1", 110, "del2", "CAG", "C", "0|1"))            # 2-base deletion
    V.append(vline("1", 118, "ins3", "T", "TACG", "1|0"))           # 3-base insertion
    V.append(vline("1", 125, "mnp", "AC", "GT", "0|1"))             # MNP: ref_len 2, no length change
    V.append(vline("1", 131, "mai", "A", "AT,ATT", "1|2"))          # multi-allelic insertion, sample carries both ALTs
    V.append(vline("1", 136, "rs6", "C", "T", "0/1"))
    V.append(vline("1", 140, "del1", "GA", "G", "0|1"))
    V.append(vline("1", 146, "cpx", "ATG", "AC", "1|0"))            # complex: ref 3 -> 2
    V.append(vline("1", 160, "rs9", "T", "C", "0|1"))
    V.append(vline("2", 58, "edge", "AAAA", "A", "0|1"))            # REF runs over the end of short reads / the first exon
    V.append(vline("2", 300, "rs21", "C", "G", "0|1"))
    V.append(vline("2", 305, "ins2", "A", "AGG", "0|1"))
    S = []

    def rd(name, flag, chrom, pos, cigar, seq, qual=None, **kw):
        S.append((chrom, pos, sline(name, flag, chrom, pos, 255, cigar, seq, qual, **kw)))

    def ref1(lo, n):                       # reference sequence of contig 1 around the sites (A everywhere else)
        g = {105: "A", 110: "C", 111: "A", 112: "G", 118: "T", 125: "A", 126: "C", 131: "A", 136: "C", 140: "G", 141: "A",
             146: "A", 147: "T", 148: "G", 160: "T"}
        return "".join(g.get(p, "A") for p in range(lo, lo + n))

    def edit(seq, lo, changes):
        s = list(seq)
        for p, b in changes.items():
            s[p - lo] = b
        return "".join(s)

    # haplotype R: reference everywhere.  60 bases from 100
    for i in range(4):
        rd("refhap%d" % i, 99, "1", 100, "62M", ref1(100, 62))
    # haplotype X: every non-reference allele: SNV G@105, del CAG>C (2D after 110), ins ACG after 118, MNP GT@125-126,
    # ATT after 131 (allele index 2), T@136, del GA>G (1D after 140), ATG>AC: 146 A, 147 C, 148 deleted, C@160
    def alt_read(name, flag, second_ins="TT", AS=140):
        left = edit(ref1(100, 11), 100, {105: "G"})                       # 100..110
        mid1 = ref1(113, 6)                                               # 113..118
        mid2 = edit(ref1(119, 13), 119, {125: "G", 126: "T"})             # 119..131
        mid3 = edit(ref1(132, 9), 132, {136: "T"})                        # 132..140
        mid4 = edit(ref1(142, 6), 142, {147: "C"})                        # 142..147
        tail = edit(ref1(149, 13), 149, {160: "C"})                       # 149..161
        seq = left + mid1 + "ACG" + mid2 + second_ins + mid3 + mid4 + tail
        cig = "11M2D6M3I13M%dI9M1D6M1D13M" % len(second_ins)
        rd(name, flag, "1", 100, cig, seq, AS=AS)
    for i in range(4):
        alt_read("althap%d" % i, 99)
    alt_read("althapT", 99, second_ins="T")                                # allele index 1 of the multi-allelic site
    alt_read("althapTTT", 99, second_ins="TTT")                            # neither allele -> other
    # partial / odd carriers
    rd("mnp_half", 99, "1", 120, "20M", edit(ref1(120, 20), 120, {125: "G"}))                  # GC: matches neither -> other
    rd("mnp_lowq", 99, "1", 120, "20M", edit(ref1(120, 20), 120, {125: "G", 126: "T"}), "FFFFF#FFFFFFFFFFFFFF")   # NT -> other
    rd("del_n", 99, "1", 104, "20M", edit(ref1(104, 20), 104, {111: "N"}))                     # CNG -> other
    rd("short_end", 99, "1", 100, "12M", ref1(100, 12))                                        # 110 + 3 > read end: nothing for del2
    rd("short_exact", 99, "1", 100, "13M", ref1(100, 13))                                      # ends exactly at 112: CAG
    rd("del_long", 99, "1", 100, "9M6D30M", ref1(100, 9) + ref1(115, 30))                      # deletion swallows the whole del2 REF -> ""
    rd("del_partial", 99, "1", 100, "11M1D30M", ref1(100, 11) + ref1(112, 30))                 # CAG with A deleted -> CG -> other
    rd("ins_wrong", 99, "1", 110, "9M2I20M", ref1(110, 9) + "GG" + ref1(119, 20))              # TGG -> other
    rd("softclip", 99, "1", 108, "4S30M", "TTTT" + ref1(108, 30))
    # contig 2: exon 1 = 45..60, intron, exon 2 = 291..320 ; "edge" REF AAAA at 58..61 crosses the junction
    for i in range(3):
        rd("sp%d" % i, 99, "2", 45, "16M230N30M", "A" * 16 + edit("A" * 30, 291, {300: "C"}))
        rd("spalt%d" % i, 99, "2", 45, "16M230N15M2I15M", "A" * 16 + edit("A" * 15, 291, {300: "G"}) + "GG" + "A" * 15)
    rd("edge_ok", 99, "2", 50, "20M", "A" * 20)                                                # AAAA inside one segment
    rd("edge_del", 99, "2", 50, "9M3D20M", "A" * 29)                                           # A + 3 deleted -> "A" = alt
    rd("q3ins", 99, "2", 45, "16M230N10M2I20M", "A" * 16 + "A" * 9 + "C" + "GG" + "A" * 20)    # Q3: insertion keyed whole-read offset
    S.sort(key=lambda t: (t[0], t[1]))
    return VCF_HEAD + "".join(V), SAM_HEAD + "".join(x[2] for x in S)


def fuzz_indel_case(seed=77, n_sites=140, n_reads=900):
    """Seeded random sites (REF / ALT strings of 1-4 bases, multi-allelic, overlapping REF spans) under reads drawn from a
    random reference with random CIGARs mixing M = X I D N S H P; reads carry one haplotype's SNV alleles, pure
    insertions / deletions are written into the CIGAR, and a few percent of the bases are errors, N or low quality:
    every branch of split_read / identify_allele with ref_length >= 1."""
    import random
    rnd = random.Random(seed)
    G = [rnd.choice("ACGT") for _ in range(2000)]           # G[p] = reference base at 1-based position p
    V = []; sites = []
    # distinct positions: two sites at one position inside a multi-site block are ordered by CPython's set iteration in
    # the reference, which no canonical order can reproduce; the quirks case keeps such a pair
    pos = sorted(rnd.sample(range(100, 1600), n_sites))
    for i, p in enumerate(pos):
        rl = rnd.choice([1, 1, 1, 1, 2, 3, 4])
        ref = "".join(G[p:p + rl])
        alts = []
        while len(alts) < rnd.choice([1, 1, 1, 1, 2]):
            k = rnd.random()
            if k < 0.5:
                a = rnd.choice("ACGT") if rl == 1 else "".join(rnd.choice("ACGT") for _ in range(rl))
            elif k < 0.75:
                a = ref[0]                                                      # pure deletion (or a no-op for SNVs)
            else:
                a = ref + "".join(rnd.choice("ACGT") for _ in range(rnd.randrange(1, 4))) if rl == 1 else ref[0] + rnd.choice("ACGT")
            if a != ref and a not in alts:
                alts.append(a)
        gt = rnd.choice(["0|1", "1|0", "0/1"]) if len(alts) == 1 else rnd.choice(["1|2", "0|2", "2|1", "0/1"])
        idx = [int(c) for c in gt if c.isdigit()]
        alls = [ref] + alts
        V.append(vline("1", p, "rs%d" % i if rnd.random() < 0.9 else ".", ref, ",".join(alts), gt))
        sites.append((p, rl, [alls[idx[0]], alls[idx[1]]], [rnd.randrange(2), ]))
    by_pos = {}
    for st in sites:
        by_pos.setdefault(st[0], st)
    S = []
    for i in range(n_reads):
        hap = rnd.randrange(2)
        start = rnd.randrange(60, 1600)
        ops = []; seq = []
        if rnd.random() < 0.1:
            ops.append((rnd.randrange(1, 5), "H"))
        if rnd.random() < 0.15:
            n = rnd.randrange(1, 6); ops.append((n, "S")); seq += [rnd.choice("ACGT") for _ in range(n)]
        g = start
        n_blocks = rnd.choice([1, 1, 2, 2, 3, 4])
        for b in range(n_blocks):
            n = rnd.randrange(8, 60)
            kind = rnd.choice("MMMM=X")
            x = g; run = 0
            while x < g + n:
                st = by_pos.get(x)
                done = False
                if st is not None and x + st[1] <= g + n and x > g:
                    al = st[2][hap ^ st[3][0]]
                    ref = "".join(G[x:x + st[1]])
                    if len(al) == st[1]:                                         # SNV / MNP: substitute
                        seq += list(al); run += st[1]; x += st[1]; done = True
                    elif len(al) == 1 and al == ref[0] and st[1] > 1:            # deletion
                        seq.append(al); run += 1
                        ops.append((run, kind)); ops.append((st[1] - 1, "D")); run = 0
                        x += st[1]; done = True
                    elif st[1] == 1 and al.startswith(ref) and x + 1 < g + n:    # insertion
                        seq.append(ref); run += 1
                        ops.append((run, kind)); ops.append((len(al) - 1, "I")); seq += list(al[1:]); run = 0
                        x += 1; done = True
                if not done:
                    seq.append(G[x]); run += 1; x += 1
            if run:
                ops.append((run, kind))
            g += n
            if b + 1 < n_blocks:
                k = rnd.random()
                if k < 0.15:
                    m = rnd.randrange(1, 4); ops.append((m, "I")); seq += [rnd.choice("ACGT") for _ in range(m)]
                elif k < 0.3:
                    m = rnd.randrange(1, 6); ops.append((m, "D")); g += m
                elif k < 0.85:
                    m = rnd.randrange(5, 120); ops.append((m, "N")); g += m
                elif k < 0.93:
                    m = rnd.randrange(1, 3); ops.append((m, "D")); g += m
                    m = rnd.randrange(1, 3); ops.append((m, "I")); seq += [rnd.choice("ACGT") for _ in range(m)]
                else:
                    ops.append((1, "P"))
        if rnd.random() < 0.15:
            n = rnd.randrange(1, 6); ops.append((n, "S")); seq += [rnd.choice("ACGT") for _ in range(n)]
        if rnd.random() < 0.1:
            ops.append((rnd.randrange(1, 5), "H"))
        merged = []
        for n, o in ops:                                                         # M5 M3 -> M8 (same op twice in a row)
            if merged and merged[-1][1] == o:
                merged[-1] = (merged[-1][0] + n, o)
            else:
                merged.append((n, o))
        seq = [(c if rnd.random() > 0.01 else rnd.choice("ACGTN")) for c in seq]
        qual = "".join("F" if rnd.random() > 0.05 else "#" for _ in seq)
        S.append((start, sline("f%d" % (i // 2), 99 if i % 2 == 0 else 147, "1", start, 255, "".join("%d%s" % x for x in merged),
                               "".join(seq), qual, AS=rnd.randrange(100, 150), tlen=rnd.choice([0, 250, -300, 800]))))
    S.sort(key=lambda t: t[0])
    return VCF_HEAD + "".join(V), SAM_HEAD + "".join(x[1] for x in S)


def run_case(name, vcf_text, sams, args, mapq="255", paired_end="1"):
    d = os.path.join(CASES, name)
    if os.path.isdir(d):
        shutil.rmtree(d)
    os.makedirs(d)
    vcf = os.path.join(d, "in.vcf.gz")
    with gzip.open(vcf, "wt") as f:
        f.write(vcf_text)
    paths = []
    for bn, text in sams:
        p = os.path.join(d, bn)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(text)
        paths.append(p)
    tmp = tempfile.mkdtemp()
    r = rr.run_reference(vcf, paths, os.path.join(tmp, "ref"), "S1", mapq=mapq, paired_end=paired_end,
                         extra_args=args, hashseed=0)
    if r["returncode"] != 0:
        raise RuntimeError(r["log"][-2000:])
    for suf in rr.OUTPUT_SUFFIXES:
        text = rr.read_text(r[suf])
        with open(os.path.join(d, "ref." + suf.replace(".gz", "")), "w") as f:
            f.write(text)
    if "--output_network" in args:                       # phaser.py:1128-1157
        for suf in ("network.links.txt", "network.nodes.txt"):
            shutil.copy(os.path.join(tmp, "ref." + suf), os.path.join(d, "ref." + suf))
    for p in paths:
        os.remove(p + ".bai")
    os.remove(vcf + ".tbi")
    with open(os.path.join(d, "case.json"), "w") as f:
        json.dump(dict(bams=[bn for bn, _ in sams], args=[str(a) for a in args], mapq=mapq, paired_end=paired_end,
                       sample="S1"), f, indent=1)
    # mapper-level golden: the reference mapper on each BAM with the table the product's VCF parser yields
    from phaser_b200 import vcfio
    col = vcfio.sample_column_map(vcf)["S1"]
    vt, _ = vcfio.parse_vcf(vcf, col, include_indels=int(args[args.index("--include_indels") + 1]) if "--include_indels" in args else 0)
    table = os.path.join(tmp, "table.tsv")
    with open(table, "w") as f:
        for v in range(vt.n_variants):
            c = [c for c in range(len(vt.contigs)) if vt.contig_var_off[c] <= v < vt.contig_var_off[c + 1]][0]
            f.write("\t".join([vt.contigs[c], str(int(vt.pos[v])), vt.ids[v], vt.rsids[v], ",".join(vt.all_alleles[v]),
                               str(int(vt.ref_len[v])), vt.gt[v], vt.maf[v]]) + "\n")
    for p in paths:
        # the mapper is fed what the two samtools stages would pass: here everything on the VCF's contigs
        out = os.path.join(d, "ref.mapper." + os.path.relpath(p, d).replace(os.sep, "_") + ".tsv")
        filt = os.path.join(tmp, "filt.sam")
        with open(p) as fin, open(filt, "w") as fo:
            for ln in fin:
                if ln[0] == "@":
                    fo.write(ln); continue
                c = ln.split("\t")
                fl = int(c[1])
                if (fl & 0x400) or (paired_end == "1" and not fl & 2) or int(c[4]) < int(mapq.split(",")[0]):
                    continue
                fo.write(ln)
        rr.run_mapper(filt, table, out, baseq=10, isize_cutoff=0)
    shutil.rmtree(tmp)
    print("golden case", name, "ok")


def synth_case(name, seed, n_variants, n_pairs, n_bams, args, qname=None, bam_files=None, phased_frac=0.9, **read_kw):
    """`qname`: QNAME prefix of every BAM (default: one prefix per BAM, so no read name is shared between BAMs);
    `bam_files`: file names inside the case directory (default b0.bam, b1.bam, ...)."""
    g = synth.make_genome(seed, n_variants, contigs=[("21", 120000), ("22", 80000)], n_genes=max(2, n_variants // 8),
                          phased_frac=phased_frac)
    tmp = tempfile.mkdtemp()
    vcf = synth.write_vcf(g, os.path.join(tmp, "x.vcf.gz"))
    sams = []
    for b in range(n_bams):
        rec = synth.make_reads(g, seed * 100 + b, n_pairs, dup_frac=0.05, **read_kw)
        p = synth.write_sam(rec, g, os.path.join(tmp, "b%d.bam" % b), bam_name=qname if qname else "b%d" % b)
        sams.append((bam_files[b] if bam_files else "b%d.bam" % b, open(p).read()))
    with gzip.open(vcf, "rt") as f:
        vt = f.read()
    shutil.rmtree(tmp)
    run_case(name, vt, sams, args)


def option_case(name, base, args, mapq="255", paired_end="1"):
    """Same inputs as case `base` (not stored twice), other command-line options."""
    b = os.path.join(CASES, base)
    meta = json.load(open(os.path.join(b, "case.json")))
    vcf_text = gzip.open(os.path.join(b, "in.vcf.gz"), "rt").read()
    sams = [(bn, open(os.path.join(b, bn)).read()) for bn in meta["bams"]]
    run_case(name, vcf_text, sams, args, mapq=mapq, paired_end=paired_end)
    d = os.path.join(CASES, name)
    for bn in meta["bams"]:
        os.remove(os.path.join(d, bn))
    os.remove(os.path.join(d, "in.vcf.gz"))
    m = json.load(open(os.path.join(d, "case.json")))
    m["inputs"] = base
    json.dump(m, open(os.path.join(d, "case.json"), "w"), indent=1)


def blacklist_cases():
    """--blacklist / --haplo_count_blacklist (phaser.py:218-243): BED files derived from the case's own variants."""
    b = os.path.join(CASES, "rna_two_bams")
    lines = [l for l in gzip.open(os.path.join(b, "in.vcf.gz"), "rt") if not l.startswith("#")]
    pos = [(l.split("\t")[0], int(l.split("\t")[1])) for l in lines]
    os.makedirs(os.path.join(HERE, "beds"), exist_ok=True)
    bl = os.path.join(HERE, "beds", "blacklist.bed"); hbl = os.path.join(HERE, "beds", "haplo_blacklist.bed")
    with open(bl, "w") as f:          # drops every 9th variant from phasing altogether
        for c, p_ in pos[4::9]:
            f.write("%s\t%d\t%d\n" % (c, p_ - 1, p_))
    with open(hbl, "w") as f:         # a few wider intervals: blocks lose members from the haplotypic counts
        for c, p_ in pos[2::7]:
            f.write("%s\t%d\t%d\n" % (c, max(0, p_ - 40), p_ + 40))
    option_case("opt_blacklists", "rna_two_bams", ["--blacklist", bl, "--haplo_count_blacklist", hbl])
    m = json.load(open(os.path.join(CASES, "opt_blacklists", "case.json")))
    m["args"] = ["--blacklist", "beds/blacklist.bed", "--haplo_count_blacklist", "beds/haplo_blacklist.bed"]
    json.dump(m, open(os.path.join(CASES, "opt_blacklists", "case.json"), "w"), indent=1)


def tie_case():
    """Q16 (phaser.py:708-726): an edge whose cis and trans support tie survives the binomial test, stays in the variant
    graph -- so it glues its variants into one component -- but adds no allele links.  Contig 1: rs1-rs2 are joined by 40
    clean reads; rs3 hangs on rs2 by exactly one cis and one trans read.  The reference therefore sees ONE component of
    three variants, cannot resolve it, enumerates 2^3 configurations, finds two equally supported ones and drops the whole
    block: rs1 and rs2 end up unphased.  (Without the glue rs1-rs2 would be an ordinary phased block.)  One read with a
    third base at rs1 gives the noise estimate the non-zero value without which the tie edge would fail the test.
    Contig 2: the same picture on a longer chain, a tie in the middle of two clean pairs."""
    V = [vline("1", 105, "rs1", "A", "G", "0|1"), vline("1", 112, "rs2", "C", "T", "0|1"), vline("1", 128, "rs3", "G", "A", "0|1"),
         vline("2", 105, "rs4", "A", "G", "0|1"), vline("2", 112, "rs5", "C", "T", "1|0"), vline("2", 128, "rs6", "G", "A", "0|1"),
         vline("2", 135, "rs7", "T", "C", "0/1")]
    S = []

    def rd(name, chrom, pos, n, edits):
        seq = ["A"] * n
        for k, b in edits.items():
            seq[k] = b
        S.append((chrom, pos, sline(name, 99, chrom, pos, 255, "%dM" % n, "".join(seq))))

    for chrom in ("1", "2"):
        for i in range(40):          # rs1/rs4 - rs2/rs5, clean cis, both haplotypes
            rd("c%s_%d" % (chrom, i), chrom, 100, 20, {5: "A", 12: "C"} if i % 2 else {5: "G", 12: "T"})
        rd("x%s" % chrom, chrom, 100, 20, {5: "C", 12: "C"})                      # third base at the first site -> 'other'
        rd("t%s_cis" % chrom, chrom, 108, 26, {4: "C", 20: "G"})                  # second - third site: one cis ...
        rd("t%s_trans" % chrom, chrom, 108, 26, {4: "C", 20: "A"})               # ... one trans read
    for i in range(6):               # contig 2: rs6 - rs7 clean cis on the far side of the tie
        rd("d%d" % i, "2", 120, 20, {8: "G", 15: "T"} if i % 2 else {8: "A", 15: "C"})
    S.sort(key=lambda t: (t[0], t[1]))
    return VCF_HEAD + "".join(V), SAM_HEAD + "".join(x[2] for x in S)


def engine_quirk_cases():
    """Reference runs that pin the engine-side quirks of SURVEY.md section 8a which no earlier case provably reaches
    (tests/test_quirk_coverage.py shows, per quirk, that a port WITHOUT the quirk no longer reproduces these files)."""
    # Q9 (phaser.py:578): both BAMs name their reads r.<i>, so hundreds of QNAMEs occur in both files, at unrelated
    # loci; read_vars of a shared name is overwritten by the later BAM
    synth_case("q9_shared_qnames", 300, 160, 700, 2, [], qname="r")
    # Q22 (phaser.py:935-939) / Q21: most genotypes are unphased, so many blocks have no known phase at all
    synth_case("q22_unphased_blocks", 301, 160, 900, 1, [], phased_frac=0.3)
    # Q26 (phaser.py:469-480): two BAMs with the same basename in different directories -> display names x.1, x.2
    synth_case("q26_same_basename", 302, 120, 500, 2, ["--haplo_count_bam_exclude", "1"], bam_files=["a/x.bam", "b/x.bam"])
    # Q16: tie edges glue components
    v, s = tie_case()
    run_case("q16_tie_glue", v, [("tie.bam", s)], ["--as_q_cutoff", "0"])
    # --chr (phaser.py:205-207, 1679-1680): het sites and output VCF lines of one contig only
    option_case("opt_chr", "rna_two_bams", ["--chr", "22", "--haplo_count_bam_exclude", "2"])


def config1_inputs(tmp):
    """BASELINE.json configs[0] shape: one contig (chr22 length), ~10k het SNVs, 300k read pairs.  The inputs are
    NOT stored (180 MB of SAM text): they are regenerated from the seed (torch CPU generator) and checked by hash."""
    import hashlib
    from phaser_b200 import engine as eng
    g = synth.make_genome(1000, 10000, contigs=[("22", 50818468)], n_genes=1250)
    vcf = synth.write_vcf(g, os.path.join(tmp, "c1.vcf.gz"))
    rec = synth.make_reads(g, 100000, 300000, dup_frac=0.05)
    sam = eng.write_sam_native(rec, g.contigs, os.path.join(tmp, "c1.bam"), "c1")     # host-only entry point of _phz.so
    h = hashlib.sha256()
    h.update(gzip.open(vcf, "rb").read()); h.update(open(sam, "rb").read())
    return vcf, sam, h.hexdigest()


def config1_case():
    tmp = tempfile.mkdtemp()
    vcf, sam, digest = config1_inputs(tmp)
    r = rr.run_reference(vcf, [sam], os.path.join(tmp, "ref"), "S1", hashseed=0)
    if r["returncode"] != 0:
        raise RuntimeError(r["log"][-2000:])
    d = os.path.join(HERE, "config1")
    os.makedirs(d, exist_ok=True)
    for suf in rr.OUTPUT_SUFFIXES:
        with gzip.open(os.path.join(d, "ref." + suf.replace(".gz", "") + ".gz"), "wt") as f:
            f.write(rr.read_text(r[suf]))
    json.dump({"sha256_inputs": digest, "note": "reference outputs for the seeded configs[0]-shape sample; inputs regenerated by "
               "make_golden.config1_inputs"}, open(os.path.join(d, "case.json"), "w"), indent=1)
    shutil.rmtree(tmp)
    print("golden config1 ok", digest[:12])


def gene_ae_cases():
    """phaser_gene_ae.py (UNMODIFIED, with the stub intervaltree of oracle/harness/stub) on the reference's own
    haplotypic_counts.txt of existing cases; features = seeded random, overlapping and nested intervals."""
    import random
    import subprocess
    GA = os.path.join(HERE, "gene_ae")
    script = os.path.join(os.path.dirname(rr.REFERENCE_DIR), "phaser_gene_ae", "phaser_gene_ae.py")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oracle", "harness", "stub"), PYTHONHASHSEED="0")
    for name, base, seed, args in (("two_bams_default", "rna_two_bams", 1, []),
                                   ("two_bams_maf_cov", "rna_two_bams", 2, ["--min_haplo_maf", "0.2", "--gw_cutoff", "0.6", "--min_cov", "3"]),
                                   ("conflict_default", "rna_conflict", 3, []),
                                   ("blacklists_cutoff1", "opt_blacklists", 4, ["--gw_cutoff", "1.0"])):
        # (a file written with --output_read_ids 1 crashes the reference script: its rows and header disagree, phaser.py:837 vs 1120)
        d = os.path.join(GA, name)
        os.makedirs(d, exist_ok=True)
        rnd = random.Random(seed)
        hc = os.path.join(CASES, base, "ref.haplotypic_counts.txt")
        feats = []
        for c, L in (("21", 120000), ("22", 80000), ("7", 50000)):
            for _ in range(60):
                a = rnd.randrange(0, L - 10); ln = rnd.choice([1, 40, 300, 2500, 9000, 40000])
                feats.append((c, a, min(L, a + ln)))
        rnd.shuffle(feats)
        bed = os.path.join(d, "features.bed")
        with open(bed, "w") as f:
            for i, (c, a, b) in enumerate(feats):
                f.write("%s\t%d\t%d\tgene%d\n" % (c, a, b, i))
        out = os.path.join(d, "ref.gene_ae.txt")
        r = subprocess.run([sys.executable, script, "--haplotypic_counts", hc, "--features", bed, "--o", out] + args,
                           env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stdout[-1000:] + r.stderr[-2000:])
        json.dump(dict(base=base, args=args), open(os.path.join(d, "case.json"), "w"), indent=1)
        print("golden gene_ae", name, "ok")


def main():
    os.makedirs(CASES, exist_ok=True)
    v, s = quirk_case()
    run_case("quirks", v, [("quirks.bam", s)], ["--as_q_cutoff", "0"])
    run_case("quirks_isize", v, [("quirks.bam", s)], ["--as_q_cutoff", "0", "--isize", "500"])
    synth_case("rna_small", 21, 160, 900, 1, [])
    synth_case("rna_two_bams", 22, 160, 700, 2, ["--haplo_count_bam_exclude", "2"])
    synth_case("rna_conflict", 23, 120, 1500, 1, ["--max_block_size", "4"], switch_per_base=0.03)
    option_case("opt_maf_gwvcf2", "rna_conflict", ["--gw_phase_method", "1", "--gw_phase_vcf", "2", "--max_block_size", "4"])
    option_case("opt_gwvcf1", "rna_small", ["--gw_phase_vcf", "1", "--gw_phase_vcf_min_confidence", "0.6"])
    option_case("opt_nounphased_uid", "rna_two_bams", ["--unphased_vars", "0", "--unique_ids", "1", "--id_separator", "-"])
    option_case("opt_filters", "rna_small", ["--pass_only", "0", "--remove_dups", "0", "--cc_threshold", "0.05", "--as_q_cutoff", "0.2"],
                paired_end="0")
    option_case("opt_quirks_baseq", "quirks", ["--as_q_cutoff", "0", "--max_block_size", "3"], mapq="0")
    blacklist_cases()
    option_case("opt_read_ids", "rna_two_bams", ["--output_read_ids", "1"])
    option_case("opt_network", "rna_two_bams", ["--output_network", "21_24913_A_C"])
    v, s = indel_case()
    run_case("indels", v, [("indels.bam", s)], ["--include_indels", "1", "--as_q_cutoff", "0"])
    vf, sf = fuzz_indel_case()
    run_case("fuzz_indels", vf, [("fuzz.bam", sf)], ["--include_indels", "1", "--as_q_cutoff", "0", "--isize", "500"])
    option_case("fuzz_snvs", "fuzz_indels", ["--as_q_cutoff", "0.1", "--max_block_size", "5"])
    vq, sq = quirk_case()
    run_case("opt_quirks_indels", vq, [("quirks.bam", sq)], ["--include_indels", "1", "--as_q_cutoff", "0"])
    engine_quirk_cases()
    config1_case()
    gene_ae_cases()


if __name__ == "__main__":
    if len(sys.argv) > 1:          # python make_golden.py engine_quirk_cases  -> only that group
        for fn in sys.argv[1:]:
            globals()[fn]()
    else:
        main()
