"""SAM-text ingest -> ReadBatch (host).  Replaces the two `samtools view` stages of the reference
pipeline (phaser/phaser.py:1346) and the per-line field parsing of read_variant_map.py:25-64.

Filter semantics kept (phaser.py:505-513, 1346): contig must be one the VCF names, drop duplicates
(0x400) iff --remove_dups 1, require proper pair (0x2) iff paired_end, MAPQ >= per-BAM mapq.
Secondary / supplementary / QC-fail records are NOT filtered (the reference does not either).
The `-L bed` stage is only a pre-filter (records overlapping no het site emit nothing) and is not
reproduced.  Records keep file order inside a contig; contigs are laid out in VCF order, the order
in which the reference maps them.
"""
import gzip
from typing import Dict, List

import numpy as np

from .layout import ReadBatch, BASE_CODE, CODE_N, CIGAR_CODE, AS_MISSING


class FragmentDictionary:
    """QNAME -> dense fragment id, shared by all BAMs of a run (mates and same-named reads of
    different BAMs are ONE read to the reference: its sets hold QNAME strings, phaser.py:1305-1324)."""

    def __init__(self):
        self.ids: Dict[str, int] = {}
        self.names: List[str] = []

    def get(self, name: str) -> int:
        i = self.ids.get(name)
        if i is None:
            i = len(self.names)
            self.ids[name] = i
            self.names.append(name)
        return i


def _open_text(path):
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic == b"\x1f\x8b":
        return gzip.open(path, "rt")
    return open(path, "r")


def parse_sam(path, contigs: List[str], fragdict: FragmentDictionary, remove_dups=True, proper_pair=True,
              min_mapq=0) -> ReadBatch:
    with _open_text(path) as f:
        return parse_sam_stream(f, contigs, fragdict, remove_dups, proper_pair, min_mapq)


def parse_sam_stream(f, contigs: List[str], fragdict: FragmentDictionary, remove_dups=True, proper_pair=True,
                     min_mapq=0) -> ReadBatch:
    cidx = {c: i for i, c in enumerate(contigs)}
    nc = len(contigs)
    per = [[] for _ in range(nc)]
    if True:
        for line in f:
            if line[0] == "@":
                continue
            c = line.rstrip().split("\t")
            ci = cidx.get(c[2])
            if ci is None:
                continue
            flag = int(c[1])
            if remove_dups and (flag & 0x400):
                continue
            if proper_pair and not (flag & 2):
                continue
            if int(c[4]) < min_mapq:
                continue
            a = AS_MISSING
            for t in c[11:]:
                if t.startswith("AS:"):
                    a = int(t.split(":")[2])
                    break
            per[ci].append((c[0], int(c[3]), int(c[8]), a, c[5], c[9], c[10]))
    recs = [r for p in per for r in p]
    R = len(recs)
    off = np.zeros(nc + 1, np.int64)
    off[1:] = np.cumsum([len(p) for p in per])
    pos = np.empty(R, np.int32); tlen = np.empty(R, np.int32); aln = np.empty(R, np.int16)
    frag = np.empty(R, np.uint32)
    cig_off = np.zeros(R + 1, np.uint32); seq_off = np.zeros(R + 1, np.uint64)
    cig: List[int] = []
    bases = bytearray(); quals = bytearray()
    for i, (qn, p, tl, a, cg, sq, ql) in enumerate(recs):
        pos[i] = p; tlen[i] = tl
        if a != AS_MISSING and not (-32767 <= a <= 32767):
            raise ValueError("AS:i value %d outside the int16 range of the packed layout" % a)
        aln[i] = a
        frag[i] = fragdict.get(qn)
        n = 0
        if cg != "*":
            for ch in cg:
                o = ord(ch)
                if 48 <= o <= 57:
                    n = n * 10 + o - 48
                else:
                    cig.append((n << 4) | CIGAR_CODE[ch])
                    n = 0
        cig_off[i + 1] = len(cig)
        if sq == "*":
            sq = ""
        if ql == "*" or len(ql) != len(sq):
            raise ValueError("record %s: QUAL missing or not the length of SEQ (unsupported)" % qn)
        bases.extend(BASE_CODE.get(ch, CODE_N) for ch in sq)
        quals.extend(ql.encode())
        seq_off[i + 1] = len(bases)
    b = np.frombuffer(bytes(bases), np.uint8)
    if b.shape[0] & 1:
        b = np.concatenate([b, np.zeros(1, np.uint8)])
    seq = ((b[0::2] << 4) | b[1::2]).astype(np.uint8)
    qual = (np.frombuffer(bytes(quals), np.uint8).astype(np.int16) - 33).clip(0, 255).astype(np.uint8)
    return ReadBatch(nc, off, pos, tlen, aln, frag, cig_off, np.asarray(cig, np.uint32), seq_off, seq, qual,
                     fragdict.names)


def read_alignments(path, contigs, fragdict, remove_dups=True, proper_pair=True, min_mapq=0) -> ReadBatch:
    """BAM (BGZF, magic "BAM\\1" after inflate) or SAM text (optionally gzipped), by content."""
    with open(path, "rb") as f:
        magic = f.read(18)
    if magic[:2] == b"\x1f\x8b" and len(magic) >= 14 and magic[3] & 4 and magic[12:14] == b"BC":
        from . import bamio
        rb = bamio.read_bam(path, contigs, fragdict, remove_dups, proper_pair, min_mapq)
    else:
        rb = parse_sam(path, contigs, fragdict, remove_dups, proper_pair, min_mapq)
    check_sorted(rb, path)
    return rb


def check_sorted(rb: ReadBatch, path=""):
    """The kernels (like the reference's streaming mapper) need coordinate-sorted records."""
    for c in range(rb.n_contigs):
        p = rb.pos[int(rb.contig_rec_off[c]):int(rb.contig_rec_off[c + 1])]
        if p.shape[0] > 1 and np.any(np.diff(p.astype(np.int64)) < 0):
            raise ValueError("%s: records are not sorted by coordinate" % path)
