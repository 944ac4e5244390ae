"""BAM ingest / writer (host).  Replaces `samtools view` of the reference pipeline (phaser.py:1346).

Reader: BGZF inflate (zlib) + BAM record decode straight into the packed ReadBatch layout -- BAM's
own encodings (4-bit bases, len<<4|op CIGAR words, phred bytes) ARE the device layout, so nothing is
converted to text.  Filters as in samio.parse_sam.  Writer: used by tests and the synthetic generator.
Specification: SAM/BAM format v1.6, section 4.2.
"""
import struct
from typing import List

import numpy as np

from . import bgzf
from .layout import ReadBatch, AS_MISSING, BASE_CODE, CODE_N, CIGAR_CODE

_TAG_SIZE = {b"A": 1, b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4}
_TAG_FMT = {b"c": "<b", b"C": "<B", b"s": "<h", b"S": "<H", b"i": "<i", b"I": "<I"}


def _find_as(buf, p, end):
    """First AS tag (integer types) in the aux block, AS_MISSING if absent (read_variant_map.py:53-64)."""
    while p + 3 <= end:
        tag = buf[p:p + 2]; t = buf[p + 2:p + 3]; p += 3
        if t in _TAG_SIZE:
            n = _TAG_SIZE[t]
            if tag == b"AS" and t in _TAG_FMT:
                return struct.unpack_from(_TAG_FMT[t], buf, p)[0]
            p += n
        elif t in (b"Z", b"H"):
            q = buf.index(b"\0", p)
            if tag == b"AS":
                try:
                    return int(buf[p:q])
                except ValueError:
                    return AS_MISSING
            p = q + 1
        elif t == b"B":
            st = buf[p:p + 1]; cnt = struct.unpack_from("<i", buf, p + 1)[0]
            p += 5 + cnt * _TAG_SIZE.get(st, 1)
        else:
            break
    return AS_MISSING


def read_bam(path, contigs: List[str], fragdict, remove_dups=True, proper_pair=True, min_mapq=0) -> ReadBatch:
    data = bgzf.read_all(path)
    if data[:4] != b"BAM\1":
        raise ValueError("%s is not a BAM file" % path)
    l_text = struct.unpack_from("<i", data, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", data, p)[0]; p += 4
    ref_names = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", data, p)[0]; p += 4
        ref_names.append(data[p:p + ln - 1].decode()); p += ln + 4
    cidx = {c: i for i, c in enumerate(contigs)}
    ref_to_contig = [cidx.get(n, -1) for n in ref_names]
    nc = len(contigs)
    per = [[] for _ in range(nc)]
    n = len(data)
    while p + 4 <= n:
        bs = struct.unpack_from("<i", data, p)[0]
        s = p + 4; p = s + bs
        ref_id, pos, l_rn, mapq, _bin, n_cig, flag, l_seq, _nref, _npos, tlen = struct.unpack_from("<iiBBHHHiiii", data, s)
        ci = ref_to_contig[ref_id] if ref_id >= 0 else -1
        if ci < 0:
            continue
        if remove_dups and (flag & 0x400):
            continue
        if proper_pair and not (flag & 2):
            continue
        if mapq < min_mapq:
            continue
        o = s + 32
        qname = data[o:o + l_rn - 1].decode(); o += l_rn
        cig = np.frombuffer(data, "<u4", n_cig, o); o += 4 * n_cig
        seq = data[o:o + (l_seq + 1) // 2]; o += (l_seq + 1) // 2
        qual = data[o:o + l_seq]; o += l_seq
        a = _find_as(data, o, p)
        per[ci].append((qname, pos + 1, tlen, a, cig, seq, qual, l_seq))
    recs = [r for pc in per for r in pc]
    R = len(recs)
    off = np.zeros(nc + 1, np.int64); off[1:] = np.cumsum([len(pc) for pc in per])
    posa = np.empty(R, np.int32); tl = np.empty(R, np.int32); aln = np.empty(R, np.int16); frag = np.empty(R, np.uint32)
    cig_off = np.zeros(R + 1, np.uint32); seq_off = np.zeros(R + 1, np.uint64)
    cigs = []; total = 0
    for i, (qn, ps, t, a, cg, sq, ql, ls) in enumerate(recs):
        posa[i] = ps; tl[i] = t
        if a != AS_MISSING and not (-32767 <= a <= 32767):
            raise ValueError("AS:i value %d outside the int16 range of the packed layout" % a)
        aln[i] = a; frag[i] = fragdict.get(qn)
        cigs.append(cg); cig_off[i + 1] = cig_off[i] + cg.shape[0]
        total += ls; seq_off[i + 1] = total
    # bases: BAM packs per record from a byte boundary; the device layout is one continuous nibble stream
    codes = np.empty(total, np.uint8); qual = np.empty(total, np.uint8)
    w = 0
    for (qn, ps, t, a, cg, sq, ql, ls) in recs:
        b = np.frombuffer(sq, np.uint8)
        u = np.empty(b.shape[0] * 2, np.uint8); u[0::2] = b >> 4; u[1::2] = b & 15
        codes[w:w + ls] = u[:ls]
        q = np.frombuffer(ql, np.uint8)
        if ls and q[0] == 0xFF:
            raise ValueError("record %s: QUAL missing (unsupported)" % qn)
        qual[w:w + ls] = q
        w += ls
    if total & 1:
        codes = np.concatenate([codes, np.zeros(1, np.uint8)])
    seq = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)
    cigar = np.concatenate(cigs).astype(np.uint32) if cigs else np.zeros(0, np.uint32)
    return ReadBatch(nc, off, posa, tl, aln, frag, cig_off, cigar, seq_off, seq, qual, fragdict.names)


def write_bam(path, ref_names_lengths, records):
    """records: iterable of (qname, flag, ref_index, pos1, mapq, cigar[(len, opchar)], seq str, qual bytes (phred),
    tlen, AS or None).  Minimal but valid BAM (no index)."""
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, l) for n, l in ref_names_lengths)
    with bgzf.BGZFWriter(path) as w:
        h = b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(ref_names_lengths))
        for n, l in ref_names_lengths:
            h += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
        w.write(h)
        for (qn, flag, ref, pos1, mapq, cigar, seq, qual, tlen, a) in records:
            rn = qn.encode() + b"\0"
            cg = b"".join(struct.pack("<I", (l << 4) | CIGAR_CODE[o]) for l, o in cigar)
            codes = [BASE_CODE.get(c, CODE_N) for c in seq]
            if len(codes) & 1:
                codes.append(0)
            sq = bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2))
            aux = b"NHC\x01" + (b"ASs" + struct.pack("<h", a) if a is not None else b"")
            body = struct.pack("<iiBBHHHiiii", ref, pos1 - 1, len(rn), mapq, 4680, len(cigar), flag, len(seq), ref,
                               pos1 - 1, tlen) + rn + cg + sq + bytes(qual) + aux
            w.write(struct.pack("<i", len(body)) + body)
