"""Tabix (.tbi) / CSI (.csi) index of the BGZF-compressed output VCF -- what `tabix -f -p vcf [--csi]` writes
after `bgzip -f` in the reference (phaser/phaser.py:1847-1853); done in-process here.

Formats: "The Tabix index file format" and "Coordinate Sorted Index (CSI) format" of the hts-specs.  Binning is
the UCSC scheme with min_shift = 14 and depth = 5 (bins of 16 kb .. 512 Mb); the linear index has one entry per
16 kb window.  Like htslib, every reference also carries the pseudo-bin 37450 with the span of its records and
their count.  A VCF record covers [POS-1, POS-1+len(REF)), or up to INFO/END when that is larger.
"""
import struct

from . import bgzf

MIN_SHIFT = 14
DEPTH = 5
META_BIN = ((1 << (3 * (DEPTH + 1))) - 1) // 7 + 1        # 37450


def reg2bin(beg, end):
    """bin of the 0-based half-open interval [beg, end)"""
    end -= 1
    s = MIN_SHIFT; t = ((1 << (3 * DEPTH)) - 1) // 7
    for level in range(DEPTH, 0, -1):
        if beg >> s == end >> s:
            return t + (beg >> s)
        s += 3; t -= 1 << (3 * (level - 1))
    return 0


def reg2bins(beg, end):
    """all bins that may hold records overlapping [beg, end)"""
    end -= 1
    out = [0]
    s = MIN_SHIFT + 3 * (DEPTH - 1); t = 1
    for level in range(1, DEPTH + 1):
        out += range(t + (beg >> s), t + (end >> s) + 1)
        s -= 3; t += 1 << (3 * level)
    return out


def bin_start_window(b):
    """first 16 kb window covered by bin b"""
    t = 0
    for level in range(DEPTH + 1):
        n = 1 << (3 * level)
        if b < t + n:
            return (b - t) << (3 * (DEPTH - level))
        t += n
    raise ValueError(b)


class IndexBuilder:
    def __init__(self):
        self.names = []
        self.refs = {}          # name -> dict(bins={bin: [[beg, end], ...]}, lin=[], first=, last=, n=)

    def add(self, chrom, beg, end, voff_beg, voff_end):
        r = self.refs.get(chrom)
        if r is None:
            r = dict(bins={}, lin=[], first=voff_beg, last=voff_end, n=0, last_bin=None)
            self.refs[chrom] = r; self.names.append(chrom)
        if end <= beg:
            end = beg + 1
        b = reg2bin(beg, end)
        chunks = r["bins"].setdefault(b, [])
        if r["last_bin"] == b and chunks and chunks[-1][1] == voff_beg:
            chunks[-1][1] = voff_end              # consecutive records of one bin: one chunk
        else:
            chunks.append([voff_beg, voff_end])
        r["last_bin"] = b
        w0, w1 = beg >> MIN_SHIFT, (end - 1) >> MIN_SHIFT
        lin = r["lin"]
        if len(lin) <= w1:
            lin += [0] * (w1 + 1 - len(lin))
        for w in range(w0, w1 + 1):
            if lin[w] == 0:
                lin[w] = voff_beg
        r["last"] = voff_end; r["n"] += 1

    def add_many(self, chrom, beg, end, v0, v1):
        """All records of one reference at once (numpy arrays, file order, consecutive lines): same index as add()
        record by record."""
        import numpy as np
        n = int(beg.shape[0])
        if n == 0:
            return
        if chrom in self.refs:             # a reference seen before (unsorted file): keep the scalar path
            for k in range(n):
                self.add(chrom, int(beg[k]), int(end[k]), int(v0[k]), int(v1[k]))
            return
        end = np.where(end <= beg, beg + 1, end)
        e1 = end - 1
        bins = np.zeros(n, np.int64); done = np.zeros(n, bool)
        s = MIN_SHIFT; t = ((1 << (3 * DEPTH)) - 1) // 7
        for level in range(DEPTH, 0, -1):
            hit = ~done & ((beg >> s) == (e1 >> s))
            bins[hit] = t + (beg[hit] >> s); done |= hit
            s += 3; t -= 1 << (3 * (level - 1))
        # consecutive records of one bin form one chunk (their lines are adjacent in the file)
        brk = np.flatnonzero((bins[1:] != bins[:-1]) | (v0[1:] != v1[:-1])) + 1
        rs = np.concatenate([[0], brk]); re_ = np.concatenate([brk, [n]]) - 1
        bd = {}
        for b, a, z in zip(bins[rs].tolist(), v0[rs].tolist(), v1[re_].tolist()):
            bd.setdefault(b, []).append([a, z])
        w0 = beg >> MIN_SHIFT; w1 = e1 >> MIN_SHIFT
        lin = np.zeros(int(w1.max()) + 1, np.int64)
        uw, first = np.unique(w0, return_index=True)
        lin[uw] = v0[first]
        for k in np.flatnonzero(w1 > w0).tolist():          # records spanning several 16 kb windows (rare)
            for w in range(int(w0[k]) + 1, int(w1[k]) + 1):
                if lin[w] == 0 or v0[k] < lin[w]:
                    lin[w] = v0[k]
        self.refs[chrom] = dict(bins=bd, lin=lin.tolist(), first=int(v0[0]), last=int(v1[-1]), n=n, last_bin=int(bins[-1]))
        self.names.append(chrom)

    def _finish_linear(self, r):
        lin = r["lin"]
        for i in range(1, len(lin)):            # windows without a record point at the previous one (htslib)
            if lin[i] == 0:
                lin[i] = lin[i - 1]
        return lin

    def _aux(self):
        names = b"".join(n.encode() + b"\0" for n in self.names)
        # format 2 = VCF, sequence column 1, begin column 2, end column 0, meta '#', skip 0
        return struct.pack("<7i", 2, 1, 2, 0, ord("#"), 0, len(names)) + names

    def tbi_bytes(self):
        out = [b"TBI\1", struct.pack("<i", len(self.names)), self._aux()]
        for n in self.names:
            r = self.refs[n]
            bins = r["bins"]
            out.append(struct.pack("<i", len(bins) + 1))
            for b in sorted(bins):
                out.append(struct.pack("<Ii", b, len(bins[b])))
                for c in bins[b]:
                    out.append(struct.pack("<QQ", c[0], c[1]))
            out.append(struct.pack("<Ii", META_BIN, 2) + struct.pack("<QQQQ", r["first"], r["last"], r["n"], 0))
            lin = self._finish_linear(r)
            out.append(struct.pack("<i", len(lin)) + struct.pack("<%dQ" % len(lin), *lin))
        out.append(struct.pack("<Q", 0))         # records without coordinates
        return b"".join(out)

    def csi_bytes(self):
        aux = self._aux()
        out = [b"CSI\1", struct.pack("<3i", MIN_SHIFT, DEPTH, len(aux)), aux, struct.pack("<i", len(self.names))]
        for n in self.names:
            r = self.refs[n]
            bins = r["bins"]; lin = self._finish_linear(r)
            out.append(struct.pack("<i", len(bins) + 1))
            for b in sorted(bins):
                w = bin_start_window(b)
                loff = lin[w] if w < len(lin) else (lin[-1] if lin else 0)
                out.append(struct.pack("<IQi", b, loff, len(bins[b])))
                for c in bins[b]:
                    out.append(struct.pack("<QQ", c[0], c[1]))
            out.append(struct.pack("<IQi", META_BIN, 0, 2) + struct.pack("<QQQQ", r["first"], r["last"], r["n"], 0))
        out.append(struct.pack("<Q", 0))
        return b"".join(out)


def record_span(cols):
    """0-based half-open reference span of a VCF data line (already split on tabs)"""
    beg = int(cols[1]) - 1
    end = beg + len(cols[3])
    if len(cols) > 7 and "END=" in cols[7]:
        for kv in cols[7].split(";"):
            if kv.startswith("END="):
                try:
                    end = max(end, int(kv[4:]))
                except ValueError:
                    pass
    return beg, end


def write_vcf_with_index(path_vcf_gz, text, csi=False, records=None):
    """bgzip + tabix of `text` (the whole VCF): writes path_vcf_gz and path_vcf_gz + '.tbi' (or '.csi').

    The BGZF writer cuts the byte stream into fixed 0xff00-byte blocks, so the virtual offset of every line follows
    from its byte offset and the compressed block sizes: the file is written in one go, the offsets of all data lines
    are computed at once and every reference is indexed with array operations.  `records` = (chroms, begs, ends) of
    the data lines in file order when the caller already has them (the VCF writer does); otherwise the lines are split."""
    import numpy as np
    data = text.encode() if isinstance(text, str) else text          # str, bytes or a memoryview of the native writer's text
    blocks = bgzf.compress_all(data)          # same bytes as BGZFWriter, blocks compressed in parallel
    with open(path_vcf_gz, "wb") as f:
        for b in blocks:
            f.write(b)
        f.write(bgzf.EOF_BLOCK)
    sizes = [len(b) for b in blocks]
    cstart = np.concatenate([[0], np.cumsum(np.asarray(sizes, np.int64))])      # compressed offset of block b
    buf = np.frombuffer(data, np.uint8)
    nl = np.flatnonzero(buf == 10)
    starts = np.concatenate([[0], nl + 1])
    if starts.shape[0] and starts[-1] >= len(data):
        starts = starts[:-1]
    ends = np.concatenate([nl + 1, [len(data)]])[:starts.shape[0]]
    is_data = (buf[starts] != ord("#")) & (ends - starts > 1) if starts.shape[0] else np.zeros(0, bool)

    def voff(u):
        b = u // bgzf.MAX_BLOCK
        return (cstart[b] << 16) | (u - b * bgzf.MAX_BLOCK)
    ds = starts[is_data]; de = ends[is_data]
    v0 = voff(ds); v1 = voff(de)
    if records is not None and len(records) == 4 and len(records[0]) == ds.shape[0]:
        # (chromosome index per line, names, begs, ends) as the native writer reports them: runs by array operations
        cidx = np.asarray(records[0], np.int64); beg = np.asarray(records[2], np.int64); end = np.asarray(records[3], np.int64)
        ib = IndexBuilder()
        if cidx.shape[0]:
            cut = np.concatenate([[0], np.flatnonzero(np.diff(cidx) != 0) + 1, [cidx.shape[0]]])
            for k, m in zip(cut[:-1].tolist(), cut[1:].tolist()):
                ib.add_many(records[1][int(cidx[k])], beg[k:m], end[k:m], v0[k:m], v1[k:m])
        idx = path_vcf_gz + (".csi" if csi else ".tbi")
        with bgzf.BGZFWriter(idx) as w:
            w.write(ib.csi_bytes() if csi else ib.tbi_bytes())
        return idx
    if records is not None and len(records) == 3 and len(records[0]) == ds.shape[0]:
        chroms = records[0]; beg = np.asarray(records[1], np.int64); end = np.asarray(records[2], np.int64)
    else:
        chroms = []; b_ = []; e_ = []
        mv = memoryview(data)
        for a, b in zip(ds.tolist(), de.tolist()):
            cols = bytes(mv[a:b]).decode().split("\t", 8)
            x, y = record_span(cols)
            chroms.append(cols[0]); b_.append(x); e_.append(y)
        beg = np.asarray(b_, np.int64); end = np.asarray(e_, np.int64)
    ib = IndexBuilder()
    k = 0; n = len(chroms)
    while k < n:                           # runs of one reference
        c = chroms[k]; m = k + 1
        while m < n and chroms[m] == c:
            m += 1
        ib.add_many(c, beg[k:m], end[k:m], v0[k:m], v1[k:m])
        k = m
    idx = path_vcf_gz + (".csi" if csi else ".tbi")
    with bgzf.BGZFWriter(idx) as w:
        w.write(ib.csi_bytes() if csi else ib.tbi_bytes())
    return idx


# ------------------------------------------------------------------------------------------------ reader (tests, spot checks)

def read_index(path):
    data = bgzf.read_all(path)
    magic = data[:4]
    o = 4
    csi = magic == b"CSI\1"
    if not csi and magic != b"TBI\1":
        raise ValueError("not a tabix / CSI index")
    if csi:
        min_shift, depth, l_aux = struct.unpack_from("<3i", data, o); o += 12
        aux = data[o:o + l_aux]; o += l_aux
        n_ref, = struct.unpack_from("<i", data, o); o += 4
        fmt = struct.unpack_from("<7i", aux, 0); names = aux[28:28 + fmt[6]]
    else:
        n_ref, = struct.unpack_from("<i", data, o); o += 4
        fmt = struct.unpack_from("<7i", data, o); o += 28
        names = data[o:o + fmt[6]]; o += fmt[6]
    names = [x.decode() for x in names.split(b"\0")[:-1]]
    refs = []
    for _ in range(n_ref):
        n_bin, = struct.unpack_from("<i", data, o); o += 4
        bins = {}; loff = {}
        for _ in range(n_bin):
            if csi:
                b, lo, nc = struct.unpack_from("<IQi", data, o); o += 16; loff[b] = lo
            else:
                b, nc = struct.unpack_from("<Ii", data, o); o += 8
            bins[b] = [struct.unpack_from("<QQ", data, o + 16 * k) for k in range(nc)]; o += 16 * nc
        lin = []
        if not csi:
            n_intv, = struct.unpack_from("<i", data, o); o += 4
            lin = list(struct.unpack_from("<%dQ" % n_intv, data, o)); o += 8 * n_intv
        refs.append(dict(bins=bins, lin=lin, loff=loff))
    return dict(csi=csi, format=fmt, names=names, refs=refs)


def query(path_vcf_gz, index, chrom, beg, end):
    """Lines of the indexed VCF overlapping the 0-based half-open region, found through the index only."""
    if chrom not in index["names"]:
        return []
    r = index["refs"][index["names"].index(chrom)]
    min_off = 0
    if not index["csi"] and r["lin"]:
        w = beg >> MIN_SHIFT
        min_off = r["lin"][w] if w < len(r["lin"]) else r["lin"][-1]
    chunks = []
    for b in reg2bins(beg, end):
        for c in r["bins"].get(b, []):
            if c[1] > min_off:
                chunks.append(c)
    chunks.sort()
    out = []
    rd = bgzf.VirtualReader(path_vcf_gz)
    seen = set()
    for c0, c1 in chunks:
        rd.seek(c0)
        while rd.tell() < c1:
            v = rd.tell()
            line = rd.readline()
            if not line:
                break
            if v in seen:
                continue
            seen.add(v)
            cols = line.decode().split("\t", 8)
            b0, e0 = record_span(cols)
            if cols[0] == chrom and b0 < end and e0 > beg:
                out.append(line.decode())
    return out
