"""VCF ingest (host): the het-site filter of the reference -> VariantTable.

Mirrors, line for line in behaviour, the text pipeline + parser of the reference:
  * `gunzip -c VCF | cut -f 1-9,<sample col> | grep -v '0|0\\|1|1'`      phaser/phaser.py:205-225
  * the per-line het / PASS test                                           phaser/phaser.py:396-434
  * the per-contig mapping table rows (ids, indel exclusion, maf)          phaser/phaser.py:1355-1413
Fatal conditions keep the reference's messages (raised as PhaserFatal; the CLI prints
"     FATAL ERROR: ..." and exits 1 like phaser/phaser.py:2032-2034).
"""
import gzip
from collections import OrderedDict
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .layout import VariantTable, allele_code, is_indel_site, ALLELE_MULTI


class PhaserFatal(Exception):
    pass


@dataclass
class VcfStats:
    het_count: int = 0
    filter_count: int = 0
    indels_excluded: int = 0
    unphased_count: int = 0


def sample_column_map(path, start_col=9, line_key="#CHR"):
    """phaser/phaser.py:2326-2342"""
    out = OrderedDict()
    with gzip.open(path, "rt") as f:
        for line in f:
            if line_key in line:
                cols = line.rstrip().rstrip("\n").split("\t")
                for i in range(start_col, len(cols)):
                    out[cols[i]] = i
                break
    return out


def _annotation_to_dict(text, sep=";"):
    out = OrderedDict()
    for var in text.split(sep):
        if "=" in var:
            out[var.split("=")[0]] = var.split("=")[1]
    return out


class BedIntervals:
    """Interval overlap as `bedtools intersect` answers it for the two uses in phaser.py:220, 234: a VCF
    record occupies [POS-1, POS-1+len(REF)), BED intervals are [start, end) 0-based."""

    def __init__(self, path):
        import bisect
        self._bisect = bisect
        iv = {}
        with open(path) as f:
            for line in f:
                c = line.rstrip("\n").split("\t")
                if len(c) >= 3 and not line.startswith(("#", "track", "browser")):
                    iv.setdefault(c[0], []).append((int(c[1]), int(c[2])))
        self.starts = {}; self.maxend = {}
        for k, v in iv.items():
            v.sort()
            self.starts[k] = [a for a, _ in v]
            m = []; cur = -1
            for _, b in v:
                cur = max(cur, b); m.append(cur)
            self.maxend[k] = m

    def overlaps(self, chrom, pos, ref_len):
        s = pos - 1; e = s + ref_len
        st = self.starts.get(chrom)
        if not st:
            return False
        j = self._bisect.bisect_left(st, e)        # intervals with start < e
        return j > 0 and self.maxend[chrom][j - 1] > s


def parse_vcf(path, sample_column: int, pass_only=1, chrom_of_interest="", chr_prefix="", id_separator="_",
              include_indels=0, gw_phase_method=0, gw_af_field="AF", blacklist="", haplo_count_blacklist=""):
    """Returns (VariantTable, VcfStats).  `sample_column` is the 0-based VCF column of the sample.
    `blacklist`: BED of intervals whose variants are dropped before anything else (phaser.py:218-221);
    `haplo_count_blacklist`: BED whose variants are kept for phasing but left out of the haplotypic counts
    (phaser.py:231-243) -- the table's `haplo_blacklisted` flags."""
    bl = BedIntervals(blacklist) if blacklist else None
    hbl = BedIntervals(haplo_count_blacklist) if haplo_count_blacklist else None
    haplo_set = set()
    contig_ban = [id_separator, ":"]
    pool = OrderedDict()
    st = VcfStats()
    checked_chroms = set(); fmt_gt_index = {}; geno_cache = {}; pass_cache = {}
    with gzip.open(path, "rt") as f:
        for line in f:
            if line.startswith("#"):
                continue
            cols = line.rstrip("\n").split("\t", sample_column + 1)       # columns past the sample's are never looked at
            # cut -f 1-9,<col> ; grep -v '0|0\|1|1' acts on that cut line (anywhere in it)
            cut = cols[0:9] + ([cols[sample_column]] if sample_column < len(cols) else [])
            cut_line = "\t".join(cut)
            if "0|0" in cut_line or "1|1" in cut_line:
                continue
            if bl is not None and bl.overlaps(cut[0], int(cut[1]), len(cut[3])):
                continue
            chrom = cut[0]
            if hbl is not None and hbl.overlaps(chrom, int(cut[1]), len(cut[3])) and (chrom_of_interest == "" or chrom_of_interest == chrom):
                haplo_set.add(chrom + "_" + cut[1])
            if chrom not in checked_chroms:
                for item in contig_ban:
                    if item in chrom:
                        raise PhaserFatal("Character '%s' must not be present in contig name. Please change id separtor "
                                          "using --id_separator to a character not found in the contig names and try "
                                          "again." % item)
                checked_chroms.add(chrom)
            if chrom_of_interest == "" or chrom_of_interest == chrom:
                if chrom not in pool:
                    pool[chrom] = []
                # FORMAT, genotype and FILTER strings repeat line after line: each distinct one is analysed once
                gi = fmt_gt_index.get(cut[8])
                if gi is None:
                    fields = cut[8].split(":")
                    gi = fields.index("GT") if "GT" in fields else -1
                    fmt_gt_index[cut[8]] = gi
                if gi >= 0:
                    geno_string = cut[9].split(":")[gi]
                    ginfo = geno_cache.get(geno_string)
                    if ginfo is None:
                        xgeno = list(geno_string)
                        unphased = False; usable = False
                        if "." not in xgeno:
                            if "|" in xgeno:
                                xgeno.remove("|")
                            if "/" in xgeno:
                                xgeno.remove("/")
                                unphased = True
                            usable = len(set(xgeno)) > 1
                        ginfo = (xgeno, unphased, usable)
                        geno_cache[geno_string] = ginfo
                    xgeno, unphased, usable = ginfo
                    if usable:
                        ok = pass_cache.get(cut[6])
                        if ok is None:
                            ok = "PASS" in cut[6].split(";")
                            pass_cache[cut[6]] = ok
                        if pass_only == 0 or ok:
                            pool[chrom].append((cut, geno_string, list(xgeno)))
                            if unphased:
                                st.unphased_count += 1
                        else:
                            st.filter_count += 1
    contigs: List[str] = []
    off = [0]
    pos, a0, a1, rl = [], [], [], []
    ids, rsids, alls, gts, mafs = [], [], [], [], []
    for chrom, rows in pool.items():
        cname = chr_prefix + chrom
        contigs.append(cname)
        for cut, geno_string, xgeno in rows:
            alt = cut[4].split(",")
            all_alleles = [cut[3]] + alt
            maf: Optional[float] = None
            if gw_phase_method == 1:
                info = _annotation_to_dict(cut[7])
                if gw_af_field in info:
                    afs = list(map(float, info[gw_af_field].split(",")))
                    if len(afs) == len(alt):
                        use = [int(a) - 1 for a in xgeno if a != "." and int(a) != 0]
                        if use:
                            maf = min(min(afs[x], 1 - afs[x]) for x in use)
            # (common case first: one-base REF and one-base ALTs need none of the general length tests)
            all_single = len(cut[3]) == 1 and (len(cut[4]) == 1 or all(len(x) == 1 for x in alt))
            if all_single or max(len(x) for x in all_alleles) == 1 or include_indels == 1:        # phaser.py:1398-1400
                ind = [all_alleles[i] for i in range(len(all_alleles)) if str(i) in xgeno]
                if len(ind) != 2 or len(xgeno) != 2:
                    raise PhaserFatal("Variant %s:%s: only diploid genotypes with two distinct alleles are supported."
                                      % (chrom, cut[1]))
                pos.append(int(cut[1]))
                if not all_single and is_indel_site(cut[3], ind):
                    a0.append(ALLELE_MULTI); a1.append(ALLELE_MULTI)
                else:
                    a0.append(allele_code(ind[0])); a1.append(allele_code(ind[1]))
                rl.append(len(cut[3]))
                ids.append(cname + id_separator + cut[1] + id_separator + id_separator.join(all_alleles))
                rsids.append(cut[2]); alls.append(all_alleles); gts.append(geno_string); mafs.append(str(maf))
                st.het_count += 1
            else:
                st.indels_excluded += 1
        off.append(len(pos))
    vt = VariantTable(contigs, np.asarray(off, np.int64), np.asarray(pos, np.int32), np.asarray(a0, np.uint8),
                      np.asarray(a1, np.uint8), np.asarray(rl, np.int32), ids, rsids, alls, gts, mafs)
    # phaser.py:1070: key = contig name as in the ids (chr_prefix applied) + "_" + pos, against raw VCF names
    vt.haplo_blacklisted = np.zeros(len(pos), np.uint8)
    if haplo_set:
        for c in range(len(contigs)):
            for v in range(off[c], off[c + 1]):
                if contigs[c] + "_" + str(int(vt.pos[v])) in haplo_set:
                    vt.haplo_blacklisted[v] = 1
    for c in range(len(contigs)):
        p = vt.pos[off[c]:off[c + 1]]
        if p.shape[0] > 1 and np.any(np.diff(p) < 0):
            raise PhaserFatal("VCF records of contig %s are not sorted by position." % contigs[c])
    return vt, st


# ---------------------------------------------------------------------------------------------- native ingest

class NativeVcf:
    """The VCF held by the native library (include/phz.h: phz_vcf_open): inflated once (BGZF blocks in parallel), split
    into lines once, parsed on --threads host threads, and written back out from the same text (phz_vcf_write)."""

    def __init__(self, path, lib, threads=0):
        import ctypes
        import os
        self.lib = lib
        self.threads = int(threads or (os.cpu_count() or 1))
        self.h = lib.phz_vcf_open(path.encode(), self.threads)
        if not self.h:
            raise PhaserFatal(lib.phz_last_error().decode())
        p = ctypes.c_void_p(); n = ctypes.c_int64(0); nl = ctypes.c_int64(0); cr = ctypes.c_int(0)
        lib.phz_vcf_text(self.h, ctypes.byref(p), ctypes.byref(n), ctypes.byref(nl), ctypes.byref(cr))
        self.n_bytes = n.value; self.n_lines = nl.value; self.has_cr = bool(cr.value)
        self.text = (ctypes.c_char * max(1, n.value)).from_address(p.value) if n.value else b""
        self.mv = memoryview(self.text) if n.value else memoryview(b"")

    def close(self):
        if getattr(self, "h", None):
            self.mv = None; self.text = None
            self.lib.phz_vcf_close(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def line(self, off, n):
        return bytes(self.mv[off:off + n]).decode()

    def sample_column_map(self, start_col=9):
        """phaser/phaser.py:2326-2342 on the text already in memory"""
        import ctypes
        off = ctypes.c_int64(0); n = ctypes.c_int64(0)
        self.lib.phz_vcf_chrom_line(self.h, ctypes.byref(off), ctypes.byref(n))
        out = OrderedDict()
        if off.value >= 0:
            cols = self.line(off.value, n.value).rstrip().split("\t")
            for i in range(start_col, len(cols)):
                out[cols[i]] = i
        return out


class _LazyColumn:
    """One text column of the variant table, produced from the site's VCF line on demand: a whole-genome table has
    millions of sites, the outputs name a fraction of them."""

    def __init__(self, owner, kind):
        self.o = owner; self.kind = kind

    def __len__(self):
        return self.o.n

    def __bool__(self):
        return self.o.n > 0

    def __getitem__(self, v):
        if isinstance(v, slice):
            return [self[i] for i in range(*v.indices(self.o.n))]
        return self.o.get(int(v))[self.kind]

    def __iter__(self):
        for i in range(self.o.n):
            yield self[i]


class _LazyRows:
    def __init__(self, nv: NativeVcf, line_off, line_len, contig_name_of, sample_column, id_separator, gw_phase_method, gw_af_field):
        self.nv = nv; self.off = line_off; self.len = line_len; self.cname = contig_name_of
        self.col = sample_column; self.sep = id_separator; self.gwm = gw_phase_method; self.af = gw_af_field
        self.n = int(line_off.shape[0])
        self.cache = {}
        self.site_rows = {}        # v -> "POS\tID\tREF\tALT\tallele\tallele\tphase\tphase" (prefetch_sites)

    def prefetch_sites(self, sites):
        """One native call (phz_vcf_site_text, all host threads) for the text-side view of `sites`; the table writers then
        build their per-variant records from one split each instead of re-reading the VCF line in Python."""
        import ctypes
        if self.gwm == 1 or not hasattr(self.nv.lib, "phz_vcf_site_text"):
            return                  # the allele-frequency reading stays with get()
        idx = np.ascontiguousarray(np.asarray(sites, np.int64))
        idx = idx[~np.isin(idx, np.fromiter(self.site_rows.keys(), np.int64, len(self.site_rows)))] if self.site_rows else idx
        if idx.shape[0] == 0:
            return
        text = ctypes.c_void_p(); nb = ctypes.c_int64(0)
        rc = self.nv.lib.phz_vcf_site_text(self.nv.h, idx.ctypes.data, int(idx.shape[0]), self.nv.threads,
                                           ctypes.cast(ctypes.byref(text), ctypes.POINTER(ctypes.c_char_p)), ctypes.byref(nb))
        if rc != 0:
            raise PhaserFatal(self.nv.lib.phz_last_error().decode())
        rows = ctypes.string_at(text.value, nb.value).decode().split("\n")
        self.site_rows.update(zip(idx.tolist(), rows))

    def get(self, v):
        r = self.cache.get(v)
        if r is None:
            if v < 0:
                v += self.n
            cols = self.nv.line(int(self.off[v]), int(self.len[v])).split("\t", self.col + 1)
            cut = cols[0:9] + [cols[self.col]]
            gi = cut[8].split(":").index("GT")
            geno_string = cut[9].split(":")[gi]
            xgeno = list(geno_string)
            if "|" in xgeno:
                xgeno.remove("|")
            if "/" in xgeno:
                xgeno.remove("/")
            alt = cut[4].split(",")
            all_alleles = [cut[3]] + alt
            maf = None
            if self.gwm == 1:
                info = _annotation_to_dict(cut[7])
                if self.af in info:
                    afs = list(map(float, info[self.af].split(",")))
                    if len(afs) == len(alt):
                        use = [int(a) - 1 for a in xgeno if a != "." and int(a) != 0]
                        if use:
                            maf = min(min(afs[x], 1 - afs[x]) for x in use)
            r = (self.cname(v) + self.sep + cut[1] + self.sep + self.sep.join(all_alleles), cut[2], all_alleles, geno_string, str(maf))
            self.cache[v] = r
        return r


def parse_vcf_native(nv: NativeVcf, sample_column: int, pass_only=1, chrom_of_interest="", chr_prefix="", id_separator="_",
                     include_indels=0, gw_phase_method=0, gw_af_field="AF"):
    """parse_vcf through the native library (no blacklists: the caller takes parse_vcf for those).  Same VariantTable and
    VcfStats; the text columns (ids, rsids, all_alleles, gt, maf) are produced from the VCF lines on demand."""
    import ctypes
    from .engine import phz_vcf_table
    if nv.has_cr:
        raise PhaserFatal("native VCF ingest: carriage returns in the file")        # the caller falls back to parse_vcf
    t = phz_vcf_table()
    rc = nv.lib.phz_vcf_parse(nv.h, int(sample_column), int(pass_only), chrom_of_interest.encode(), int(include_indels),
                              nv.threads, ctypes.byref(t))
    if rc != 0:
        raise PhaserFatal(nv.lib.phz_last_error().decode())

    def names(ptr, n):
        out = []; p = ptr
        for _ in range(n):
            s = ctypes.string_at(p); out.append(s.decode()); p += len(s) + 1
        return out

    def arr(ptr, n, dt):
        if n == 0 or not ptr:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy()

    for chrom in names(t.seen_names, t.n_seen):                # phaser.py:404-408
        for item in (id_separator, ":"):
            if item in chrom:
                raise PhaserFatal("Character '%s' must not be present in contig name. Please change id separtor "
                                  "using --id_separator to a character not found in the contig names and try again." % item)
    V = int(t.n_variants); nc = int(t.n_contigs)
    raw_contigs = names(t.contig_names, nc)
    contigs = [chr_prefix + c for c in raw_contigs]
    off = arr(t.contig_var_off, nc + 1, np.int64)
    pos = arr(t.pos, V, np.int32)
    line_off = arr(t.var_line_off, V, np.int64); line_len = arr(t.var_line_len, V, np.int32)
    st = VcfStats(int(t.stats[0]), int(t.stats[1]), int(t.stats[2]), int(t.stats[3]))
    contig_idx = np.repeat(np.arange(nc, dtype=np.int64), np.diff(off)) if V else np.zeros(0, np.int64)
    rows = _LazyRows(nv, line_off, line_len, lambda v: contigs[int(contig_idx[v])], sample_column, id_separator, gw_phase_method, gw_af_field)
    vt = VariantTable(contigs, off, pos, arr(t.a0, V, np.uint8), arr(t.a1, V, np.uint8), arr(t.ref_len, V, np.int32),
                      _LazyColumn(rows, 0), _LazyColumn(rows, 1), _LazyColumn(rows, 2), _LazyColumn(rows, 3), _LazyColumn(rows, 4))
    vt.haplo_blacklisted = np.zeros(V, np.uint8)
    for c in range(nc):
        p = pos[off[c]:off[c + 1]]
        if p.shape[0] > 1 and np.any(np.diff(p) < 0):
            raise PhaserFatal("VCF records of contig %s are not sorted by position." % contigs[c])
    return vt, st
