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
