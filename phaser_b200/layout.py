"""Packed SoA layouts shared by the host code, the C-ABI (include/phz.h) and the oracle.

Reads and variants are parsed ONCE into these arrays; every kernel works on them.

ReadBatch  -- the records of one BAM that survive the samtools-stage filters of the reference
              (phaser/phaser.py:505-513, 1346: -F 0x400, -f 2, -q MAPQ, contig in the VCF), in
              BAM order grouped by contig (contigs in VCF first-appearance order, the order in
              which the reference maps them, phaser/phaser.py:437-442, 532-533).
VariantTable -- the het sites the reference writes into its per-contig mapping tables
              (phaser/phaser.py:396-434, 1355-1413), sorted by (contig, VCF order).

Base codes are BAM's 4-bit alphabet "=ACMGRSVTWYHKDBN" so that SEQ travels nibble-packed.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

BASE_ALPHABET = "=ACMGRSVTWYHKDBN"
BASE_CODE = {c: i for i, c in enumerate(BASE_ALPHABET)}
CODE_N = 15
CODE_D = 13          # IUPAC 'D' -- the reference strips it like a deletion placeholder (read_variant_map.py:254)
ALLELE_NONE = 0xFF   # an allele string that no single read base can equal
ALLELE_MULTI = 0xFE  # --include_indels: multi-base REF or allele string, resolved through the indel side tables

# CIGAR op codes as in BAM: MIDNSHP=X
CIGAR_OPS = "MIDNSHP=X"
CIGAR_CODE = {c: i for i, c in enumerate(CIGAR_OPS)}

AS_MISSING = -32768  # record had no AS:i tag (read_variant_map.py:53-64 leaves "")

# tuple classes (read_variant_map.py:236-258 + phaser.py:1312-1324)
CLS_A0, CLS_A1, CLS_OTHER, CLS_NONE = 0, 1, 2, 3


def allele_code(s: str) -> int:
    """4-bit code of a single-base allele string, ALLELE_NONE if it cannot equal one read base."""
    if len(s) == 1 and s in BASE_CODE:
        return BASE_CODE[s]
    return ALLELE_NONE


@dataclass
class ReadBatch:
    """One BAM's filtered records.  All arrays are numpy (host) here; engine uploads them."""
    n_contigs: int
    contig_rec_off: np.ndarray   # i64[n_contigs+1]  records of contig c are [off[c], off[c+1])
    pos: np.ndarray              # i32[R]   1-based leftmost reference position (SAM POS)
    tlen: np.ndarray             # i32[R]   SAM TLEN (sign kept; the gate uses abs, read_variant_map.py:35)
    aln_score: np.ndarray        # i16[R]   AS:i value, AS_MISSING if absent
    frag: np.ndarray             # u32[R]   fragment id: one per distinct QNAME (mates share it)
    cigar_off: np.ndarray        # u32[R+1]
    cigar: np.ndarray            # u32[sum C]  BAM encoding len<<4|op
    seq_off: np.ndarray          # u64[R+1]  base offsets; base j of record r is nibble seq_off[r]+j
    seq: np.ndarray              # u8[ceil(total/2)]  4-bit packed, even nibble index = high nibble
    qual: np.ndarray             # u8[total]  phred (no +33)
    qnames: Optional[List[str]] = None   # per fragment id (host only, for text output)

    @property
    def n_records(self) -> int:
        return int(self.pos.shape[0])


@dataclass
class VariantTable:
    contigs: List[str]               # names as they appear in ids (chr_prefix applied)
    contig_var_off: np.ndarray       # i64[n_contigs+1]
    pos: np.ndarray                  # i32[V] 1-based VCF POS
    a0: np.ndarray                   # u8[V] base code of the sample's first allele (allele-index order)
    a1: np.ndarray                   # u8[V] base code of the sample's second allele
    ref_len: np.ndarray              # i32[V] len(REF)
    # host-side text metadata, index-aligned with the arrays above
    ids: List[str] = field(default_factory=list)        # chr_pos_ref_alt[...]  (phaser.py:1376)
    rsids: List[str] = field(default_factory=list)      # VCF ID column verbatim
    all_alleles: List[List[str]] = field(default_factory=list)
    gt: List[str] = field(default_factory=list)         # GT string verbatim (e.g. "0|1")
    maf: List[str] = field(default_factory=list)        # str(maf) as the mapping table carries it
    haplo_blacklisted: Optional[np.ndarray] = None      # u8[V] 1 = left out of haplotypic counts (phaser.py:1070)

    @property
    def n_variants(self) -> int:
        return int(self.pos.shape[0])


def is_indel_site(ref: str, sample_alleles) -> bool:
    """A site the SNV fast path cannot call: REF spans several bases or one of the sample's alleles is a string."""
    return len(ref) != 1 or any(len(a) != 1 for a in sample_alleles)


def sample_alleles(all_alleles, gt: str):
    """The alleles the sample carries, in allele-index order (phaser.py:1431-1435)."""
    g = [ch for ch in gt if ch not in "|/"]
    return [all_alleles[i] for i in range(len(all_alleles)) if str(i) in g]


def indel_tables(vt: "VariantTable"):
    """(ref_len i32[V], al_off u32[2V+1], al_codes u8[]) for phz_set_indel_alleles, or None when every site is a
    plain SNV.  Strings are stored only for the sites flagged ALLELE_MULTI."""
    if not bool(np.any(vt.a0 == ALLELE_MULTI)):
        return None
    V = vt.n_variants
    off = np.zeros(2 * V + 1, np.uint32)
    codes = []
    for v in np.nonzero(vt.a0 == ALLELE_MULTI)[0].tolist():
        ind = sample_alleles(vt.all_alleles[v], vt.gt[v])
        off[2 * v + 1] = len(ind[0]); off[2 * v + 2] = len(ind[1])
    lens = off[1:].copy()
    off[1:] = np.cumsum(lens, dtype=np.uint64).astype(np.uint32)
    for v in np.nonzero(vt.a0 == ALLELE_MULTI)[0].tolist():
        for a in sample_alleles(vt.all_alleles[v], vt.gt[v]):
            codes.extend(BASE_CODE.get(ch, ALLELE_NONE) for ch in a)
    return (np.ascontiguousarray(vt.ref_len, np.int32), off, np.asarray(codes if codes else [0], np.uint8))


def pack_nibbles(codes: np.ndarray) -> np.ndarray:
    """u8 base codes (one per base) -> BAM-style packed nibbles (even index = high nibble)."""
    n = codes.shape[0]
    if n & 1:
        codes = np.concatenate([codes, np.zeros(1, np.uint8)])
    return ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)


def unpack_nibbles(packed: np.ndarray, n: int) -> np.ndarray:
    out = np.empty(packed.shape[0] * 2, np.uint8)
    out[0::2] = packed >> 4
    out[1::2] = packed & 15
    return out[:n]
