"""Drop-in for the reference's read_variant_map module (seam S1).

    do_read_variant_map(variant_table, baseq, o, splice, isize_cutoff)

has the signature and the file protocol of phaser/read_variant_map.py:3 -- SAM text (with @SQ lines)
on stdin, the per-contig variant table written by generate_mapping_table (phaser/phaser.py:1402-1404)
as `variant_table`, the 7-column TSV of read_variant_map.py:117 written to `o` -- but the join runs in
the K1 kernel through the C ABI (phz_map_reads).  The reference's README has users swap in a compiled
`read_variant_map.so` (phaser/README.md:20-25); putting this module first on the import path does the
same for the GPU mapper, so the UNMODIFIED reference phaser.py can drive it.
"""
import sys

import numpy as np

from .layout import VariantTable, ReadBatch, BASE_ALPHABET, AS_MISSING, allele_code, is_indel_site, ALLELE_MULTI
from . import samio

_ENGINE = None


def set_engine(engine):
    """Tests inject the engine; by default the CUDA engine on cuda:0 is created on first use."""
    global _ENGINE
    _ENGINE = engine


def _engine():
    global _ENGINE
    if _ENGINE is None:
        from .engine import Engine
        _ENGINE = Engine(device="cuda:0")
    return _ENGINE


def read_variant_table(path):
    """Rows: chr, pos, unique_id, rsid, "ref,alt[,alt2]", len(ref), GT string, maf (phaser.py:1402-1404)."""
    contigs, off = [], [0]
    pos, a0, a1, rl, ids, rs, alls, gts, mafs = [], [], [], [], [], [], [], [], []
    with open(path) as f:
        for line in f:
            c = line.rstrip().split("\t")
            if len(c) < 8:
                continue
            if not contigs or contigs[-1] != c[0]:
                if c[0] in contigs:
                    raise ValueError("variant table is not grouped by contig")
                if contigs:
                    off.append(len(pos))
                contigs.append(c[0])
            alleles = c[4].split(",")
            g = list(c[6])
            for sep in "|/":
                if sep in g:
                    g.remove(sep)
            ind = [alleles[i] for i in range(len(alleles)) if str(i) in g]
            if len(ind) != 2:
                raise NotImplementedError("only diploid het sites with two distinct alleles are supported (variant %s)" % c[2])
            pos.append(int(c[1])); rl.append(int(c[5]))
            if int(c[5]) != 1 or is_indel_site(alleles[0], ind):        # --include_indels 1 tables (phaser.py:1398-1404)
                a0.append(ALLELE_MULTI); a1.append(ALLELE_MULTI)
            else:
                a0.append(allele_code(ind[0])); a1.append(allele_code(ind[1]))
            ids.append(c[2]); rs.append(c[3]); alls.append(alleles); gts.append(c[6]); mafs.append(c[7])
    off.append(len(pos))
    if not contigs:
        off = [0]
    return VariantTable(contigs, np.asarray(off, np.int64), np.asarray(pos, np.int32), np.asarray(a0, np.uint8),
                        np.asarray(a1, np.uint8), np.asarray(rl, np.int32), ids, rs, alls, gts, mafs)


def allele_string(batch: ReadBatch, r, seg_index, vpos, baseq, ref_len=1):
    """Text of a multi-base call (bases + inserted bases, read_variant_map.py:245-258) for the TSV."""
    lo = int(batch.seq_off[r])
    n_b = int(batch.seq_off[r + 1]) - lo
    codes = [(int(batch.seq[(lo + j) >> 1]) >> 4) if ((lo + j) & 1) == 0 else (int(batch.seq[(lo + j) >> 1]) & 15) for j in range(n_b)]
    bases = ["N" if int(batch.qual[lo + j]) < baseq else BASE_ALPHABET[codes[j]] for j in range(n_b)]
    g = q = 0; seg = 0; seg_start = 0; pseudo = []; ins = {}
    for k in range(int(batch.cigar_off[r]), int(batch.cigar_off[r + 1])):
        n = int(batch.cigar[k]) >> 4; op = int(batch.cigar[k]) & 15
        if op in (0, 7, 8):
            pseudo += bases[q:q + n]; q += n; g += n
        elif op == 2:
            pseudo += ["D"] * n; g += n
        elif op == 1:
            ins[g - 1] = "".join(bases[q:q + n]); q += n
        elif op == 4:
            q += n
        elif op == 3:
            if seg == seg_index:
                break
            g += n; seg += 1; seg_start = g; pseudo = []; ins = {}
    st = vpos - (int(batch.pos[r]) + seg_start)
    s = "".join(pseudo[x] + ins.get(x, "") for x in range(st, st + ref_len))
    return s.replace("D", "")


def do_read_variant_map(variant_table, baseq, o, splice, isize_cutoff):
    if int(splice) != 1:
        raise NotImplementedError("--splice 0 is never used by phaser.py (phaser/phaser.py:1346) and is not supported")
    vt = read_variant_table(variant_table)
    fd = samio.FragmentDictionary()
    # no flag / MAPQ filtering here: in the reference pipeline samtools has already done it (phaser.py:1346)
    batch = samio.parse_sam_stream(sys.stdin, vt.contigs, fd, remove_dups=False, proper_pair=False, min_mapq=0)
    samio.check_sorted(batch, "<stdin>")
    e = _engine()
    e.set_variants(vt)
    e.map_reads(e.upload_reads(batch), int(baseq), float(isize_cutoff))
    rec = e.download("t_rec"); var = e.download("t_var"); misc = e.download("t_misc")
    keep = (misc & 3) != 3
    with open(o, "w") as out:
        for r, v, m in zip(rec[keep].tolist(), var[keep].tolist(), misc[keep].tolist()):
            if (m >> 2) & 1:
                allele = allele_string(batch, r, (m >> 8) & 0xFF, int(vt.pos[v]), int(baseq), int(vt.ref_len[v]))
            else:
                allele = BASE_ALPHABET[(m >> 4) & 15]
            a = np.int16(np.uint16(m >> 16))
            out.write("\t".join([fd.names[int(batch.frag[r])], vt.ids[v], vt.rsids[v], allele,
                                 "" if a == AS_MISSING else str(int(a)), vt.gt[v], vt.maf[v]]) + "\n")
