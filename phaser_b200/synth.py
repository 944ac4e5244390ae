"""Synthetic genome / variant / read generator, device-agnostic torch.

One seed -> (i) the packed SoA arrays the kernels consume, (ii) SAM text + VCF text twins for the
oracle and for the reference harness.  Not on the hot path: it only manufactures inputs.

RNA-seq shape: genes with 2-8 exons (120-600 bp) and 200-3000 bp introns, Pareto expression,
paired reads in transcript coordinates (-> M/N CIGARs), 1 % small indels, 2 % soft clips, 0.3 %
substitutions at het sites, 0.2 %/base chimeric haplotype switches (forces conflicting edges),
5 % low-quality bases, AS ~ U(120,150)*Lr/76.
"""
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np
import torch

from .layout import ReadBatch, VariantTable, BASE_ALPHABET, CIGAR_OPS, pack_nibbles

GRCH38 = [("1", 248956422), ("2", 242193529), ("3", 198295559), ("4", 190214555), ("5", 181538259),
          ("6", 170805979), ("7", 159345973), ("8", 145138636), ("9", 138394717), ("10", 133797422),
          ("11", 135086622), ("12", 133275309), ("13", 114364328), ("14", 107043718), ("15", 101991189),
          ("16", 90338345), ("17", 83257441), ("18", 80373285), ("19", 58617616), ("20", 64444167),
          ("21", 46709983), ("22", 50818468), ("X", 156040895), ("Y", 57227415)]

OP_M, OP_I, OP_D, OP_N, OP_S = 0, 1, 2, 3, 4
MAX_EXONS = 8


@dataclass
class Genome:
    contigs: List[Tuple[str, int]]
    # variants, sorted by (contig, pos)
    v_contig: torch.Tensor     # i64[V]
    v_pos: torch.Tensor        # i64[V] 1-based
    v_ref: torch.Tensor        # u8[V]  base code (1,2,4,8)
    v_alt: torch.Tensor        # u8[V]
    v_hap0_alt: torch.Tensor   # bool[V] true phase: haplotype 0 carries ALT
    v_phased: torch.Tensor     # bool[V] GT written with '|'
    v_gt_first_alt: torch.Tensor  # bool[V] GT string is "1|0" (or "1/0" never; unphased -> "0/1")
    v_named: torch.Tensor      # bool[V] ID is rs<i> (else ".")
    v_af: torch.Tensor         # f32[V]
    # genes
    g_contig: torch.Tensor     # i64[G]
    g_nexon: torch.Tensor      # i64[G]
    exon_start: torch.Tensor   # i64[G,8] 0-based genome start
    exon_len: torch.Tensor     # i64[G,8] 0 beyond n_exons
    tcum: torch.Tensor         # i64[G,9] transcript offset of each exon start; [:, n] = transcript length
    g_weight: torch.Tensor     # f64[G] expression

    @property
    def device(self):
        return self.v_pos.device


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def _randint(lo, hi, shape, g, device):
    return torch.randint(lo, hi, shape, generator=g, device=device, dtype=torch.int64)


def _rand(shape, g, device):
    return torch.rand(shape, generator=g, device=device, dtype=torch.float64)


def make_genome(seed, n_variants, n_genes=None, contigs=None, exonic_frac=1.0, device="cpu",
                phased_frac=0.9, phase_error=0.03):
    contigs = list(contigs if contigs is not None else GRCH38)
    g = _gen(seed, device)
    n_genes = int(n_genes if n_genes is not None else max(1, n_variants // 8))
    clen = torch.tensor([c[1] for c in contigs], dtype=torch.float64, device=device)
    # ---- genes
    g_contig = torch.multinomial(clen / clen.sum(), n_genes, replacement=True, generator=g)
    g_nexon = _randint(2, MAX_EXONS + 1, (n_genes,), g, device)
    exon_len = _randint(120, 601, (n_genes, MAX_EXONS), g, device)
    intron_len = _randint(200, 3001, (n_genes, MAX_EXONS), g, device)
    col = torch.arange(MAX_EXONS, device=device)[None, :]
    live = col < g_nexon[:, None]
    exon_len = exon_len * live
    intron_len = intron_len * live
    step = exon_len + intron_len
    span = step.sum(1)
    room = (clen[g_contig] - span.double() - 2000.0).clamp(min=1.0)
    g_start = 1000 + (_rand((n_genes,), g, device) * room).long()
    exon_start = g_start[:, None] + torch.cumsum(step, 1) - step
    tcum = torch.zeros((n_genes, MAX_EXONS + 1), dtype=torch.int64, device=device)
    tcum[:, 1:] = torch.cumsum(exon_len, 1)
    u = _rand((n_genes,), g, device).clamp(min=1e-9)
    g_weight = u.pow(-1.0 / 1.2)              # Pareto(1.2)
    # ---- variants
    n_ex = int(round(n_variants * exonic_frac))
    n_bg = n_variants - n_ex
    tlen = tcum[:, -1].double()
    vg = torch.multinomial(tlen / tlen.sum(), n_ex, replacement=True, generator=g) if n_ex > 0 else \
        torch.zeros(0, dtype=torch.int64, device=device)
    toff = (_rand((n_ex,), g, device) * tlen[vg]).long()
    ve = (tcum[vg, 1:] <= toff[:, None]).sum(1)
    vpos_ex = exon_start[vg, ve] + (toff - tcum[vg, ve]) + 1
    vcon_ex = g_contig[vg]
    vcon_bg = torch.multinomial(clen / clen.sum(), n_bg, replacement=True, generator=g) if n_bg > 0 else \
        torch.zeros(0, dtype=torch.int64, device=device)
    vpos_bg = 1 + (_rand((n_bg,), g, device) * (clen[vcon_bg] - 1)).long()
    v_contig = torch.cat([vcon_ex, vcon_bg])
    v_pos = torch.cat([vpos_ex, vpos_bg])
    key = torch.unique(v_contig * (1 << 32) + v_pos)   # sorted, duplicates dropped
    v_contig = key >> 32
    v_pos = key & 0xFFFFFFFF
    V = key.shape[0]
    code = torch.tensor([1, 2, 4, 8], dtype=torch.uint8, device=device)
    r = _randint(0, 4, (V,), g, device)
    a = (r + _randint(1, 4, (V,), g, device)) % 4
    hap0_alt = _rand((V,), g, device) < 0.5
    phased = _rand((V,), g, device) < phased_frac
    wrong = _rand((V,), g, device) < phase_error
    return Genome(contigs, v_contig, v_pos, code[r], code[a], hap0_alt, phased, hap0_alt ^ wrong,
                  _rand((V,), g, device) < 0.95, (0.01 + 0.49 * _rand((V,), g, device)).float(),
                  g_contig, g_nexon, exon_start, exon_len, tcum, g_weight)


def _overlay(genome, bases, rec_contig, hap, sw, blocks, err_rate, g):
    """Write the haplotype's allele into every read base that sits on a het site."""
    device = bases.device
    vkey = genome.v_contig * (1 << 32) + (genome.v_pos - 1)
    n, lr = bases.shape
    code = torch.tensor([1, 2, 4, 8], dtype=torch.uint8, device=device)
    for gs, qs, ln in blocks:
        lo = torch.searchsorted(vkey, rec_contig * (1 << 32) + gs)
        hi = torch.searchsorted(vkey, rec_contig * (1 << 32) + gs + ln)
        cnt = (hi - lo).clamp(min=0) * (ln > 0)
        tot = int(cnt.sum())
        if tot == 0:
            continue
        ridx = torch.repeat_interleave(torch.arange(n, device=device), cnt)
        first = torch.cumsum(cnt, 0) - cnt
        vidx = lo[ridx] + (torch.arange(tot, device=device) - first[ridx])
        q = qs[ridx] + (genome.v_pos[vidx] - 1 - gs[ridx])
        h = hap[ridx] ^ (q >= sw[ridx])
        use_alt = (h == 0) == genome.v_hap0_alt[vidx]
        b = torch.where(use_alt, genome.v_alt[vidx], genome.v_ref[vidx])
        err = _rand((tot,), g, device) < err_rate
        b = torch.where(err, code[_randint(0, 4, (tot,), g, device)], b)
        bases[ridx, q] = b


def make_reads(genome, seed, n_pairs, read_len=76, dup_frac=0.0, lowq_frac=0.05, indel_frac=0.01,
               clip_frac=0.02, switch_per_base=0.002, err_rate=0.003, insert_lo=150, insert_hi=350,
               lowmapq_frac=0.0, mapq=255, chunk_pairs=2_000_000):
    """Raw (pre-filter) records of one BAM, coordinate sorted.  Returns a dict of torch tensors."""
    device = genome.device
    g = _gen(seed, device)
    lr = read_len
    parts = []
    done = 0
    tl = genome.tcum[:, -1]
    w = genome.g_weight * (tl >= max(insert_lo, lr)).double()
    w = w / w.sum()
    while done < n_pairs:
        n = min(chunk_pairs, n_pairs - done)
        gi = torch.multinomial(w, n, replacement=True, generator=g)
        ins = torch.minimum(_randint(max(insert_lo, lr), max(insert_hi, lr) + 1, (n,), g, device), tl[gi])
        s = (_rand((n,), g, device) * (tl[gi] - ins + 1).double()).long()
        hap_pair = _randint(0, 2, (n,), g, device)
        # two records per pair
        gi2 = torch.cat([gi, gi]); a = torch.cat([s, s + ins - lr]); hap = torch.cat([hap_pair, hap_pair])
        pair = torch.cat([torch.arange(n, device=device), torch.arange(n, device=device)]) + done
        mate2 = torch.cat([torch.zeros(n, dtype=torch.bool, device=device), torch.ones(n, dtype=torch.bool, device=device)])
        N = 2 * n
        tc = genome.tcum[gi2]
        e = (tc[:, 1:] <= a[:, None]).sum(1)
        o = a - tc.gather(1, e[:, None])[:, 0]
        el = genome.exon_len[gi2]; es = genome.exon_start[gi2]
        rem = torch.full((N,), lr, dtype=torch.int64, device=device)
        blk_gs, blk_len, gap = [], [], []
        prev_end = None
        for k in range(3):
            ek = (e + k).clamp(max=MAX_EXONS - 1)
            valid = (e + k) < genome.g_nexon[gi2]
            avail = el.gather(1, ek[:, None])[:, 0] - (o if k == 0 else 0)
            ln = torch.minimum(rem, avail) * valid
            gs = es.gather(1, ek[:, None])[:, 0] + (o if k == 0 else 0)
            blk_gs.append(gs); blk_len.append(ln)
            gap.append(torch.zeros_like(gs) if prev_end is None else (gs - prev_end) * (ln > 0))
            prev_end = gs + ln
            rem = rem - ln
        # soft clip at the read start
        clip = (_rand((N,), g, device) < clip_frac)
        ck = _randint(1, 11, (N,), g, device)
        clip = clip & (blk_len[0] > ck + 5)
        ck = ck * clip
        blk_len[0] = blk_len[0] - ck
        blk_gs[0] = blk_gs[0] + ck
        # one small indel inside the last aligned block
        nblk = (blk_len[0] > 0).long() + (blk_len[1] > 0).long() + (blk_len[2] > 0).long()
        last = nblk - 1
        L_last = torch.stack(blk_len, 1).gather(1, last[:, None])[:, 0]
        ik = _randint(1, 4, (N,), g, device)
        has = (_rand((N,), g, device) < indel_frac) & (L_last >= ik + 4)
        is_ins = _rand((N,), g, device) < 0.5
        x = 2 + (_rand((N,), g, device) * (L_last - ik - 3).clamp(min=1).double()).long()
        x = torch.minimum(x, (L_last - ik - 2).clamp(min=2))
        # op table: per block j: [N_j] M_ja (I|D)_j M_jb ; column 0 is the leading S
        ops = torch.zeros((N, 13), dtype=torch.int64, device=device)
        opl = torch.zeros((N, 13), dtype=torch.int64, device=device)
        ops[:, 0] = OP_S; opl[:, 0] = ck
        blocks = []
        q = ck.clone()
        for j in range(3):
            base = 1 + 4 * j
            ops[:, base] = OP_N; opl[:, base] = gap[j]
            here = has & (last == j)
            ins_here = here & is_ins
            del_here = here & ~is_ins
            la = torch.where(here, x, blk_len[j])
            lb = torch.where(ins_here, blk_len[j] - x - ik, torch.where(del_here, blk_len[j] - x, torch.zeros_like(x)))
            ops[:, base + 1] = OP_M; opl[:, base + 1] = la
            ops[:, base + 2] = torch.where(is_ins, OP_I, OP_D); opl[:, base + 2] = ik * here
            ops[:, base + 3] = OP_M; opl[:, base + 3] = lb
            blocks.append((blk_gs[j], q.clone(), la))
            gs_b = blk_gs[j] + x + torch.where(del_here, ik, torch.zeros_like(ik))
            qs_b = q + x + torch.where(ins_here, ik, torch.zeros_like(ik))
            blocks.append((gs_b, qs_b, lb))
            q = q + blk_len[j]
        rec_contig = genome.g_contig[gi2]
        code = torch.tensor([1, 2, 4, 8], dtype=torch.uint8, device=device)
        bases = code[_randint(0, 4, (N, lr), g, device)]
        sw = torch.where(_rand((N,), g, device) < min(1.0, switch_per_base * lr),
                         _randint(1, lr, (N,), g, device), torch.full((N,), lr + 1, dtype=torch.int64, device=device))
        _overlay(genome, bases, rec_contig, hap, sw, blocks, err_rate, g)
        qual = torch.where(_rand((N, lr), g, device) < lowq_frac, 2, 37).to(torch.uint8)
        pos = blk_gs[0] + 1
        ref_end = torch.stack([b[0] + b[2] for b in blocks], 1).max(1)[0]
        start1 = pos[:n]; end2 = ref_end[n:]
        tl_pair = end2 - start1 + 1
        tlen = torch.cat([tl_pair, -tl_pair])
        flag = torch.where(mate2, 147, 99) | torch.where(_rand((N,), g, device) < dup_frac, 0x400, 0)
        mq = torch.where(_rand((N,), g, device) < lowmapq_frac, _randint(0, 20, (N,), g, device),
                         torch.full((N,), mapq, dtype=torch.int64, device=device))
        aln = _randint(120, 151, (N,), g, device) * lr // 76
        parts.append(dict(contig=rec_contig, pos=pos, tlen=tlen, flag=flag, mapq=mq, aln=aln, frag=pair,
                          ops=ops, opl=opl, bases=bases, qual=qual))
        done += n
    rec = {k: torch.cat([p[k] for p in parts]) for k in parts[0]}
    order = torch.argsort(rec["contig"] * (1 << 32) + rec["pos"], stable=True)
    rec = {k: v[order] for k, v in rec.items()}
    rec["read_len"] = lr
    return rec


def make_wgs_reads(genome, seed, n_pairs, read_len=150, insert_mean=400.0, insert_sd=60.0, **kw):
    """WGS shape: unspliced pairs uniform over the genome (a single-exon 'gene' per contig)."""
    device = genome.device
    nc = len(genome.contigs)
    clen = torch.tensor([c[1] for c in genome.contigs], dtype=torch.int64, device=device)
    exon_len = torch.zeros((nc, MAX_EXONS), dtype=torch.int64, device=device); exon_len[:, 0] = clen - 1
    exon_start = torch.zeros((nc, MAX_EXONS), dtype=torch.int64, device=device)
    tcum = torch.zeros((nc, MAX_EXONS + 1), dtype=torch.int64, device=device); tcum[:, 1:] = (clen - 1)[:, None]
    flat = Genome(genome.contigs, genome.v_contig, genome.v_pos, genome.v_ref, genome.v_alt, genome.v_hap0_alt,
                  genome.v_phased, genome.v_gt_first_alt, genome.v_named, genome.v_af,
                  torch.arange(nc, device=device), torch.ones(nc, dtype=torch.int64, device=device),
                  exon_start, exon_len, tcum, clen.double())
    lo = int(max(read_len, insert_mean - 3 * insert_sd)); hi = int(insert_mean + 3 * insert_sd)
    return make_reads(flat, seed, n_pairs, read_len=read_len, insert_lo=lo, insert_hi=hi, **kw)


# ----------------------------------------------------------------------------------------------
# conversions

def filter_raw(rec, remove_dups=True, proper_pair=True, min_mapq=0):
    """The samtools-stage filters of the reference pipeline (phaser.py:505-513, 1346)."""
    keep = rec["mapq"] >= min_mapq
    if remove_dups:
        keep &= (rec["flag"] & 0x400) == 0
    if proper_pair:
        keep &= (rec["flag"] & 2) == 2
    return {k: (v[keep] if torch.is_tensor(v) else v) for k, v in rec.items()}


def pack_records(rec, n_contigs):
    """dict of torch tensors (make_reads) -> dict of packed SoA torch tensors on the same device."""
    device = rec["pos"].device
    N = rec["pos"].shape[0]; lr = rec["read_len"]
    live = rec["opl"] > 0
    ncig = live.sum(1)
    cigar_off = torch.zeros(N + 1, dtype=torch.int64, device=device)
    cigar_off[1:] = torch.cumsum(ncig, 0)
    cig = ((rec["opl"][live].to(torch.int32) << 4) | rec["ops"][live].to(torch.int32))
    seq_off = torch.arange(N + 1, dtype=torch.int64, device=device) * lr
    flat = rec["bases"].reshape(-1)
    if flat.shape[0] & 1:
        flat = torch.cat([flat, torch.zeros(1, dtype=torch.uint8, device=device)])
    seq = (flat[0::2] << 4) | flat[1::2]
    counts = torch.bincount(rec["contig"].to(torch.int64), minlength=n_contigs)
    contig_rec_off = torch.zeros(n_contigs + 1, dtype=torch.int64, device=device)
    contig_rec_off[1:] = torch.cumsum(counts, 0)
    return dict(contig_rec_off=contig_rec_off, pos=rec["pos"].to(torch.int32), tlen=rec["tlen"].to(torch.int32),
                aln_score=rec["aln"].to(torch.int16), frag=rec["frag"].to(torch.int32),
                cigar_off=cigar_off.to(torch.int32), cigar=cig.to(torch.int32),
                seq_off=seq_off, seq=seq.contiguous(), qual=rec["qual"].reshape(-1).contiguous())


def compact_raw(rec):
    """Shrink the dtypes of a make_reads chunk (bench-scale generation keeps ~200 B per record)."""
    out = dict(rec)
    out["contig"] = rec["contig"].to(torch.int16)
    out["pos"] = rec["pos"].to(torch.int32); out["tlen"] = rec["tlen"].to(torch.int32)
    out["flag"] = rec["flag"].to(torch.int16); out["mapq"] = rec["mapq"].to(torch.int16)
    out["aln"] = rec["aln"].to(torch.int16); out["frag"] = rec["frag"].to(torch.int32)
    out["ops"] = rec["ops"].to(torch.uint8); out["opl"] = rec["opl"].to(torch.int16)
    return out


def concat_sorted(parts):
    """Concatenate chunks and restore the global coordinate order."""
    rec = {k: (torch.cat([p[k] for p in parts]) if torch.is_tensor(parts[0][k]) else parts[0][k]) for k in parts[0]}
    order = torch.argsort(rec["contig"].to(torch.int64) * (1 << 32) + rec["pos"].to(torch.int64), stable=True)
    return {k: (v[order] if torch.is_tensor(v) else v) for k, v in rec.items()}


def to_variant_table_arrays(genome) -> VariantTable:
    """VariantTable with the numeric columns only (bench scale: no per-variant Python strings)."""
    vc = genome.v_contig.cpu().numpy()
    nc = len(genome.contigs)
    off = np.zeros(nc + 1, np.int64); off[1:] = np.cumsum(np.bincount(vc, minlength=nc))
    V = vc.shape[0]
    return VariantTable([c[0] for c in genome.contigs], off, genome.v_pos.cpu().numpy().astype(np.int32),
                        genome.v_ref.cpu().numpy().astype(np.uint8), genome.v_alt.cpu().numpy().astype(np.uint8),
                        np.ones(V, np.int32))


def to_read_batch(rec, n_contigs, bam_name="bam0") -> ReadBatch:
    p = {k: v.cpu().numpy() for k, v in pack_records(rec, n_contigs).items()}
    # fragment ids -> dense, first-seen order; names stay "<bam>.<pair index>"
    uniq, inv = np.unique(p["frag"], return_inverse=True)
    first = np.full(uniq.shape[0], np.iinfo(np.int64).max, np.int64)
    np.minimum.at(first, inv, np.arange(inv.shape[0]))
    rank = np.empty(uniq.shape[0], np.int64); rank[np.argsort(first, kind="stable")] = np.arange(uniq.shape[0])
    dense = rank[inv]
    names = [None] * uniq.shape[0]
    for u, r in zip(uniq.tolist(), rank.tolist()):
        names[r] = "%s.%d" % (bam_name, u)
    return ReadBatch(n_contigs, p["contig_rec_off"].astype(np.int64), p["pos"].astype(np.int32),
                     p["tlen"].astype(np.int32), p["aln_score"].astype(np.int16), dense.astype(np.uint32),
                     p["cigar_off"].astype(np.uint32), p["cigar"].astype(np.uint32),
                     p["seq_off"].astype(np.uint64), p["seq"].astype(np.uint8), p["qual"].astype(np.uint8), names)


def to_variant_table(genome) -> VariantTable:
    vc = genome.v_contig.cpu().numpy(); vp = genome.v_pos.cpu().numpy()
    r = genome.v_ref.cpu().numpy(); a = genome.v_alt.cpu().numpy()
    ph = genome.v_phased.cpu().numpy(); fa = genome.v_gt_first_alt.cpu().numpy(); nm = genome.v_named.cpu().numpy()
    nc = len(genome.contigs)
    off = np.zeros(nc + 1, np.int64); off[1:] = np.cumsum(np.bincount(vc, minlength=nc))
    names = [c[0] for c in genome.contigs]
    ids, rs, al, gt = [], [], [], []
    for i in range(vp.shape[0]):
        rb = BASE_ALPHABET[r[i]]; ab = BASE_ALPHABET[a[i]]
        ids.append("%s_%d_%s_%s" % (names[vc[i]], vp[i], rb, ab))
        rs.append("rs%d" % (i + 1) if nm[i] else ".")
        al.append([rb, ab])
        gt.append(("1|0" if fa[i] else "0|1") if ph[i] else "0/1")
    return VariantTable(names, off, vp.astype(np.int32), r.astype(np.uint8), a.astype(np.uint8),
                        np.ones(vp.shape[0], np.int32), ids, rs, al, gt, ["None"] * vp.shape[0])


def write_vcf(genome, path, sample="S1", extra_lines=()):
    """VCF text twin (gzip).  `extra_lines` lets tests add hom / non-PASS / indel records."""
    import gzip
    vt = to_variant_table(genome)
    af = genome.v_af.cpu().numpy()
    vc = genome.v_contig.cpu().numpy()
    rows = []
    for i in range(vt.n_variants):
        rows.append((int(vc[i]), int(vt.pos[i]), "%s\t%d\t%s\t%s\t%s\t100\tPASS\tAF=%.4f\tGT\t%s" % (
            vt.contigs[vc[i]], vt.pos[i], vt.rsids[i], vt.all_alleles[i][0], vt.all_alleles[i][1], af[i], vt.gt[i])))
    cidx = {c: i for i, c in enumerate(vt.contigs)}
    for ln in extra_lines:
        c = ln.split("\t")
        rows.append((cidx[c[0]], int(c[1]), ln))
    rows.sort(key=lambda t: (t[0], t[1]))
    with gzip.open(path, "wt") as f:
        f.write("##fileformat=VCFv4.2\n")
        for name, ln in genome.contigs:
            f.write("##contig=<ID=%s,length=%d>\n" % (name, ln))
        f.write('##INFO=<ID=AF,Number=A,Type=Float,Description="Allele frequency">\n')
        f.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n" % sample)
        for _, _, ln in rows:
            f.write(ln + "\n")
    return path


def write_sam(rec, genome, path, bam_name="bam0"):
    """SAM text twin of the raw records (what `samtools view -h` would print)."""
    c = rec["contig"].cpu().numpy(); pos = rec["pos"].cpu().numpy(); tl = rec["tlen"].cpu().numpy()
    fl = rec["flag"].cpu().numpy(); mq = rec["mapq"].cpu().numpy(); aln = rec["aln"].cpu().numpy()
    fr = rec["frag"].cpu().numpy(); ops = rec["ops"].cpu().numpy(); opl = rec["opl"].cpu().numpy()
    bases = rec["bases"].cpu().numpy(); qual = rec["qual"].cpu().numpy()
    lut = np.frombuffer(BASE_ALPHABET.encode(), np.uint8)
    names = [x[0] for x in genome.contigs]
    with open(path, "w") as f:
        f.write("@HD\tVN:1.6\tSO:coordinate\n")
        for name, ln in genome.contigs:
            f.write("@SQ\tSN:%s\tLN:%d\n" % (name, ln))
        for i in range(pos.shape[0]):
            cg = "".join("%d%s" % (opl[i, j], CIGAR_OPS[ops[i, j]]) for j in range(ops.shape[1]) if opl[i, j] > 0)
            f.write("%s.%d\t%d\t%s\t%d\t%d\t%s\t=\t%d\t%d\t%s\t%s\tNH:i:1\tAS:i:%d\n" % (
                bam_name, fr[i], fl[i], names[c[i]], pos[i], mq[i], cg, max(1, pos[i] + tl[i]) if tl[i] > 0 else pos[i],
                tl[i], lut[bases[i]].tobytes().decode(), (qual[i] + 33).astype(np.uint8).tobytes().decode(), aln[i]))
    return path
