"""CPU restatement ("port") of the reference read -> variant -> haplotype path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this file; nothing under phaser_b200/ does.  It is the checker,
never the thing shipped or measured as the product.

It restates, on the packed arrays of phaser_b200/layout.py and with integer ids instead of
strings, what the reference computes:
  mapper      read_variant_map.py:25-123 (join), 165-234 (split_read), 236-258 (identify_allele)
  tuples      phaser.py:545-553 (AS cutoff), 1287-1328 (process_mapping_result), 558-581 (merge, Q9)
  noise       phaser.py:610-632
  edges       phaser.py:1265-1285, 660-678, 1594-1654, 686-726
  blocks      phaser.py:1861-1887, 1985-1998
  phasing     phaser.py:2107-2324
  outputs     phaser.py:737-749, 865-1239, 1661-1845
Pinned against the UNMODIFIED reference run in the build container (oracle/harness) on seeded
synthetic inputs and the hand-written quirk vectors; the resulting fixtures are committed under
tests/golden/ (tests/golden/make_golden.py).  The reference itself has no tests or golden vectors
(SURVEY.md section 4), so parity is pinned by reference *runs*, not by reference-owned KATs.

Where the reference's output depends on CPython set iteration order (SURVEY.md Q12, Q13, Q29 and
the row order of variant_connections.txt) the port emits a canonical order; oracle/compare.py
compares those fields canonically.
"""
import bisect
import itertools
import math
from collections import OrderedDict

import numpy as np

BASES = "=ACMGRSVTWYHKDBN"
AS_MISSING = -32768
CLS_OTHER = 2


# ------------------------------------------------------------------------------------- mapper

def _unpack(seq, lo, hi):
    return [BASES[(seq[i >> 1] >> 4) if (i & 1) == 0 else (seq[i >> 1] & 15)] for i in range(lo, hi)]


def split_record(batch, r, baseq):
    """read_variant_map.py:165-234 with --splice 1.  Returns [(genome_start, pseudo_read, insertions)]."""
    lo = int(batch.seq_off[r]); hi = int(batch.seq_off[r + 1])
    bases = _unpack(batch.seq, lo, hi)
    q = batch.qual
    for j in range(hi - lo):
        if int(q[lo + j]) < baseq:           # read_variant_map.py:179-184
            bases[j] = "N"
    segs = []
    read_pos = 0; genome_start = 0; genome_pos = 0
    pseudo = []; ins = {}
    for k in range(int(batch.cigar_off[r]), int(batch.cigar_off[r + 1])):
        n = int(batch.cigar[k]) >> 4; op = int(batch.cigar[k]) & 15
        if op in (0, 7, 8):                  # M = X
            pseudo.extend(bases[read_pos:read_pos + n]); read_pos += n; genome_pos += n
        elif op == 3:                        # N closes the segment
            segs.append((genome_start, pseudo, ins))
            genome_pos += n; genome_start = genome_pos; pseudo = []; ins = {}
        elif op == 2:                        # D
            pseudo.extend("D" * n); genome_pos += n
        elif op == 1:                        # I, keyed by WHOLE-read reference offset (Q3)
            ins[genome_pos - 1] = "".join(bases[read_pos:read_pos + n]); read_pos += n
        elif op == 4:                        # S
            read_pos += n
        # H, P: nothing
    segs.append((genome_start, pseudo, ins))
    return segs


def identify_allele(seg, read_pos, vpos, ref_len):
    """read_variant_map.py:236-258"""
    gstart, pseudo, ins = seg
    st = vpos - (read_pos + gstart)
    en = st + ref_len
    if st >= 0 and en <= len(pseudo):
        s = "".join(pseudo[st:en])
        offset = 0
        for gp in range(st, en):             # segment-relative lookup into whole-read keys (Q3)
            if gp in ins:
                at = (gp - st) + offset + 1
                s = s[:at] + ins[gp] + s[at:]
                offset += len(ins[gp])
        s = s.replace("D", "")
        if s != "N":
            return s
    return ""


def map_reads(batch, vt, baseq, isize_cutoff):
    """Canonical tuple list of one BAM: [(rec, seg, var, allele_string, AS)] in (record, segment,
    variant VCF order).  SURVEY.md 'K1 spec': the streaming buffer of read_variant_map.py:38-49,
    89-112 has no observable effect on coordinate-sorted input."""
    out = []
    for c in range(batch.n_contigs):
        v0 = int(vt.contig_var_off[c]); v1 = int(vt.contig_var_off[c + 1])
        vpos = vt.pos[v0:v1].tolist()
        for r in range(int(batch.contig_rec_off[c]), int(batch.contig_rec_off[c + 1])):
            if not (isize_cutoff == 0 or abs(int(batch.tlen[r])) <= isize_cutoff):   # read_variant_map.py:51
                continue
            rp = int(batch.pos[r])
            a = int(batch.aln_score[r])
            for si, seg in enumerate(split_record(batch, r, baseq)):
                lo = bisect.bisect_left(vpos, rp + seg[0])
                hi = bisect.bisect_right(vpos, rp + seg[0] + len(seg[1]))
                for j in range(lo, hi):
                    s = identify_allele(seg, rp, vpos[j], int(vt.ref_len[v0 + j]))
                    if s != "":
                        out.append((r, si, v0 + j, s, a))
    return out


def tuples_tsv(batch, vt, tuples):
    """The mapper's TSV (read_variant_map.py:117)."""
    names = batch.qnames
    lines = []
    for r, _, v, s, a in tuples:
        lines.append("\t".join([names[int(batch.frag[r])], vt.ids[v], vt.rsids[v], s,
                                "" if a == AS_MISSING else str(a), vt.gt[v], vt.maf[v]]) + "\n")
    return "".join(lines)


# ------------------------------------------------------------------------------------- helpers

class Params:
    def __init__(self, **kw):
        self.baseq = 10
        self.isize = [0.0]
        self.as_q_cutoff = 0.05
        self.cc_threshold = 0.01
        self.max_block_size = 15
        self.unphased_vars = 1
        self.gw_phase_method = 0
        self.gw_phase_vcf = 0
        self.gw_phase_vcf_min_confidence = 0.90
        self.haplo_count_bam_exclude = []     # 0-based
        self.bam_names = ["bam0"]
        self.id_separator = "_"
        self.unique_ids = 0
        self.output_read_ids = 0
        self.read_names = None                # QNAME per fragment id, needed with output_read_ids = 1
        self.output_network = ""              # variant id whose block is dumped as a network
        self.pass_only = 1
        self.remove_dups = 1
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


def variant_info(vt, v):
    """generate_variant_dict, phaser.py:1418-1462 (text metadata part)."""
    alls = vt.all_alleles[v]
    g = list(vt.gt[v])
    phased = "|" in g
    if phased:
        g.remove("|")
    if "/" in g:
        g.remove("/")
    ind = [alls[i] for i in range(len(alls)) if str(i) in g]
    phase = [alls[int(i)] for i in g] if phased else ["-", "-"]
    try:
        maf = float(vt.maf[v])
    except ValueError:
        maf = 0
    rsid = vt.rsids[v] if vt.rsids[v] not in (".", "") else vt.ids[v]
    return dict(id=vt.ids[v], rsid=rsid, ref=alls[0], alleles=ind, phase=phase, maf=maf)


def fmt(x):
    return str(x)


# ------------------------------------------------------------------------------------- the path

class PortResult:
    pass


# Quirk mutants (tests/test_quirk_coverage.py): naming a quirk here makes the port follow the behaviour the reference's
# author presumably INTENDED instead of what the reference does.  A fixture produced by the unmodified reference pins a
# quirk exactly when the mutant no longer reproduces it; the set is empty in every other use of this module.
MUTANTS = set()


def run(vt, batches, params, binom_cdf=None):
    """Whole path.  `batches`: one ReadBatch per BAM (fragment ids share one namespace)."""
    if binom_cdf is None:
        from scipy.stats import binom
        binom_cdf = binom.cdf
    P = params
    res = PortResult()
    nb = len(batches)
    isz = list(P.isize) * nb if len(P.isize) == 1 else list(P.isize)
    contig_of = np.zeros(vt.n_variants, np.int64)
    for c in range(len(vt.contigs)):
        contig_of[int(vt.contig_var_off[c]):int(vt.contig_var_off[c + 1])] = c
    info = {}

    def vinfo(v):
        if v not in info:
            info[v] = variant_info(vt, v)
        return info[v]

    # ---- #2 mapping, AS cutoff, per-variant read lists (phaser.py:526-591)
    var = OrderedDict()        # v -> dict(reads=[[],[]], other=[], haplo=[{bam:[]},{bam:[]}])
    read_vars = OrderedDict()  # contig -> OrderedDict(frag -> [v,...])
    res.tuples = []; res.as_cutoff = []; res.total_tuples = 0
    for b, batch in enumerate(batches):
        tup = map_reads(batch, vt, P.baseq, isz[b])
        res.tuples.append(tup)
        cutoff = None
        if P.as_q_cutoff > 0:
            scores = [t[4] for t in tup if t[4] != AS_MISSING]
            if scores:
                cutoff = np.percentile(scores, P.as_q_cutoff * 100)            # phaser.py:551
        res.as_cutoff.append(cutoff)
        per_contig_rv = OrderedDict()
        last_variant = None
        for (r, si, v, s, a) in tup:
            if cutoff is not None:
                if a == AS_MISSING:
                    raise ValueError("record without AS tag while an alignment-score cutoff is active")
                if not (a >= cutoff):
                    continue
            res.total_tuples += 1
            f = int(batch.frag[r])
            c = int(contig_of[v])
            if c not in per_contig_rv:
                per_contig_rv[c] = OrderedDict()
            vi = vinfo(v)
            if v not in var:
                var[v] = dict(reads=[[], []], other=[], haplo=[OrderedDict(), OrderedDict()])
            e = var[v]
            if s in vi["alleles"]:
                ai = vi["alleles"].index(s)
                per_contig_rv[c].setdefault(f, []).append(v)
                e["reads"][ai].append(f)
                if b not in P.haplo_count_bam_exclude or len(P.haplo_count_bam_exclude) == 0:
                    e["haplo"][ai].setdefault(b, []).append(f)
            else:
                e["other"].append(f)
        for c, rv in per_contig_rv.items():
            if c not in read_vars:
                read_vars[c] = OrderedDict()
            for f, lst in rv.items():
                if "Q9" in MUTANTS and f in read_vars[c]:
                    read_vars[c][f] = read_vars[c][f] + lst      # what phaser.py:578-581 was meant to do
                    continue
                read_vars[c][f] = lst          # always overwritten: stale-variable test, phaser.py:578 (Q9)
    res.var = var

    # ---- noise (phaser.py:610-632)
    match = mism = 0
    for v, e in var.items():
        mis = len(e["other"]); mat = len(e["reads"][0]) + len(e["reads"][1])
        if mat > 0 and (float(mis) / float(mis + mat)) < 0.05:
            match += mat; mism += mis
    if match == 0:
        raise RuntimeError("No reads could be matched to variants.")
    noise_e = float(mism) / (float(match + mism) * 2)
    res.noise_e = noise_e; res.match = match; res.mismatch = mism
    sets = {v: (set(e["reads"][0]), set(e["reads"][1]), set(e["other"])) for v, e in var.items()}

    # ---- connectivity (phaser.py:1265-1285, 660-662)
    overlap = OrderedDict()      # contig -> OrderedDict(v -> set(w))
    for c, rv in read_vars.items():
        for f, lst in rv.items():
            for v in lst:
                for w in lst:
                    if w != v:
                        overlap.setdefault(c, OrderedDict()).setdefault(v, set()).add(w)

    # ---- edge tests (phaser.py:667-680, 1594-1654)
    p_success = 1 - ((6 * noise_e) + (10 * math.pow(noise_e, 2)))
    edges = []                   # (c, a, b, p, c_sup, c_tot, phase_concordant, chosen)
    seen = set()
    for c in overlap:
        for a in overlap[c]:
            for b in sorted(overlap[c][a]):
                if (a, b) in seen or (b, a) in seen:
                    continue
                seen.add((a, b))
                A = sets[a]; B = sets[b]
                cis = len(A[0] & B[0]) + len(A[1] & B[1])
                trans = len(A[1] & B[0]) + len(A[0] & B[1])
                other = len(A[2] & B[0]) + len(A[2] & B[1]) + len(A[0] & B[2]) + len(A[1] & B[2]) + len(A[2] & B[2])
                pc = "."
                ia = vinfo(a); ib = vinfo(b)
                if "-" not in ia["phase"] and "-" not in ib["phase"]:
                    if cis > trans:
                        pc = 1 if ia["phase"].index(ia["alleles"][0]) == ib["phase"].index(ib["alleles"][0]) else 0
                    elif cis < trans:
                        pc = 1 if ia["phase"].index(ia["alleles"][1]) == ib["phase"].index(ib["alleles"][0]) else 0
                sup = max(cis, trans); tot = cis + trans + other
                chosen = 0 if cis > trans else (1 if cis < trans else -1)
                if sup == 0:
                    p = 0
                elif tot - sup > 0:
                    p = binom_cdf(sup, tot, p_success)
                else:
                    p = 1
                edges.append((c, a, b, p, sup, tot, pc, chosen))
    res.edges = edges

    # ---- drop + allele graph (phaser.py:686-726)
    links = OrderedDict()        # (v, allele) -> set((w, allele))
    res.dropped = 0
    conn_lines = ["variant_a\tvariant_b\tsupporting_connections\ttotal_connections\tconflicting_configuration_p\tphase_concordant\n"]
    for (c, a, b, p, sup, tot, pc, chosen) in edges:
        conn_lines.append("\t".join(map(str, [vt.ids[a], vt.ids[b], sup, tot, p, pc])) + "\n")
        if p < P.cc_threshold:
            overlap[c][a].remove(b); overlap[c][b].remove(a)
            if len(overlap[c][a]) == 0:
                del overlap[c][a]
            if len(overlap[c][b]) == 0:
                del overlap[c][b]
            res.dropped += 1
        else:
            for k in ((a, 0), (a, 1), (b, 0), (b, 1)):
                if k not in links:
                    links[k] = set()
            if chosen == -1 and "Q16" in MUTANTS:        # a tie edge that does not glue components
                overlap[c][a].remove(b); overlap[c][b].remove(a)
                if len(overlap[c][a]) == 0:
                    del overlap[c][a]
                if len(overlap[c][b]) == 0:
                    del overlap[c][b]
            if chosen == 0:
                links[(a, 0)].add((b, 0)); links[(b, 0)].add((a, 0)); links[(a, 1)].add((b, 1)); links[(b, 1)].add((a, 1))
            elif chosen == 1:
                links[(a, 0)].add((b, 1)); links[(b, 0)].add((a, 1)); links[(a, 1)].add((b, 0)); links[(b, 1)].add((a, 0))
    res.variant_connections = "".join(conn_lines)

    # ---- allelic counts (phaser.py:737-749)
    lines = ["contig\tposition\tvariantID\trefAllele\taltAllele\trefCount\taltCount\ttotalCount\n"]
    for v in var:
        r0 = len(sets[v][0]); r1 = len(sets[v][1])
        if r0 + r1 > 0:
            vi = vinfo(v)
            lines.append("\t".join([vt.contigs[contig_of[v]], str(int(vt.pos[v])), vt.ids[v], vi["alleles"][0],
                                    vi["alleles"][1], str(r0), str(r1), str(r0 + r1) + "\n"]))
    res.allelic_counts = "".join(lines)

    # ---- prune (phaser.py:756-774)
    if P.unphased_vars == 0:
        rm = [v for v in var if contig_of[v] not in overlap or v not in overlap[contig_of[v]]]
    else:
        rm = [v for v in var if len(var[v]["reads"][0]) + len(var[v]["reads"][1]) == 0]
    for v in rm:
        del var[v]

    # ---- #4 blocks = connected components, listed by first remaining key (phaser.py:1861-1882)
    blocks = []
    for c in overlap:
        pool = OrderedDict((v, set(s)) for v, s in overlap[c].items())
        remaining = set(pool.keys())
        while pool:
            seed_var = next(iter(pool))
            hap = set([seed_var]) | pool[seed_var]
            del pool[seed_var]; remaining.remove(seed_var)
            ov = hap & remaining
            while ov:
                for v in ov:
                    hap |= pool[v]; del pool[v]; remaining.remove(v)
                ov = hap & remaining
            blocks.append((c, sorted(hap)))       # variant index order == (contig, pos, VCF order)
    res.blocks = blocks

    # ---- #5 phasing (phaser.py:2107-2324)
    final = []
    for c, members in blocks:
        vconn = OrderedDict((v, overlap[c][v]) for v in members if v in overlap[c])
        for blk in phase_v3(members, vconn, links, P.max_block_size):
            if blk:
                final.append(blk)
    res.final_blocks = final

    # ---- #6 outputs
    _outputs(res, vt, var, sets, links, final, contig_of, vinfo, P, nb)
    res.info = info
    return res


# ------------------------------------------------------------------------------------- phasing

def _reach(seed, links, allowed=None):
    """build_haplotype_v3 closure (phaser.py:1985-1998) from `seed` over allele links."""
    got = {seed}
    stack = [seed]
    while stack:
        k = stack.pop()
        for n in links.get(k, ()):
            if allowed is not None and n[0] not in allowed:
                continue
            if n not in got:
                got.add(n); stack.append(n)
    return got


def resolve_phase(variants, links, clean=False):
    """phaser.py:2172-2207.  The seed is the first allele key in block order, i.e. variants[0]:0."""
    allowed = set(variants) if clean else None
    got = _reach((variants[0], 0), links, allowed)
    if len(got) == len(variants):
        out = ""
        for v in variants:
            if (v, 0) in got:
                out += "0"
            elif (v, 1) in got:
                out += "1"
        return [[out, inverse_config(out)]]
    return None


def inverse_config(cfg):
    return "".join("-" if a == "-" else str(1 - int(a)) for a in cfg)


def sub_block_phase(variants, links, sub_block_configs=(), attempt_resolve=False):
    """phaser.py:2209-2258"""
    if len(sub_block_configs) > 0:
        A, B = sub_block_configs
        configurations = [A[0] + B[0], A[0] + B[1], A[1] + B[0], A[1] + B[1]]
    else:
        if attempt_resolve:
            x = resolve_phase(variants, links, clean=True)
            if x is not None:
                return x[0]
        configurations = ["".join(s) for s in itertools.product("01", repeat=len(variants))]
        STATS["enumerations"] += 1
    support = OrderedDict()
    for cfg in configurations:
        inv = inverse_config(cfg)
        if cfg + "|" + inv not in support and inv + "|" + cfg not in support:
            s = 0
            for v, a in zip(variants, cfg):
                if a != "-" and (v, int(a)) in links:
                    lk = links[(v, int(a))]
                    for w, b in zip(variants, cfg):
                        if w != v and b != "-" and (w, int(b)) in lk:
                            s += 1
            support[cfg + "|" + inv] = s
    best = max(support.values())
    winners = [k for k, s in support.items() if s == best]
    if len(winners) == 1:
        return winners[0].split("|")
    return ["-" * len(variants), "-" * len(variants)]


def find_weak_points(variants, vconn):
    """phaser.py:2309-2324: edges crossing the cut left of index p, p in [2, n-2]."""
    idx = {v: i for i, v in enumerate(variants)}
    counts = OrderedDict()
    for p in range(2, len(variants) - 1):
        n = 0
        for x in vconn:
            for y in vconn[x]:
                if idx[x] < p - 0.5 and idx[y] > p - 0.5:
                    n += 1
        counts[p] = n
    return counts


def split_by_weak(variants, vconn, max_size):
    """phaser.py:2271-2307"""
    weak = find_weak_points(variants, vconn)
    frags = []
    points = []
    split_at = 1
    max_frag = len(variants)
    while max_frag > max_size or split_at == 1:
        for p in sorted(weak.keys()):
            if weak[p] == split_at and p + 1 not in points and p - 1 not in points:
                points.append(p)
        if points:
            sp = sorted(points)
            frags = [variants[:sp[0]]] + [variants[sp[i - 1]:sp[i]] for i in range(1, len(sp))] + [variants[sp[-1]:]]
        else:
            frags = [variants]
        max_frag = max(len(x) for x in frags)
        split_at += 1
        if split_at > 4 * len(variants) * len(variants) + 8:
            raise RuntimeError("split_by_weak cannot reach max_block_size (the reference would loop forever)")
    return frags


STATS = {"hard_blocks": 0, "enumerations": 0, "merge_fail": 0, "quirk_short": 0}


def phase_v3(variants, vconn, links, max_block_size):
    """phaser.py:2107-2170.  Returns [[(variant, allele_char), ...], ...]."""
    x = resolve_phase(variants, links)
    if x is not None:
        final_blocks = x
        if len(x[0][0]) != len(variants):
            STATS["quirk_short"] += 1
    else:
        STATS["hard_blocks"] += 1
        xmax = len(variants) if max_block_size == 0 else max_block_size
        subs = split_by_weak(variants, vconn, xmax)
        if len(subs) == 1:
            phases = [sub_block_phase(s, links) for s in subs]
        else:
            phases = [sub_block_phase(s, links, attempt_resolve=True) for s in subs]
        split_phases = []
        final_phase = phases[0]
        split_start = 0
        for i in range(1, len(phases)):
            step = [final_phase, phases[i]]
            used = math.ceil(sum(sum(len(y) for y in x) for x in step) / 2)
            new_phase = sub_block_phase(variants[split_start:split_start + used], links, step)
            if "-" in new_phase[0]:
                STATS["merge_fail"] += 1
                split_phases.append(final_phase)
                if "Q14" in MUTANTS:
                    split_start = split_start + len(final_phase[0])      # the offset of the sub-block that starts now
                else:
                    split_start = used                      # Q14: not an offset sum
                final_phase = phases[i]
            else:
                final_phase = new_phase
        final_blocks = split_phases + [final_phase]
    out = []
    vi = 0
    for blk in final_blocks:
        ob = []
        for a in blk[0]:
            ob.append((variants[vi], a))
            vi += 1
        if "-" not in ob[0][1]:
            out.append(ob)
    return out


# ------------------------------------------------------------------------------------- outputs

def _lts(xs, sep=","):
    return sep.join(str(x) for x in xs)


def _relabel_first_occurrence(lists):
    """Canonical form of the aReads/bReads column: fragment -> index of its
    first occurrence walking variants in order, reads in list order."""
    ids = {}
    out = []
    for lst in lists:
        cur = []
        for f in lst:
            if f not in ids:
                ids[f] = len(ids)
            cur.append(ids[f])
        out.append(_lts(cur))
    return _lts(out, ";")


def _outputs(res, vt, var, sets, links, final, contig_of, vinfo, P, nb):
    hc = ["\t".join(["contig", "start", "stop", "variants", "variantCount", "variantsBlacklisted",
                     "variantCountBlacklisted", "haplotypeA", "haplotypeB", "aCount", "bCount", "totalCount",
                     "blockGWPhase", "gwStat", "max_haplo_maf", "bam", "aReads", "bReads"] +
                    (["read_ids_a", "read_ids_b"] if P.output_read_ids == 1 else [])) + "\n"]     # phaser.py:837-838

    def read_ids(lists):
        # phaser.py:1087-1088 prints list(set(...)); the canonical form is first-occurrence order (oracle/compare.py sorts)
        seen = {}
        for lst in lists:
            for f in lst:
                seen.setdefault(f, None)
        return _lts(P.read_names[f] for f in seen)
    hp = ["\t".join(['contig', 'start', 'stop', 'length', 'variants', 'variant_ids', 'variant_alleles', 'reads_hap_a',
                     'reads_hap_b', 'reads_total', 'edges_supporting', 'edges_total', 'annotated_phase',
                     'phase_concordant', 'gw_phase', 'gw_confidence']) + "\n"]
    ac = ["\t".join(['variant_a', 'rsid_a', 'variant_b', 'rsid_b', 'configuration']) + "\n"]
    lookup = {}          # v -> (members, "a|b", block_index)
    gw_stat_of = {}      # block_index -> stat
    gw_phase = {}        # v -> [phase of allele0, phase of allele1] after correction
    all_variants = []
    nan = float("nan")
    block_index = 0
    res.block_rows = []
    for blk in final:
        block_index += 1
        variants = sorted(v for v, _ in blk)
        all_variants += variants
        hap_a = "".join(a for _, a in blk)
        hap_b = "".join(str(int(not int(x))) for x in hap_a)
        sup = tot = 0
        for (v, a) in blk:
            for (w, b) in blk:
                if (v, a) != (w, b):
                    lk = links[(v, int(a))]
                    if (w, int(b)) in lk:
                        sup += 1
                    if (w, 0) in lk:
                        tot += 1
                    if (w, 1) in lk:
                        tot += 1
        sup = sup / 2; tot = tot / 2
        vis = [vinfo(v) for v in variants]
        rsids = [vi["rsid"] for vi in vis] if P.unique_ids == 0 else [vt.ids[v] for v in variants]
        positions = [int(vt.pos[v]) for v in variants]
        chrom = vt.contigs[contig_of[variants[0]]]
        alleles = [[], []]; phases = [[], []]; hap_counts = [0, 0]
        for h in (0, 1):
            hx = (hap_a, hap_b)[h]
            u = set()
            for i, v in enumerate(variants):
                al = vis[i]["alleles"][int(hx[i])]
                alleles[h].append(al)
                try:
                    phases[h].append(vis[i]["phase"].index(al))
                except ValueError:
                    phases[h].append(nan)
                u |= sets[v][vis[i]["alleles"].index(al)]
            hap_counts[h] = len(u)
        use_phases = [x for x in phases[0] if str(x) != "nan"]
        phase_concordant = 1 if len(set(use_phases)) <= 1 else 0            # Q22: also 1 when no phase is known
        if "Q22" in MUTANTS:
            phase_concordant = 1 if len(set(use_phases)) == 1 else 0
        ps = ["".join(str(x).replace("nan", "-") for x in phases[h]) for h in (0, 1)]
        nan_strip = [int(x) for x in phases[0] if x >= 0]
        corrected = [phases[0], phases[1]]
        stat = 0.5
        mafs = [vi["maf"] for vi in vis]
        if len(nan_strip) > 0:
            # set() of the phase list: every float('nan') object is distinct (Q23)
            n_nan = sum(1 for x in phases[0] if x != x)
            distinct = len(set(x for x in phases[0] if x == x)) + n_nan
            if "Q23" in MUTANTS:
                distinct = len(set(x for x in phases[0] if x == x))          # nan-aware "all known phases agree"
            if distinct == 1:
                stat = 1
            elif P.gw_phase_method == 0:
                stat = sum(nan_strip) / len(nan_strip)
                if stat < 0.5:
                    corrected = [[0] * len(variants), [1] * len(variants)]
                elif stat > 0.5:
                    corrected = [[1] * len(variants), [0] * len(variants)]
                stat = max([stat, 1 - stat])
            elif P.gw_phase_method == 1:
                support = [0, 0]
                for ph, maf in zip(phases[0], mafs):
                    if ph == 0:
                        support[0] += maf
                    elif ph == 1:
                        support[1] += maf
                if sum(support) > 0:
                    stat = max(support) / sum(support)
                    if support[0] > support[1]:
                        corrected = [[0] * len(variants), [1] * len(variants)]
                    elif support[1] > support[0]:
                        corrected = [[1] * len(variants), [0] * len(variants)]
                else:
                    stat = sum(nan_strip) / len(nan_strip)
                    if stat < 0.5:
                        corrected = [[0] * len(variants), [1] * len(variants)]
                    elif stat > 0.5:
                        corrected = [[1] * len(variants), [0] * len(variants)]
                    stat = max([stat, 1 - stat])
        if "Q28" in MUTANTS:
            stat = float(stat)                # one print format for the statistic
        gw_stat_of[block_index] = stat
        max_maf = max(mafs)
        for i, v in enumerate(variants):
            lookup[v] = (variants, hap_a[i] + "|" + hap_b[i], block_index, max_maf)
            ai = vis[i]["alleles"].index(alleles[0][i])
            g = [None, None]
            g[ai] = corrected[0][i]; g[1 - ai] = corrected[1][i]
            gw_phase[v] = g
        cps = ["".join(str(x).replace("nan", "-") for x in corrected[h]) for h in (0, 1)]
        hp.append("\t".join(map(str, [chrom, min(positions), max(positions), max(positions) - min(positions),
                                      len(variants), _lts(rsids), _lts(alleles[0]) + "|" + _lts(alleles[1]),
                                      hap_counts[0], hap_counts[1], sum(hap_counts), sup, tot, ps[0] + "|" + ps[1],
                                      phase_concordant, cps[0] + "|" + cps[1], stat])) + "\n")
        res.block_rows.append(dict(variants=variants, hap_a=hap_a, counts=hap_counts, sup=sup, tot=tot, stat=stat))
        black = getattr(vt, "haplo_blacklisted", None)
        used = [i for i, v in enumerate(variants) if black is None or not black[v]]      # phaser.py:1070
        blacklisted = [vt.ids[v] for i, v in enumerate(variants) if i not in used]
        for b in range(nb):
            if b in P.haplo_count_bam_exclude:
                continue
            cnt = [0, 0]; vreads = [[], []]
            for h in (0, 1):
                hx = (hap_a, hap_b)[h]
                u = set()
                for i in used:
                    v = variants[i]
                    ai = vis[i]["alleles"].index(vis[i]["alleles"][int(hx[i])])
                    lst = var[v]["haplo"][ai].get(b, [])
                    vreads[h].append(lst)
                    u |= set(lst)
                cnt[h] = len(u)
            gwp = "0/1"
            if corrected[0][0] == 0:
                gwp = "0|1"
            elif corrected[0][0] == 1:
                gwp = "1|0"
            if sum(cnt) > 0:
                hc.append("\t".join(map(str, [chrom, min(positions), max(positions), _lts(vt.ids[variants[i]] for i in used),
                                              len(used), _lts(sorted(blacklisted)), len(blacklisted),
                                              _lts(alleles[0][i] for i in used), _lts(alleles[1][i] for i in used), cnt[0], cnt[1],
                                              sum(cnt), gwp, stat] +
                                        ([read_ids(vreads[0]), read_ids(vreads[1])] if P.output_read_ids == 1 else []) +   # phaser.py:1120-1121
                                        [str(max_maf), P.bam_names[b],
                                              _relabel_first_occurrence(vreads[0]),
                                              _relabel_first_occurrence(vreads[1])])) + "\n")
        if P.output_network != "" and P.output_network in [vt.ids[v] for v in variants]:
            # generate_hap_network_all (phaser.py:1928-1949) + writers (phaser.py:1128-1157)
            counted = set(); junctions = []
            for i in range(len(variants)):
                for j in range(len(variants)):
                    if i == j:
                        continue
                    for a in (0, 1):
                        for oa in (0, 1):
                            if (i, a, j, oa) in counted or (j, oa, i, a) in counted:
                                continue
                            k = len(sets[variants[i]][a] & sets[variants[j]][oa])
                            junctions.append([vt.ids[variants[i]] + ":" + vis[i]["alleles"][a],
                                              vt.ids[variants[j]] + ":" + vis[j]["alleles"][oa], k, 0])
                            junctions.append([vt.ids[variants[i]] + ":" + vis[i]["alleles"][1 - a],
                                              vt.ids[variants[j]] + ":" + vis[j]["alleles"][1 - oa], k, 1])
                            counted.add((i, a, j, oa))
            lk = ["variantA\tvariantB\tconnections\tinferred\n"]; nodes = []
            for item in junctions:
                if item[2] > 0:
                    lk.append(_lts(item, "\t") + "\n"); nodes += [item[0], item[1]]
            nd = ["id\tindex\tassigned_hap\n"]
            for item in dict.fromkeys(nodes):          # a set in the reference: compared as a set
                xv, xa = item.split(":")[0], item.split(":")[1]
                vi_ = [vt.ids[v] for v in variants].index(xv)
                nd.append(item + "\t" + str(vi_) + "\t" + ("A" if alleles[0][vi_] == xa else "B") + "\n")
            res.network_links = "".join(lk); res.network_nodes = "".join(nd)
        for i, va in enumerate(variants):
            for j, vb in enumerate(variants):
                if va != vb:
                    ra = vis[i]["ref"] == alleles[0][i]; rb = vis[j]["ref"] == alleles[1][j]
                    cfg = "trans" if ra == rb else "cis"
                    ac.append("\t".join([vt.ids[va], vis[i]["rsid"], vt.ids[vb], vis[j]["rsid"], cfg]) + "\n")
    res.singletons = []
    if P.unphased_vars == 1:
        phased = set(all_variants)
        singles = [v for v in var if v not in phased]
        res.singletons = singles
        for v in singles:
            vi = vinfo(v)
            if getattr(vt, "haplo_blacklisted", None) is not None and vt.haplo_blacklisted[v]:
                continue                                   # phaser.py:1189
            for b in range(nb):
                if b in P.haplo_count_bam_exclude:
                    continue
                ca = len(set(var[v]["haplo"][0].get(b, []))); cb = len(set(var[v]["haplo"][1].get(b, [])))
                if ca + cb > 0:
                    if "-" not in vi["phase"]:
                        pstr = str(vi["phase"].index(vi["alleles"][0])) + "|" + str(vi["phase"].index(vi["alleles"][1]))
                    else:
                        pstr = "0/1"
                    hc.append("\t".join([vt.contigs[contig_of[v]], str(int(vt.pos[v])), str(int(vt.pos[v])), vt.ids[v],
                                         "1", "", "0", vi["alleles"][0], vi["alleles"][1], str(ca), str(cb), str(ca + cb),
                                         pstr, "1"] +
                                        ([read_ids([var[v]["haplo"][h].get(b, [])]) for h in (0, 1)] if P.output_read_ids == 1 else []) +   # phaser.py:1216-1217
                                        [str(vi["maf"]), P.bam_names[b], "", ""]) + "\n")
        for v in singles:
            vi = vinfo(v)
            if "-" not in vi["phase"]:
                pstr = str(vi["phase"].index(vi["alleles"][0])) + "|" + str(vi["phase"].index(vi["alleles"][1]))
            else:
                pstr = "-|-"
            name = vi["rsid"] if P.unique_ids == 0 else vt.ids[v]
            n0 = len(sets[v][0]); n1 = len(sets[v][1])
            hp.append("\t".join([vt.contigs[contig_of[v]], str(int(vt.pos[v]) - 1), str(int(vt.pos[v])), "1", "1", name,
                                 vi["alleles"][0] + "|" + vi["alleles"][1], str(n0), str(n1), str(n0 + n1), "0", "0",
                                 pstr, "nan", pstr, "nan"]) + "\n")
    res.haplotypic_counts = "".join(hc)
    res.haplotypes = "".join(hp)
    res.allele_config = "".join(ac)
    res.all_variants = all_variants
    res.lookup = lookup
    res.gw_stat_of = gw_stat_of
    res.gw_phase = gw_phase


def write_vcf_text(res, vt, vcf_lines, sample_column, P, chrom_of_interest=""):
    """write_vcf, phaser.py:1661-1845, on the already cut (cols 1-9 + sample) text lines.
    Returns (text, unphased_phased, phase_corrections)."""
    id_to_v = {s: i for i, s in enumerate(vt.ids)}
    out = []
    fmt_text = ""
    corrections = unphased_phased = 0
    tags = ['PG', 'PB', 'PI', 'PW', 'PC', 'PM']
    for line in vcf_lines:
        cols = line.replace("\n", "").split("\t")
        cols = cols[0:9] + ([cols[sample_column]] if len(cols) > sample_column else [])
        line = "\t".join(cols) + "\n"
        if "##FORMAT" in line:
            fmt_text += line
            out.append(line)
        elif line.startswith("#CHROM"):
            for t, d in (("PG", "phASER Local Genotype"), ("PB", "phASER Local Block"),
                         ("PI", "phASER Local Block Index (unique for each block)"),
                         ("PM", "phASER Local Block Maximum Variant MAF"), ("PW", "phASER Genome Wide Genotype"),
                         ("PC", "phASER Genome Wide Confidence")):
                if "##FORMAT=<ID=%s," % t not in fmt_text:
                    out.append('##FORMAT=<ID=%s,Number=1,Type=String,Description="%s">\n' % (t, d))
            if P.gw_phase_vcf == 2 and "##FORMAT=<ID=PS," not in fmt_text:
                out.append('##FORMAT=<ID=PS,Number=1,Type=String,Description="Phase Set">\n')
            out.append("\t".join(cols[0:9] + [cols[9]]) + "\n")
        elif line[0:1] == "#":
            out.append(line)
        else:
            chrom = cols[0]; pos = int(cols[1])
            if chrom_of_interest == "" or chrom == chrom_of_interest:
                if "GT" in cols[8]:
                    gt_index = cols[8].split(":").index("GT")
                    genotype = list(cols[9].split(":")[gt_index])
                    if "|" in genotype:
                        genotype.remove("|")
                    if "/" in genotype:
                        genotype.remove("/")
                    all_alleles = [cols[3]] + cols[4].split(",")
                    n_fields = len(cols[8].split(":"))
                    for i in range(9, len(cols)):
                        sf = len(cols[i].split(":"))
                        if sf != n_fields:
                            cols[i] += ":" * (n_fields - sf)
                    ff = cols[8].split(":")
                    for t in tags:
                        if t not in ff:
                            ff.append(t)
                    cols[8] = ":".join(ff)
                    uid = chrom + P.id_separator + str(pos) + P.id_separator + P.id_separator.join(all_alleles)
                    v = id_to_v.get(uid)
                    if v is not None and v in res.lookup:
                        members, ab, bidx, max_maf = res.lookup[v]
                        vi = res.info[v]
                        alleles_out = []; gw_out = ["", ""]
                        for al in ab.split("|"):
                            base = vi["alleles"][int(al)]
                            vidx = all_alleles.index(base)
                            g = res.gw_phase[v][int(al)]
                            if isinstance(g, int):
                                gw_out[g] = str(vidx)
                            alleles_out.append(str(vidx))
                        names = [res.info[m]["rsid"].replace(":", "_") for m in members]
                        stat = res.gw_stat_of[bidx]
                        if "-" not in gw_out:
                            xf = cols[9].split(":")
                            new_phase = "|".join(gw_out)
                            if stat >= P.gw_phase_vcf_min_confidence:
                                if "|" in xf[gt_index] and xf[gt_index] != new_phase:
                                    corrections += 1
                                if "/" in xf[gt_index] and xf[gt_index] != "./." and xf[gt_index] != new_phase:
                                    unphased_phased += 1
                                if P.gw_phase_vcf in (1, 2):
                                    xf[gt_index] = new_phase
                                    cols[9] = ":".join(xf)
                            if P.gw_phase_vcf == 2 and stat < P.gw_phase_vcf_min_confidence:
                                xf[gt_index] = "|".join(alleles_out)
                                cols[9] = ":".join(xf)
                        sf = cols[9].split(":")
                        sf += [''] * (len(ff) - len(sf))
                        sf[ff.index('PG')] = "|".join(alleles_out)
                        sf[ff.index('PB')] = _lts(names)
                        sf[ff.index('PI')] = str(bidx)
                        sf[ff.index('PM')] = str(max_maf)
                        sf[ff.index('PW')] = "|".join(gw_out)
                        sf[ff.index('PC')] = str(stat)
                        if P.gw_phase_vcf == 2 and stat < P.gw_phase_vcf_min_confidence:
                            if 'PS' not in ff:
                                cols[8] += ":PS"; ff.append("PS"); sf.append('')
                            sf[ff.index('PS')] = str(bidx)
                        cols[9] = ":".join(sf)
                    else:
                        sf = cols[9].split(":")
                        sf += [''] * (len(ff) - len(sf))
                        sf[ff.index('PG')] = "/".join(sorted(genotype))
                        sf[ff.index('PB')] = '.'; sf[ff.index('PI')] = '.'; sf[ff.index('PM')] = '.'
                        sf[ff.index('PW')] = cols[9].split(":")[gt_index]
                        sf[ff.index('PC')] = '.'
                        cols[9] = ":".join(sf)
                out.append("\t".join(cols[0:9] + [cols[9]]) + "\n")
    return "".join(out), unphased_phased, corrections
