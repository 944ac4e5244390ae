"""Text outputs from the device result arrays: the five .txt files and the annotated VCF.

Formats, column orders and number formatting follow the reference's writers:
  allelic_counts.txt       phaser/phaser.py:737-749
  variant_connections.txt  phaser/phaser.py:683-695
  haplotypes.txt           phaser/phaser.py:843, 865-1043, 1224-1239
  haplotypic_counts.txt    phaser/phaser.py:836-839, 1048-1125, 1180-1221
  allele_config.txt        phaser/phaser.py:847, 1160-1172
  VCF                      phaser/phaser.py:1661-1845
Fields whose order in the reference comes from CPython set iteration are written in a canonical
order instead: singleton rows in first-seen order, variant_connections rows
in edge-table order, aReads/bReads as first-occurrence indices.
"""
from collections import OrderedDict

import numpy as np

from .pipeline import PhaseResult, edge_pvalues

NONE32 = 0xFFFFFFFF


class PhzUnsupported(Exception):
    """the native writer met an input it leaves to the general Python writer"""
NAN = float("nan")


class VariantMeta:
    """generate_variant_dict (phaser.py:1418-1462): text-side view of one het site."""
    __slots__ = ("id", "rsid", "ref", "alleles", "phase", "maf", "chrom", "pos")

    def __init__(self, vt, v, chrom):
        alls = vt.all_alleles[v]
        g = list(vt.gt[v])
        phased = "|" in g
        if phased:
            g.remove("|")
        if "/" in g:
            g.remove("/")
        self.alleles = [alls[i] for i in range(len(alls)) if str(i) in g]
        self.phase = [alls[int(i)] for i in g] if phased else ["-", "-"]
        try:
            self.maf = float(vt.maf[v])
        except ValueError:
            self.maf = 0
        self.id = vt.ids[v]
        self.rsid = vt.rsids[v] if vt.rsids[v] not in (".", "") else vt.ids[v]
        self.ref = alls[0]
        self.chrom = chrom
        self.pos = int(vt.pos[v])

    @classmethod
    def from_site_text(cls, row, chrom, sep, pos):
        """the same record from one line of phz_vcf_site_text (POS, ID, REF, ALT, the genotype's two alleles in index
        order, the two alleles in genotype order or "-" twice); None when the native reader left the site alone ("?")"""
        f = row.split("\t")
        if len(f) != 8:
            return None
        m = cls.__new__(cls)
        m.id = chrom + sep + f[0] + sep + f[2] + sep + (f[3].replace(",", sep) if "," in f[3] else f[3])
        m.rsid = f[1] if f[1] not in (".", "") else m.id
        m.ref = f[2]
        m.alleles = [f[4], f[5]]
        m.phase = [f[6], f[7]]
        m.maf = 0                     # gw_phase_method 0: no allele frequency is read (float("None") -> 0 in __init__)
        m.chrom = chrom
        m.pos = pos
        return m


class Outputs:
    def __init__(self, res: PhaseResult, vt, bam_names, params, unphased_vars=1, gw_phase_method=0, unique_ids=0,
                 read_names=None, output_network="", lib=None, threads=0):
        """`read_names` (QNAME per fragment id) switches the --output_read_ids 1 columns on (phaser.py:837-838);
        `output_network` names the variant whose block is dumped as a network (phaser.py:1128-1157)."""
        self.output_network = output_network
        self.lib = lib; self.threads = int(threads or 0)         # native formatter of the aReads / bReads columns (optional)
        self.network = None        # (links text, nodes text) once block_tables() met the block
        self.res = res; self.vt = vt; self.bam_names = bam_names; self.P = params
        self.read_names = read_names
        self.unphased_vars = unphased_vars; self.gw_phase_method = gw_phase_method; self.unique_ids = unique_ids
        self._meta = {}
        self._site_rows = None; self._site_sep = "_"
        self.contig_of = np.zeros(vt.n_variants, np.int64)
        for c in range(len(vt.contigs)):
            self.contig_of[int(vt.contig_var_off[c]):int(vt.contig_var_off[c + 1])] = c
        self.lookup = {}         # v -> (members, "a|b", block_index, max_maf)
        self.block_stats = []; self.block_mafs = []          # per final block, in output order (native VCF writer)
        self.gw_stat_of = {}
        self.gw_phase = {}
        self.all_variants = []

    def meta(self, v) -> VariantMeta:
        m = self._meta.get(v)
        if m is None:
            chrom = self.vt.contigs[self.contig_of[v]]
            row = self._site_rows.get(v) if self._site_rows is not None else None
            if row is not None:
                m = VariantMeta.from_site_text(row, chrom, self._site_sep, int(self.vt.pos[v]))
            if m is None:
                m = VariantMeta(self.vt, v, chrom)
            self._meta[v] = m
        return m

    def prefetch_meta(self):
        """Every variant the tables will name (edge ends, block members, covered sites), read from the VCF text in one
        native call when the variant table is the native one."""
        rows = getattr(getattr(self.vt.ids, "o", None), "prefetch_sites", None)
        if rows is None:
            return
        r = self.res; owner = self.vt.ids.o
        sz = r.setsize.reshape(-1, 3)
        need = [np.asarray(r.ed_a, np.int64), np.asarray(r.ed_b, np.int64), np.nonzero(sz[:, 0] + sz[:, 1] > 0)[0].astype(np.int64)]
        if "members" in r.arrays:
            mem = np.asarray(r.members, np.int64)
            need.append(mem[(mem >= 0) & (mem < self.vt.n_variants)])
        owner.prefetch_sites(np.unique(np.concatenate(need)))
        self._site_rows = owner.site_rows; self._site_sep = owner.sep

    # ------------------------------------------------------------------ simple tables
    def first_seen_order(self):
        vf = self.res.vfirst
        seen = np.nonzero(vf != np.iinfo(vf.dtype).max)[0]     # u32 tuple index, or u64 merged key (shard.py)
        return seen[np.argsort(vf[seen], kind="stable")]

    def allelic_counts(self):
        r = self.res
        out = ["contig\tposition\tvariantID\trefAllele\taltAllele\trefCount\taltCount\ttotalCount\n"]
        sz = r.setsize.reshape(-1, 3)
        n = 0
        for v in self.first_seen_order().tolist():
            a, b = int(sz[v, 0]), int(sz[v, 1])
            if a + b > 0:
                m = self.meta(v)
                out.append("\t".join([m.chrom, str(m.pos), m.id, m.alleles[0], m.alleles[1], str(a), str(b), str(a + b) + "\n"]))
                n += 1
        self.covered_count = n
        return "".join(out)

    def variant_connections(self):
        r = self.res
        out = ["variant_a\tvariant_b\tsupporting_connections\ttotal_connections\tconflicting_configuration_p\tphase_concordant\n"]
        pv = edge_pvalues(r.ed_sup, r.ed_tot, r.noise_e)
        meta = self.meta; add = out.append
        # (pv: python ints 0 / 1 or numpy.float64, printed by str() as the reference prints them -- Q28)
        for a, b, sup, tot, cfg, p in zip(r.ed_a.tolist(), r.ed_b.tolist(), r.ed_sup.tolist(), r.ed_tot.tolist(),
                                         r.ed_cfg.tolist(), pv.tolist()):
            ma = meta(a); mb = meta(b)
            pc = "."
            pa = ma.phase; pb = mb.phase
            if "-" not in pa and "-" not in pb:       # phaser.py:1609-1620
                if cfg == 0:
                    pc = 1 if pa.index(ma.alleles[0]) == pb.index(mb.alleles[0]) else 0
                elif cfg == 1:
                    pc = 1 if pa.index(ma.alleles[1]) == pb.index(mb.alleles[0]) else 0
            add("%s\t%s\t%d\t%d\t%s\t%s\n" % (ma.id, mb.id, sup, tot, p, pc))
        return "".join(out)

    # ------------------------------------------------------------------ blocks
    def _read_list_rows(self):
        """(final block, bam, hap) -> list (one per variant, in variant order) of fragment lists"""
        r = self.res
        rows = {}
        if "rl_row" not in r.arrays:
            return rows
        row = r.rl_row; var = r.rl_var; frag = r.rl_frag
        n = row.shape[0]
        if n == 0:
            return rows
        brk = np.nonzero((row[1:] != row[:-1]) | (var[1:] != var[:-1]))[0] + 1
        starts = np.concatenate([[0], brk]); ends = np.concatenate([brk, [n]])
        for s, e in zip(starts.tolist(), ends.tolist()):
            rows.setdefault(int(row[s]), OrderedDict())[int(var[s])] = frag[s:e].tolist()
        return rows

    def _format_read_lists_native(self, keys, offs, variants):
        """aReads / bReads columns of the requested rows through the native formatter (phz_format_read_lists)"""
        import ctypes
        import os
        r = self.res
        row = np.ascontiguousarray(r.rl_row, np.uint32); var = np.ascontiguousarray(r.rl_var, np.uint32)
        frag = np.ascontiguousarray(r.rl_frag, np.uint32)
        k = np.asarray(keys, np.uint32); o = np.asarray(offs, np.int64); v = np.asarray(variants if variants else [0], np.uint32)
        text = ctypes.c_void_p(); toff = ctypes.c_void_p()
        rc = self.lib.phz_format_read_lists(int(row.shape[0]), row.ctypes.data, var.ctypes.data, frag.ctypes.data, int(k.shape[0]),
                                            k.ctypes.data, o.ctypes.data, v.ctypes.data, self.threads or (os.cpu_count() or 1),
                                            ctypes.byref(text), ctypes.byref(toff))
        if rc != 0:
            raise RuntimeError(self.lib.phz_last_error().decode())
        n = int(k.shape[0])
        off = np.ctypeslib.as_array(ctypes.cast(toff, ctypes.POINTER(ctypes.c_int64)), shape=(n + 1,))
        total = int(off[n])
        buf = ctypes.string_at(text.value, total) if total else b""
        s = buf.decode()
        ol = off.tolist()
        return [s[ol[i]:ol[i + 1]] for i in range(n)]

    def _read_id_columns(self, frag_lists):
        """read_ids_a / read_ids_b of one row: the distinct QNAMEs of `frag_lists` in first-occurrence order, i.e.
        the list the aReads/bReads indices point into (phaser.py:1087-1105; the reference's order is list(set()))"""
        seen = {}
        for lst in frag_lists:
            for f in lst:
                if f not in seen:
                    seen[f] = len(seen)
        return ",".join(self.read_names[f] for f in seen)

    def _network_tables(self, variants, ms, alleles_a):
        """generate_hap_network_all + the two writers (phaser.py:1928-1949, 1128-1157): junction = number of reads shared by
        two allele read sets (all BAMs).  Node rows come from a set in the reference; here in first-mention order."""
        r = self.res
        if "g_var" not in r.arrays:
            raise RuntimeError("--output_network needs the kept tuples (PhaseParams.want_kept_tuples)")
        pos_of = {v: i for i, v in enumerate(variants)}
        sets = [[set(), set()] for _ in variants]
        sel = np.nonzero(np.isin(r.g_var, np.asarray(variants, r.g_var.dtype)) & ((r.g_cb & 3) < 2))[0]
        for v, cb, f in zip(r.g_var[sel].tolist(), r.g_cb[sel].tolist(), r.g_frag[sel].tolist()):
            sets[pos_of[v]][cb & 3].add(f)
        links = ["\t".join(["variantA", "variantB", "connections", "inferred\n"])]
        nodes = []
        n = len(variants)
        for i in range(n):
            for j in range(n):
                if j <= i:
                    continue          # (j, oa, i, a) was counted when the roles were swapped (phaser.py:1943)
                for a in (0, 1):
                    for oa in (0, 1):
                        k = len(sets[i][a] & sets[j][oa])
                        if k > 0:
                            for x, y, inf in ((a, oa, 0), (1 - a, 1 - oa, 1)):
                                na = ms[i].id + ":" + ms[i].alleles[x]; nb_ = ms[j].id + ":" + ms[j].alleles[y]
                                links.append("\t".join([na, nb_, str(k), str(inf)]) + "\n")
                                nodes += [na, nb_]
        out = ["id\tindex\tassigned_hap\n"]
        for item in dict.fromkeys(nodes):
            xvar, xallele = item.split(":")[0], item.split(":")[1]
            vi = [m.id for m in ms].index(xvar)
            out.append(item + "\t" + str(vi) + "\t" + ("A" if alleles_a[vi] == xallele else "B") + "\n")
        return "".join(links), "".join(out)

    def _singleton_read_sets(self):
        """(variant, bam, allele) -> fragment ids in tuple order, for the variants outside every block"""
        r = self.res
        out = {}
        if "sg_var" not in r.arrays:
            return out
        for v, cb, f in zip(r.sg_var.tolist(), r.sg_cb.tolist(), r.sg_frag.tolist()):
            out.setdefault((v, cb >> 2, cb & 3), []).append(f)
        return out

    @staticmethod
    def _relabel(lists):
        ids = {}
        out = []
        for lst in lists:
            cur = []
            for f in lst:
                i = ids.get(f)
                if i is None:
                    i = len(ids); ids[f] = i
                cur.append(str(i))
            out.append(",".join(cur))
        return ";".join(out)

    def block_tables(self):
        """Returns (haplotypes.txt, haplotypic_counts.txt, allele_config.txt)."""
        r = self.res; nb = r.n_bams
        excl = set(self.P.haplo_count_bam_exclude)
        hc = ["\t".join(["contig", "start", "stop", "variants", "variantCount", "variantsBlacklisted",
                         "variantCountBlacklisted", "haplotypeA", "haplotypeB", "aCount", "bCount", "totalCount",
                         "blockGWPhase", "gwStat", "max_haplo_maf", "bam", "aReads", "bReads"] +
                        (["read_ids_a", "read_ids_b"] if self.read_names is not None else [])) + "\n"]
        hp = ["\t".join(['contig', 'start', 'stop', 'length', 'variants', 'variant_ids', 'variant_alleles',
                         'reads_hap_a', 'reads_hap_b', 'reads_total', 'edges_supporting', 'edges_total',
                         'annotated_phase', 'phase_concordant', 'gw_phase', 'gw_confidence']) + "\n"]
        ac = ["\t".join(['variant_a', 'rsid_a', 'variant_b', 'rsid_b', 'configuration']) + "\n"]
        native_rl = self.lib is not None and self.read_names is None and "rl_row" in r.arrays
        rl = {} if native_rl else self._read_list_rows()
        pending = []; req_keys = []; req_off = [0]; req_vars = []          # rows whose read columns the native formatter fills
        bb = max(1, int(np.ceil(np.log2(max(nb, 2)))))
        fcnt = r.fb_cnt.reshape(-1, 2).tolist()
        fbc = (r.fb_bcnt.reshape(-1, nb, 2) if r.fb_bcnt.size else r.fb_bcnt.reshape(0, nb, 2)).tolist()
        fb_first = r.fb_first.tolist(); fb_len = r.fb_len.tolist(); fb_sup = r.fb_sup.tolist(); fb_tot = r.fb_tot.tolist()
        members = r.members; v_hap = r.v_hap; meta = self.meta
        black = getattr(self.vt, "haplo_blacklisted", None)
        if black is not None and not np.any(black):
            black = None                                             # nothing blacklisted: every variant is used
        for f in range(len(fb_first)):
            block_index = f + 1
            o = fb_first[f]; n = fb_len[f]
            mem = members[o:o + n]
            variants = mem.tolist()
            self.all_variants += variants
            hap_a = v_hap[mem].tolist()
            ms = [meta(v) for v in variants]
            sup = fb_sup[f] * 2 / 2; tot = fb_tot[f] * 2 / 2          # floats, phaser.py:894-895
            rsids = [m.rsid for m in ms] if self.unique_ids == 0 else [m.id for m in ms]
            positions = [m.pos for m in ms]
            alleles = [[], []]; phases = [[], []]
            for h in (0, 1):
                for i, m in enumerate(ms):
                    al = m.alleles[hap_a[i] ^ h]
                    alleles[h].append(al)
                    try:
                        phases[h].append(m.phase.index(al))
                    except ValueError:
                        phases[h].append(NAN)
            known = [x for x in phases[0] if x == x]
            phase_concordant = 1 if len(set(known)) <= 1 else 0
            ps = ["".join("-" if x != x else str(x) for x in phases[h]) for h in (0, 1)]
            corrected = [phases[0], phases[1]]
            stat = 0.5
            mafs = [m.maf for m in ms]
            if len(known) > 0:                                        # phaser.py:959-1025
                n_nan = len(phases[0]) - len(known)
                if len(set(known)) + n_nan == 1:                     # every float('nan') is its own set element (Q23)
                    stat = 1
                else:
                    use_mean = self.gw_phase_method == 0
                    if self.gw_phase_method == 1:
                        support = [0, 0]
                        for ph, maf in zip(phases[0], mafs):
                            if ph == 0:
                                support[0] += maf
                            elif ph == 1:
                                support[1] += maf
                        if sum(support) > 0:
                            stat = max(support) / sum(support)
                            if support[0] > support[1]:
                                corrected = [[0] * n, [1] * n]
                            elif support[1] > support[0]:
                                corrected = [[1] * n, [0] * n]
                        else:
                            use_mean = True
                    if use_mean:
                        stat = sum(known) / len(known)               # numpy.mean of small ints, phaser.py:970
                        if stat < 0.5:
                            corrected = [[0] * n, [1] * n]
                        elif stat > 0.5:
                            corrected = [[1] * n, [0] * n]
                        stat = max([stat, 1 - stat])
            max_maf = max(mafs)
            self.gw_stat_of[block_index] = stat
            self.block_stats.append(stat); self.block_mafs.append(max_maf)
            for i, v in enumerate(variants):
                self.lookup[v] = (variants, "%d|%d" % (hap_a[i], 1 - hap_a[i]), block_index, max_maf)
                ai = ms[i].alleles.index(alleles[0][i])
                g = [None, None]
                g[ai] = corrected[0][i]; g[1 - ai] = corrected[1][i]
                self.gw_phase[v] = g
            cps = ["".join("-" if x != x else str(x) for x in corrected[h]) for h in (0, 1)]
            ca, cb = fcnt[f]
            hp.append("\t".join(map(str, [ms[0].chrom, min(positions), max(positions), max(positions) - min(positions),
                                          n, ",".join(rsids), ",".join(alleles[0]) + "|" + ",".join(alleles[1]),
                                          ca, cb, ca + cb, sup, tot, ps[0] + "|" + ps[1], phase_concordant,
                                          cps[0] + "|" + cps[1], stat])) + "\n")
            gwp = "0/1"
            if corrected[0][0] == 0:
                gwp = "0|1"
            elif corrected[0][0] == 1:
                gwp = "1|0"
            if black is None:
                used = list(range(n)); blacklisted = []
            else:
                used = [i for i, v in enumerate(variants) if not black[v]]        # phaser.py:1070
                blacklisted = sorted(ms[i].id for i in range(n) if i not in used)
            for b in range(nb):
                if b in excl:
                    continue
                a_cnt, b_cnt = fbc[f][b]
                if a_cnt + b_cnt > 0:
                    cols = []; ids = []
                    for h in (0, 1):
                        key = (((f << bb) | b) << 1) | h
                        if native_rl:
                            req_keys.append(key); req_vars += [variants[i] for i in used]; req_off.append(len(req_vars))
                            continue
                        per_var = rl.get(key, {})
                        lists = [per_var.get(variants[i], []) for i in used]
                        cols.append(self._relabel(lists))
                        if self.read_names is not None:       # the row's own order: ids BEFORE maf/bam (phaser.py:1120-1123)
                            ids.append(self._read_id_columns(lists))
                    head = "\t".join(map(str, [ms[0].chrom, min(positions), max(positions),
                                               ",".join(ms[i].id for i in used), len(used), ",".join(blacklisted),
                                               len(blacklisted), ",".join(alleles[0][i] for i in used),
                                               ",".join(alleles[1][i] for i in used), a_cnt, b_cnt, a_cnt + b_cnt, gwp, stat] +
                                          ids + [str(max_maf), self.bam_names[b]]))
                    if native_rl:
                        pending.append((len(hc), head, len(req_keys) - 2)); hc.append(None)
                    else:
                        hc.append(head + "\t" + cols[0] + "\t" + cols[1] + "\n")
            if self.output_network != "" and self.output_network in [m.id for m in ms]:
                self.network = self._network_tables(variants, ms, alleles[0])
            al0 = alleles[0]; al1 = alleles[1]
            for i, ma in enumerate(ms):
                left = ma.id + "\t" + ma.rsid + "\t"; ra = ma.ref == al0[i]
                for j, mb in enumerate(ms):
                    if i != j:
                        ac.append(left + mb.id + "\t" + mb.rsid + ("\ttrans\n" if ra == (mb.ref == al1[j]) else "\tcis\n"))
        if pending:
            texts = self._format_read_lists_native(req_keys, req_off, req_vars)
            for at, head, k in pending:
                hc[at] = head + "\t" + texts[k] + "\t" + texts[k + 1] + "\n"
        # ---- singletons (phaser.py:1180-1239)
        if self.unphased_vars == 1:
            ncls = r.ncls.reshape(-1, 3); sz = r.setsize.reshape(-1, 3)
            vbc = r.vb_cnt.reshape(-1, nb, 2)
            fso = self.first_seen_order()
            singles = fso[((ncls[fso, 0].astype(np.int64) + ncls[fso, 1]) > 0) & (r.v_final[fso] == NONE32)].tolist()
            black = getattr(self.vt, "haplo_blacklisted", None)
            sg = self._singleton_read_sets() if self.read_names is not None else None
            for v in singles:
                m = self.meta(v)
                if black is not None and black[v]:
                    continue                                     # phaser.py:1189
                for b in range(nb):
                    if b in excl:
                        continue
                    ca, cb = int(vbc[v, b, 0]), int(vbc[v, b, 1])
                    if ca + cb > 0:
                        if "-" not in m.phase:
                            pstr = str(m.phase.index(m.alleles[0])) + "|" + str(m.phase.index(m.alleles[1]))
                        else:
                            pstr = "0/1"
                        ids = [] if sg is None else [self._read_id_columns([sg.get((v, b, h), [])]) for h in (0, 1)]
                        hc.append("\t".join([m.chrom, str(m.pos), str(m.pos), m.id, "1", "", "0", m.alleles[0], m.alleles[1],
                                             str(ca), str(cb), str(ca + cb), pstr, "1"] + ids +
                                            [str(m.maf), self.bam_names[b], "", ""]) + "\n")
            for v in singles:
                m = self.meta(v)
                if "-" not in m.phase:
                    pstr = str(m.phase.index(m.alleles[0])) + "|" + str(m.phase.index(m.alleles[1]))
                else:
                    pstr = "-|-"
                name = m.rsid if self.unique_ids == 0 else m.id
                n0, n1 = int(sz[v, 0]), int(sz[v, 1])
                hp.append("\t".join([m.chrom, str(m.pos - 1), str(m.pos), "1", "1", name, m.alleles[0] + "|" + m.alleles[1],
                                     str(n0), str(n1), str(n0 + n1), "0", "0", pstr, "nan", pstr, "nan"]) + "\n")
        return "".join(hp), "".join(hc), "".join(ac)

    # ------------------------------------------------------------------ VCF
    def vcf_native(self, nv, gw_phase_vcf=0, min_conf=0.90, chrom_of_interest="", id_separator="_", chr_prefix=""):
        """write_vcf (phaser.py:1661-1845) through the native writer (include/phz.h: phz_vcf_write) over the text the
        native parser holds.  Must run after block_tables().  Returns (text bytes view, unphased_phased, phase_corrections,
        records) with records = (chromosome index per data line, names, begs, ends) for the index; raises PhzUnsupported when
        the input needs the general Python writer."""
        import ctypes
        from .engine import phz_vcf_annot
        r = self.res; V = self.vt.n_variants
        B = int(r.fb_first.shape[0])
        v_block = np.where(r.v_final == NONE32, -1, r.v_final.astype(np.int64)).astype(np.int32)
        v_hap = np.ascontiguousarray(r.v_hap, np.uint8)
        gw = np.full((V, 2), -1, np.int8)
        for v, g in self.gw_phase.items():
            for k in (0, 1):
                if isinstance(g[k], int):
                    gw[v, k] = g[k]
        first = np.ascontiguousarray(r.fb_first, np.int64); ln = np.ascontiguousarray(r.fb_len, np.int64)
        members = np.ascontiguousarray(r.members.astype(np.int64), np.int32)
        index = np.arange(1, B + 1, dtype=np.int32)
        conf = np.asarray([1 if s >= min_conf else 0 for s in self.block_stats], np.uint8)
        stat_blob = b"".join(str(s).encode() + b"\0" for s in self.block_stats) + b"\0"
        maf_blob = b"".join(str(m).encode() + b"\0" for m in self.block_mafs) + b"\0"
        if len(self.block_stats) != B:
            raise RuntimeError("vcf_native must run after block_tables()")
        a = phz_vcf_annot()
        a.gw_phase_vcf = int(gw_phase_vcf); a.ids_match = 0 if chr_prefix else 1; a.chrom_of_interest = chrom_of_interest.encode()
        a.id_separator = id_separator.encode(); a.chr_prefix = chr_prefix.encode()
        a.n_variants = V; a.v_block = v_block.ctypes.data; a.v_hap = v_hap.ctypes.data; a.v_gw = gw.ctypes.data
        a.n_blocks = B; a.blk_first = first.ctypes.data; a.blk_len = ln.ctypes.data; a.blk_members = members.ctypes.data
        a.blk_index = index.ctypes.data; a.blk_confident = conf.ctypes.data; a.blk_stat = stat_blob; a.blk_maf = maf_blob
        text = ctypes.c_void_p(); n = ctypes.c_int64(0); counts = (ctypes.c_int64 * 2)()
        rc = nv.lib.phz_vcf_write(nv.h, ctypes.byref(a), nv.threads, ctypes.byref(text), ctypes.byref(n), counts)
        if rc != 0:
            msg = nv.lib.phz_last_error().decode()
            raise PhzUnsupported(msg)
        buf = (ctypes.c_char * max(1, n.value)).from_address(text.value) if n.value else b""
        nr = ctypes.c_int64(0); pc = ctypes.c_void_p(); pb = ctypes.c_void_p(); pe = ctypes.c_void_p(); po = ctypes.c_void_p()
        pn = ctypes.c_void_p(); nn = ctypes.c_int32(0)
        nv.lib.phz_vcf_records(nv.h, ctypes.byref(nr), ctypes.byref(pc), ctypes.byref(pb), ctypes.byref(pe), ctypes.byref(po),
                               ctypes.byref(pn), ctypes.byref(nn))

        def arr(ptr, cnt, dt):
            if cnt == 0 or not ptr.value:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(cnt,)).copy()
        names = []; p = pn.value
        for _ in range(nn.value):
            sname = ctypes.string_at(p); names.append(sname.decode()); p += len(sname) + 1
        records = (arr(pc, nr.value, np.int32), names, arr(pb, nr.value, np.int64), arr(pe, nr.value, np.int64))
        return memoryview(buf)[:n.value] if n.value else memoryview(b""), int(counts[0]), int(counts[1]), records

    @staticmethod
    def vcf_save_native(nv, path_vcf_gz, csi=False):
        """bgzip + tabix of the text the last vcf_native() produced (phz_vcf_save): path and path + .tbi / .csi"""
        if nv.lib.phz_vcf_save(nv.h, path_vcf_gz.encode(), int(bool(csi)), nv.threads) != 0:
            raise RuntimeError(nv.lib.phz_last_error().decode())

    def vcf_text(self, vcf_lines, sample_column, id_separator="_", gw_phase_vcf=0, min_conf=0.90, chrom_of_interest=""):
        """write_vcf (phaser.py:1661-1845).  Must run after block_tables().  Returns
        (text, unphased_phased, phase_corrections)."""
        id_to_v = {s: i for i, s in enumerate(self.vt.ids)}
        out = []
        fmt_text = ""
        corrections = unphased_phased = 0
        tags = ['PG', 'PB', 'PI', 'PW', 'PC', 'PM']
        fmt_cache = {}
        rec_chrom, rec_beg, rec_end = [], [], []       # reference span of every data line written (for the index)
        for line in vcf_lines:
            cols = line.replace("\n", "").split("\t", sample_column + 1)      # columns past the sample's are cut away anyway
            cols = cols[0:9] + ([cols[sample_column]] if len(cols) > sample_column else [])
            if line[0:1] != "#" and "##FORMAT" not in line:
                line = "d"            # a data line: none of the header tests below can match the cut line either
            else:
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
                if gw_phase_vcf == 2 and "##FORMAT=<ID=PS," not in fmt_text:
                    out.append('##FORMAT=<ID=PS,Number=1,Type=String,Description="Phase Set">\n')
                out.append("\t".join(cols[0:9] + [cols[9]]) + "\n")
            elif line[0:1] == "#":
                out.append(line)
            else:
                chrom = cols[0]; pos = int(cols[1])
                if chrom_of_interest == "" or chrom == chrom_of_interest:
                    fmt = cols[8]
                    if "GT" in fmt:
                        info = fmt_cache.get(fmt)
                        if info is None:          # FORMAT strings repeat line after line: split / extend them once
                            fields = fmt.split(":")
                            ff0 = list(fields)
                            for t in tags:
                                if t not in ff0:
                                    ff0.append(t)
                            info = (fields.index("GT"), len(fields), ff0, ":".join(ff0), [ff0.index(t) for t in tags])
                            fmt_cache[fmt] = info
                        gt_index, n_fields, ff0, fmt_out, (iPG, iPB, iPI, iPW, iPC, iPM) = info
                        genotype = list(cols[9].split(":")[gt_index])
                        if "|" in genotype:
                            genotype.remove("|")
                        if "/" in genotype:
                            genotype.remove("/")
                        all_alleles = [cols[3]] + cols[4].split(",")
                        for i in range(9, len(cols)):
                            sf = len(cols[i].split(":"))
                            if sf != n_fields:
                                cols[i] += ":" * (n_fields - sf)
                        cols[8] = fmt_out
                        uid = chrom + id_separator + str(pos) + id_separator + id_separator.join(all_alleles)
                        v = id_to_v.get(uid)
                        if v is not None and v in self.lookup:
                            ff = list(ff0)
                            members, ab, bidx, max_maf = self.lookup[v]
                            m = self.meta(v)
                            alleles_out = []; gw_out = ["", ""]
                            for al in ab.split("|"):
                                base = m.alleles[int(al)]
                                vidx = all_alleles.index(base)
                                g = self.gw_phase[v][int(al)]
                                if isinstance(g, int):
                                    gw_out[g] = str(vidx)
                                alleles_out.append(str(vidx))
                            names = [self.meta(x).rsid.replace(":", "_") for x in members]
                            stat = self.gw_stat_of[bidx]
                            if "-" not in gw_out:
                                xf = cols[9].split(":")
                                new_phase = "|".join(gw_out)
                                if stat >= min_conf:
                                    if "|" in xf[gt_index] and xf[gt_index] != new_phase:
                                        corrections += 1
                                    if "/" in xf[gt_index] and xf[gt_index] != "./." and xf[gt_index] != new_phase:
                                        unphased_phased += 1
                                    if gw_phase_vcf in (1, 2):
                                        xf[gt_index] = new_phase
                                        cols[9] = ":".join(xf)
                                if gw_phase_vcf == 2 and stat < min_conf:
                                    xf[gt_index] = "|".join(alleles_out)
                                    cols[9] = ":".join(xf)
                            sf = cols[9].split(":")
                            sf += [''] * (len(ff) - len(sf))
                            sf[iPG] = "|".join(alleles_out)
                            sf[iPB] = ",".join(names)
                            sf[iPI] = str(bidx)
                            sf[iPM] = str(max_maf)
                            sf[iPW] = "|".join(gw_out)
                            sf[iPC] = str(stat)
                            if gw_phase_vcf == 2 and stat < min_conf:
                                if 'PS' not in ff:
                                    cols[8] += ":PS"; ff.append("PS"); sf.append('')
                                sf[ff.index('PS')] = str(bidx)
                            cols[9] = ":".join(sf)
                        else:
                            sf = cols[9].split(":")
                            gt_text = sf[gt_index]
                            sf += [''] * (len(ff0) - len(sf))
                            sf[iPG] = "/".join(sorted(genotype))
                            sf[iPB] = '.'; sf[iPI] = '.'; sf[iPM] = '.'
                            sf[iPW] = gt_text
                            sf[iPC] = '.'
                            cols[9] = ":".join(sf)
                    out.append("\t".join(cols[0:9] + [cols[9]]) + "\n")
                    e = pos - 1 + len(cols[3])
                    if "END=" in cols[7]:
                        for kv in cols[7].split(";"):
                            if kv.startswith("END="):
                                try:
                                    e = max(e, int(kv[4:]))
                                except ValueError:
                                    pass
                    rec_chrom.append(chrom); rec_beg.append(pos - 1); rec_end.append(e)
        self.vcf_records = (rec_chrom, rec_beg, rec_end)
        return "".join(out), unphased_phased, corrections
