#!/usr/bin/env python
"""phaser_gene_ae.py -- drop-in for the reference's feature-level haplotypic counts (phaser_gene_ae/phaser_gene_ae.py).

Same flags, same output columns.  The join of haplotype rows with features and the distinct-read counting
(variant_feature_reads, phaser_gene_ae.py:172-219) run on the GPU through phz_gene_ae_pairs (include/phz.h);
parsing, the order-dependent fold over rows (:103-141) and the text are host work.  Differences a user can see:
no pandas / intervaltree needed; the per-BAM blocks come out in order of first appearance in the input instead of
CPython set order (:87).
"""
import argparse
import math
import os
import sys

import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

VERSION = "1.2.0"


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--haplotypic_counts", required=True)
    p.add_argument("--features", required=True)
    p.add_argument("--o", required=True)
    p.add_argument("--id_separator", default="_")
    p.add_argument("--gw_cutoff", type=float, default=0.9)
    p.add_argument("--min_cov", type=int, default=0)
    p.add_argument("--min_haplo_maf", type=float, default=0)
    p.add_argument("--device", default="cuda:0", help="CUDA device to run on (this implementation only)")
    return p


class Features:
    """BED rows (phaser_gene_ae.py:40-56) + the sorted arrays the device join wants."""

    def __init__(self, text):
        self.chr, self.start, self.stop, self.name = [], [], [], []
        for ln in text.split("\n"):
            if ln.strip() == "":
                continue
            c = ln.rstrip().split("\t")
            if int(c[1]) >= int(c[2]):
                raise ValueError("IntervalTree: Null Interval objects not allowed in IntervalTree: %s" % ln.rstrip())
            self.chr.append(c[0]); self.start.append(int(c[1])); self.stop.append(int(c[2])); self.name.append(c[3])
        self.contigs = list(dict.fromkeys(self.chr))
        cid = {c: i for i, c in enumerate(self.contigs)}
        ci = np.asarray([cid[c] for c in self.chr], np.int64)
        st = np.asarray(self.start, np.int64); sp = np.asarray(self.stop, np.int64)
        self.order = np.lexsort((st, ci)) if len(self.chr) else np.zeros(0, np.int64)      # sorted index -> original index
        self.contig_id = cid
        self.f_start = st[self.order].astype(np.int32); self.f_stop = sp[self.order].astype(np.int32)
        sc = ci[self.order]
        self.f_contig_off = np.searchsorted(sc, np.arange(len(self.contigs) + 1)).astype(np.int64)
        mx = self.f_stop.copy()
        for c in range(len(self.contigs)):                    # running maximum of the stops inside each contig
            a, b = int(self.f_contig_off[c]), int(self.f_contig_off[c + 1])
            if b > a:
                mx[a:b] = np.maximum.accumulate(self.f_stop[a:b])
        self.f_maxstop = mx


class Rows:
    """haplotypic_counts.txt -> arrays (phaser_gene_ae.py:58-79 column meaning).  Read ids are renumbered densely
    per (row, haplotype); the sets of the reference hold the id strings, so equal strings = equal ids."""

    def __init__(self, text, features: Features, id_separator):
        lines = text.split("\n")
        self.cols = lines[0].split("\t")
        if "bam" not in self.cols:
            raise SystemExit("ERROR - this version of phaser_gene_ae is only compatible with results from phASER v1.0.0+")
        ix = {c: i for i, c in enumerate(self.cols)}
        self.has_maf = "max_haplo_maf" in ix
        n = len(self.cols)
        self.contig, self.start, self.stop, self.a, self.b, self.ida, self.idb = [], [], [], [], [], [], []
        self.phase, self.gw, self.maf, self.bam, self.variants = [], [], [], [], []
        var_off = [0]; var_pos = []; id_off = [0]; ids = []
        for ln in lines[1:]:
            if ln == "":
                continue
            f = ln.split("\t")
            f += [""] * (n - len(f))
            xvars = f[ix["variants"]].split(",")
            total = int(f[ix["totalCount"]])
            active = total > 0 and f[ix["contig"]] in features.contig_id
            if active and (id_separator not in xvars[0] or xvars[0].count(id_separator) < 3):      # :181-184
                print("ERROR - ID separator not found in variant ID, please ensure that --id_separator is set correctly.")
                sys.exit(1)
            self.contig.append(features.contig_id[f[ix["contig"]]] if active else -1)
            self.start.append(int(f[ix["start"]])); self.stop.append(int(f[ix["stop"]]))
            self.a.append(int(f[ix["aCount"]])); self.b.append(int(f[ix["bCount"]]))
            self.phase.append(f[ix["blockGWPhase"]])
            self.gw.append(float(f[ix["gwStat"]]) if f[ix["gwStat"]] != "" else float("nan"))
            self.maf.append((float(f[ix["max_haplo_maf"]]) if f[ix["max_haplo_maf"]] != "" else float("nan")) if self.has_maf else 0.0)
            self.bam.append(f[ix["bam"]]); self.variants.append(xvars)
            if not active:
                var_off.append(var_off[-1]); self.ida.append(0); self.idb.append(0)
                continue
            for xv in xvars:
                var_pos.append(int(xv.split(id_separator)[1]))
            nd = [0, 0]
            if len(xvars) > 1:
                lists = []
                for h, col in enumerate(("aReads", "bReads")):
                    parts = f[ix[col]].split(";")
                    seen = {}
                    per_var = []
                    for xv in xvars:
                        k = xvars.index(xv)                          # :187 (first occurrence of the id)
                        cur = []
                        for tok in (parts[k].split(",") if k < len(parts) else [""]):
                            if tok == "":
                                continue                              # :210-211
                            i = seen.get(tok)
                            if i is None:
                                i = len(seen); seen[tok] = i
                            cur.append(i)
                        per_var.append(cur)
                    nd[h] = len(seen)
                    lists.append(per_var)
                for k in range(len(xvars)):
                    ids += lists[0][k]; id_off.append(len(ids)); ids += lists[1][k]; id_off.append(len(ids))
            else:
                id_off += [len(ids), len(ids)]
            self.ida.append(nd[0]); self.idb.append(nd[1])
            var_off.append(len(var_pos))
        self.n = len(self.contig)
        self.var_off = np.asarray(var_off, np.uint32); self.var_pos = np.asarray(var_pos, np.int32)
        self.id_off = np.asarray(id_off, np.uint32); self.ids = np.asarray(ids, np.uint32)


def _zero_divide(a, b):
    return float('inf') if b == 0 else float(a) / float(b)


def _zero_log(value, base):
    return float('-inf') if value == 0 else math.log(value, base)


def run_text(engine, hc_text, features_text, id_separator="_", gw_cutoff=0.9, min_cov=0, min_haplo_maf=0.0):
    """Returns the output file's text."""
    F = Features(features_text)
    R = Rows(hc_text, F, id_separator)
    pr, pf, pa, pb = engine.gene_ae_pairs(R, F)
    pf = F.order[pf.astype(np.int64)]                      # back to the features' file order
    nf = len(F.chr)
    out = ["\t".join(["contig", "start", "stop", "name", "aCount", "bCount", "totalCount", "log2_aFC", "n_variants", "variants",
                      "gw_phased", "bam"]) + "\n"]
    bams = list(dict.fromkeys(R.bam))
    bam_of_row = np.asarray([bams.index(b) for b in R.bam], np.int64) if R.n else np.zeros(0, np.int64)
    for bi, xbam in enumerate(bams):
        a = [0] * nf; b = [0] * nf; ua = [0] * nf; ub = [0] * nf
        variants = [[] for _ in range(nf)]; uvariants = [""] * nf
        sel = np.nonzero(bam_of_row[pr.astype(np.int64)] == bi)[0] if pr.shape[0] else []
        for p in sel.tolist() if len(sel) else []:
            r = int(pr[p]); fi = int(pf[p]); ca = int(pa[p]); cb = int(pb[p])

            def used():                                    # names only where they are printed (:191)
                return [xv for xv in R.variants[r]
                        if (int(xv.split(id_separator)[1]) - 1) - F.start[fi] >= 0 and (int(xv.split(id_separator)[1]) - 1) - F.stop[fi] <= 0]
            if R.phase[r] != "0/1" and float(R.gw[r] >= gw_cutoff):                          # :113
                if min_haplo_maf > 0 and R.has_maf and R.maf[r] < min_haplo_maf:          # :115-122
                    if ca + cb > ua[fi] + ub[fi]:
                        ua[fi], ub[fi], uvariants[fi] = ca, cb, used()
                    continue
                if R.phase[r] == "0|1":
                    a[fi] += ca; b[fi] += cb
                elif R.phase[r] == "1|0":
                    a[fi] += cb; b[fi] += ca
                variants[fi] += used()
            elif ca + cb > ua[fi] + ub[fi]:                                                 # :135-140
                ua[fi], ub[fi], uvariants[fi] = ca, cb, used()
        for fi in range(nf):                                                                # :147-165
            if a[fi] + b[fi] >= ua[fi] + ub[fi]:
                total = a[fi] + b[fi]
                if total >= min_cov:
                    out.append("\t".join(map(str, [F.chr[fi], F.start[fi], F.stop[fi], F.name[fi], a[fi], b[fi], total,
                                                   _zero_log(_zero_divide(a[fi], b[fi]), 2), len(variants[fi]),
                                                   ",".join(variants[fi]), 1, xbam])) + "\n")
            else:
                total = ua[fi] + ub[fi]
                if total >= min_cov:
                    out.append("\t".join(map(str, [F.chr[fi], F.start[fi], F.stop[fi], F.name[fi], ua[fi], ub[fi], total,
                                                   _zero_log(_zero_divide(ua[fi], ub[fi]), 2), len(uvariants[fi]),
                                                   ",".join(uvariants[fi]), 0, xbam])) + "\n")
    return "".join(out)


def run(args, engine=None):
    print("")
    print("##################################################")
    print("          Welcome to phASER Gene AE v%s" % VERSION)
    print("  Author: Stephane Castel (stephanecastel@gmail.com)")
    print("  B200-native join + distinct-read counting (phaser_b200)")
    print("##################################################")
    print("")
    if args.min_haplo_maf < 0 or args.min_haplo_maf > 0.5:
        print("ERROR - invalid value for min_haplo_maf specified. Value must be between 0 and 0.5.")
        sys.exit(1)
    if engine is None:
        from phaser_b200.engine import Engine
        engine = Engine(device=args.device)
    print("#1 Loading features...")
    ft = open(args.features).read()
    print("#2 Loading haplotype counts...")
    hc = open(args.haplotypic_counts).read()
    print("#3 Processing results...")
    text = run_text(engine, hc, ft, args.id_separator, args.gw_cutoff, args.min_cov, args.min_haplo_maf)
    with open(args.o, "w") as f:
        f.write(text)


def main(argv=None):
    run(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
