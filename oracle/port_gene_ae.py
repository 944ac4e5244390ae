"""CPU restatement of phaser_gene_ae.py (feature-level haplotypic counts) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product
(phaser_b200/) never does.  Pinned against runs of the UNMODIFIED reference script
(tests/golden/gene_ae/*, made by tests/golden/make_golden.py with a stub `intervaltree`).
Each step cites phaser_gene_ae/phaser_gene_ae.py.  Text is parsed directly instead of through pandas: columns
keep the meaning pandas gives them for files the phaser writers produce (ints, floats, strings).
"""
import math


def _zero_divide(a, b):            # :223-227
    return float('inf') if b == 0 else float(a) / float(b)


def _zero_log(value, base):        # :229-233
    return float('-inf') if value == 0 else math.log(value, base)


def variant_feature_reads(row, fstart, fstop, id_separator):
    """:172-219"""
    xvars = row["variants"].split(",")
    if id_separator not in xvars[0] or xvars[0].count(id_separator) < 3:
        raise SystemExit("ERROR - ID separator not found in variant ID, please ensure that --id_separator is set correctly.")
    a, b, used = [], [], []
    for xvar in xvars:
        idx = xvars.index(xvar)
        pos = int(xvar.split(id_separator)[1])
        if (pos - 1) - fstart >= 0 and (pos - 1) - fstop <= 0:
            used.append(xvar)
            if len(xvars) == 1:
                a += list(map(str, range(0, int(row["aCount"]))))
                b += list(map(str, range(0, int(row["bCount"]))))
            else:
                a += row["aReads"].split(";")[idx].split(",")
                b += row["bReads"].split(";")[idx].split(",")
    a = set(a); b = set(b)
    a.discard(""); b.discard("")
    return {"variants": used, "aCount": len(a), "bCount": len(b), "totalCount": len(a) + len(b)}


def parse_rows(text):
    lines = text.split("\n")
    cols = lines[0].split("\t")
    rows = []
    for ln in lines[1:]:
        if ln == "":
            continue
        f = ln.split("\t")
        f += [""] * (len(cols) - len(f))
        r = dict(zip(cols, f))
        for k in ("start", "stop", "aCount", "bCount", "totalCount"):
            r[k] = int(r[k])
        r["gwStat"] = float(r["gwStat"]) if r["gwStat"] != "" else float("nan")
        if "max_haplo_maf" in r:
            r["max_haplo_maf"] = float(r["max_haplo_maf"]) if r["max_haplo_maf"] != "" else float("nan")
        rows.append(r)
    return cols, rows


def run(hc_text, features_text, id_separator="_", gw_cutoff=0.9, min_cov=0, min_haplo_maf=0.0):
    feats = []                                                              # :40-56
    for ln in features_text.split("\n"):
        if ln.strip() == "":
            continue
        c = ln.rstrip().split("\t")
        feats.append(dict(chr=c[0], start=int(c[1]), stop=int(c[2]), name=c[3]))
    by_chrom = {}
    for i, f in enumerate(feats):
        by_chrom.setdefault(f["chr"], []).append(i)
    # interval index per contig (what intervaltree gives the reference): features by start + running maximum of stops
    import bisect
    index = {}
    for c, ids in by_chrom.items():
        ids = sorted(ids, key=lambda i: (feats[i]["start"], i))
        starts = [feats[i]["start"] for i in ids]
        mx = []; cur = -1
        for i in ids:
            cur = max(cur, feats[i]["stop"]); mx.append(cur)
        index[c] = (ids, starts, mx)

    def overlapping(chrom, lo, hi):
        ids, starts, mx = index[chrom]
        j1 = bisect.bisect_left(starts, hi)                 # features with start < hi
        j0 = bisect.bisect_right(mx, lo, 0, j1)             # first position whose running max of stops exceeds lo
        return sorted(ids[j] for j in range(j0, j1) if feats[ids[j]]["stop"] > lo)
    cols, rows = parse_rows(hc_text)
    if "bam" not in cols:
        raise SystemExit("ERROR - this version of phaser_gene_ae is only compatible with results from phASER v1.0.0+")
    out = ["\t".join(["contig", "start", "stop", "name", "aCount", "bCount", "totalCount", "log2_aFC", "n_variants", "variants",
                      "gw_phased", "bam"]) + "\n"]
    bams = list(dict.fromkeys(r["bam"] for r in rows))                        # a set in the reference (:87): block order differs
    for xbam in bams:
        st = [dict(a=0, b=0, variants=[], ua=0, ub=0, uvariants="") for _ in feats]      # :93-100
        for row in rows:
            if row["bam"] != xbam:
                continue
            if row["totalCount"] > 0 and row["contig"] in by_chrom:          # :105
                lo, hi = row["start"] - 1, row["stop"]
                for fi in (overlapping(row["contig"], lo, hi) if lo < hi else []):      # IntervalTree[lo:hi]
                    f = feats[fi]
                    m = variant_feature_reads(row, f["start"], f["stop"], id_separator)
                    s = st[fi]
                    if row["blockGWPhase"] != "0/1" and float(row["gwStat"] >= gw_cutoff):      # :113
                        if min_haplo_maf > 0 and "max_haplo_maf" in cols and row["max_haplo_maf"] < min_haplo_maf:   # :115-122
                            if m["totalCount"] > s["ua"] + s["ub"]:
                                s["ua"], s["ub"], s["uvariants"] = m["aCount"], m["bCount"], m["variants"]
                            continue
                        if row["blockGWPhase"] == "0|1":
                            s["a"] += m["aCount"]; s["b"] += m["bCount"]
                        elif row["blockGWPhase"] == "1|0":
                            s["a"] += m["bCount"]; s["b"] += m["aCount"]
                        s["variants"] += m["variants"]
                    else:                                                     # :135-140
                        if m["totalCount"] > s["ua"] + s["ub"]:
                            s["ua"], s["ub"], s["uvariants"] = m["aCount"], m["bCount"], m["variants"]
        for fi, f in enumerate(feats):                                        # :147-165
            s = st[fi]
            if s["a"] + s["b"] >= s["ua"] + s["ub"]:
                total = s["a"] + s["b"]
                if total >= min_cov:
                    out.append("\t".join(map(str, [f["chr"], f["start"], f["stop"], f["name"], s["a"], s["b"], total,
                                                   _zero_log(_zero_divide(s["a"], s["b"]), 2), len(s["variants"]),
                                                   ",".join(s["variants"]), 1, xbam])) + "\n")
            else:
                total = s["ua"] + s["ub"]
                if total >= min_cov:
                    out.append("\t".join(map(str, [f["chr"], f["start"], f["stop"], f["name"], s["ua"], s["ub"], total,
                                                   _zero_log(_zero_divide(s["ua"], s["ub"]), 2), len(s["uvariants"]),
                                                   ",".join(s["uvariants"]), 0, xbam])) + "\n")
    return "".join(out)


def canon(text):
    """bam blocks come out in set order in the reference (:87): compare header + per-bam blocks"""
    lines = [l for l in text.split("\n") if l != ""]
    blocks = {}
    for ln in lines[1:]:
        blocks.setdefault(ln.split("\t")[-1], []).append(ln)
    return lines[0], blocks
