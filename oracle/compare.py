"""Parity definition between two sets of phASER output files.

TEST INFRASTRUCTURE (see oracle/port.py header).

Byte-equal:       allelic_counts.txt, allele_config.txt, decompressed .vcf.gz,
                  haplotypes.txt block rows (in order), haplotypic_counts.txt block rows cols 1-16 (in order).
Canonical-equal:  haplotypic_counts cols 17-18 after first-occurrence relabelling per row / haplotype
                  (the reference prints indices into list(set(...)), CPython-hash ordered; the only
                  consumer, phaser_gene_ae.py:204-217, treats them as opaque per-row ids);
                  singleton rows of haplotypic_counts.txt and haplotypes.txt as sorted multisets
                  (the reference iterates a set difference, phaser.py:1181-1183);
                  variant_connections.txt as a set of unordered pairs (rows come from set iteration,
                  phaser.py:670-678).
"""


def _relabel(col):
    ids = {}
    out = []
    for part in col.split(";"):
        cur = []
        for tok in part.split(","):
            if tok == "":
                continue
            if tok not in ids:
                ids[tok] = len(ids)
            cur.append(str(ids[tok]))
        out.append(",".join(cur))
    return ";".join(out)


def canon_haplotypic_counts(text):
    lines = text.split("\n")
    header = lines[0]
    with_ids = header.endswith("read_ids_a\tread_ids_b")
    blocks, singles = [], []
    for ln in lines[1:]:
        if ln == "":
            continue
        c = ln.split("\t")
        ids = []
        if with_ids and len(c) >= 20:
            # --output_read_ids 1: rows carry the two id lists BEFORE max_haplo_maf (phaser.py:1120-1123, 1216-1219) although
            # the header names them last (phaser.py:837-838); both are list(set(...)) -> compared as sorted sets
            ids = [",".join(sorted(x.split(","))) for x in c[14:16]]
            c = c[:14] + c[16:]
        if len(c) >= 18:
            c[16] = _relabel(c[16]); c[17] = _relabel(c[17])
        c[5] = ",".join(sorted(c[5].split(","))) if c[5] else c[5]     # variantsBlacklisted is printed from a set (phaser.py:1119)
        row = "\t".join(c + ids)
        # singleton rows: variantCount == 1, nothing blacklisted, empty aReads/bReads (phaser.py:1214-1220)
        if c[4] == "1" and c[6] == "0" and c[16] == "" and c[17] == "" and c[13] == "1" and "," not in c[3]:
            singles.append(row)
        else:
            blocks.append(row)
    return header, blocks, sorted(singles)


def canon_haplotypes(text):
    lines = text.split("\n")
    header = lines[0]
    blocks, singles = [], []
    for ln in lines[1:]:
        if ln == "":
            continue
        c = ln.split("\t")
        if c[4] == "1":
            singles.append(ln)
        else:
            blocks.append(ln)
    return header, blocks, sorted(singles)


def canon_connections(text):
    lines = text.split("\n")
    rows = set()
    n = 0
    for ln in lines[1:]:
        if ln == "":
            continue
        c = ln.split("\t")
        a, b = c[0], c[1]
        # phase_concordant is symmetric in (a, b) (phaser.py:1607-1620 compares index equality)
        key = (min(a, b), max(a, b)) + tuple(c[2:])
        rows.add(key)
        n += 1
    return lines[0], rows, n


def diff_outputs(ref, got, files=("allelic_counts", "allele_config", "haplotypes", "haplotypic_counts",
                                  "variant_connections", "vcf")):
    """`ref` / `got`: dict name -> text.  Returns a list of human-readable mismatch strings (empty = parity)."""
    bad = []

    def first_diff(a, b, what):
        if a == b:
            return
        la, lb = (a.split("\n"), b.split("\n")) if isinstance(a, str) else (a, b)
        for i in range(max(len(la), len(lb))):
            x = la[i] if i < len(la) else "<missing>"
            y = lb[i] if i < len(lb) else "<missing>"
            if x != y:
                bad.append("%s: first difference at row %d (ref %d rows, got %d rows)\n  ref: %s\n  got: %s" % (
                    what, i, len(la), len(lb), x[:400], y[:400]))
                return

    for name in files:
        if name not in ref or name not in got:
            bad.append("%s: missing (ref %s, got %s)" % (name, name in ref, name in got))
            continue
        if name in ("allelic_counts", "allele_config", "vcf"):
            first_diff(ref[name], got[name], name)
        elif name == "haplotypes":
            hr, br, sr = canon_haplotypes(ref[name]); hg, bg, sg = canon_haplotypes(got[name])
            first_diff([hr], [hg], "haplotypes header")
            first_diff(br, bg, "haplotypes block rows")
            first_diff(sr, sg, "haplotypes singleton rows (sorted)")
        elif name == "haplotypic_counts":
            hr, br, sr = canon_haplotypic_counts(ref[name]); hg, bg, sg = canon_haplotypic_counts(got[name])
            first_diff([hr], [hg], "haplotypic_counts header")
            first_diff(br, bg, "haplotypic_counts block rows")
            first_diff(sr, sg, "haplotypic_counts singleton rows (sorted)")
        elif name == "variant_connections":
            hr, rr, nr = canon_connections(ref[name]); hg, rg, ng = canon_connections(got[name])
            if hr != hg:
                bad.append("variant_connections header differs")
            if nr != ng or rr != rg:
                only_r = sorted(rr - rg)[:3]; only_g = sorted(rg - rr)[:3]
                bad.append("variant_connections: ref %d rows, got %d rows; only in ref %s; only in got %s" % (
                    nr, ng, only_r, only_g))
    # --output_network dump (phaser.py:1128-1157): link rows in order; node rows come from set(nodes) -> compared as a set
    if "network_links" in ref or "network_links" in got:
        if ref.get("network_links") != got.get("network_links"):
            first_diff(ref.get("network_links", ""), got.get("network_links", ""), "network.links")
        if sorted(ref.get("network_nodes", "").split("\n")[1:]) != sorted(got.get("network_nodes", "").split("\n")[1:]) or \
                ref.get("network_nodes", "").split("\n")[0] != got.get("network_nodes", "").split("\n")[0]:
            bad.append("network.nodes differ (as sets)")
    return bad
