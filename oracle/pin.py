"""Pin oracle/port.py against the UNMODIFIED reference (build container only; needs /root/reference).

TEST INFRASTRUCTURE.  For each seed: generate a small synthetic sample (phaser_b200/synth.py),
write its SAM-text + VCF twins, run the real reference through oracle/harness, run the port on the
same files (parsed by the product's own host parsers) and diff with oracle/compare.py.

    python -m oracle.pin --seeds 1 2 3 --variants 400 --pairs 4000
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import port, compare                      # noqa: E402
from oracle.harness import run_reference as rr        # noqa: E402


def load_inputs(vcf_gz, sams, sample, mapq, paired_end, remove_dups=1, pass_only=1):
    from phaser_b200 import vcfio, samio
    col = vcfio.sample_column_map(vcf_gz)[sample]
    vt, st = vcfio.parse_vcf(vcf_gz, col, pass_only=pass_only)
    fd = samio.FragmentDictionary()
    mq = [int(x) for x in str(mapq).split(",")]
    pe = [int(x) for x in str(paired_end).split(",")]
    if len(mq) == 1:
        mq = mq * len(sams)
    if len(pe) == 1:
        pe = pe * len(sams)
    batches = [samio.parse_sam(s, vt.contigs, fd, bool(remove_dups), bool(pe[i]), mq[i]) for i, s in enumerate(sams)]
    return vt, st, batches, col


def bam_display_names(paths):
    """phaser.py:469-480"""
    base = [os.path.basename(p).replace(".bam", "") for p in paths]
    out = []; counter = {}
    for b in base:
        if base.count(b) > 1:
            counter[b] = counter.get(b, 0) + 1
            out.append(b + "." + str(counter[b]))
        else:
            out.append(b)
    return out


def port_outputs(vcf_gz, sams, sample, mapq="255", paired_end="1", **kw):
    import gzip
    vt, st, batches, col = load_inputs(vcf_gz, sams, sample, mapq, paired_end)
    P = port.Params(bam_names=bam_display_names(sams), **kw)
    res = port.run(vt, batches, P)
    with gzip.open(vcf_gz, "rt") as f:
        vcf_text, _, _ = port.write_vcf_text(res, vt, f.readlines(), col, P)
    return dict(allelic_counts=res.allelic_counts, allele_config=res.allele_config, haplotypes=res.haplotypes,
                haplotypic_counts=res.haplotypic_counts, variant_connections=res.variant_connections,
                vcf=vcf_text), res


def reference_outputs(vcf_gz, sams, sample, out_prefix, mapq="255", paired_end="1", extra_args=(), hashseed=0):
    r = rr.run_reference(vcf_gz, sams, out_prefix, sample, mapq=mapq, paired_end=paired_end,
                         extra_args=extra_args, hashseed=hashseed)
    if r["returncode"] != 0:
        raise RuntimeError("reference failed:\n" + r["log"][-3000:])
    out = {}
    for suf in rr.OUTPUT_SUFFIXES:
        key = suf.replace(".txt", "").replace(".gz", "")
        out[key] = rr.read_text(r[suf])
    out["log"] = r["log"]
    return out


def make_case(tmp, seed, n_variants, n_pairs, n_bams=1, contigs=None, **read_kw):
    import torch  # noqa: F401
    from phaser_b200 import synth
    contigs = contigs or [("21", 300000), ("22", 200000)]
    g = synth.make_genome(seed, n_variants, contigs=contigs, n_genes=max(2, n_variants // 8))
    vcf = synth.write_vcf(g, os.path.join(tmp, "s%d.vcf.gz" % seed))
    sams = []
    for b in range(n_bams):
        rec = synth.make_reads(g, seed * 100 + b, n_pairs, dup_frac=0.05, **read_kw)
        sams.append(synth.write_sam(rec, g, os.path.join(tmp, "s%d_b%d.bam" % (seed, b)), bam_name="b%d" % b))
    return vcf, sams


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs="+", default=[1])
    ap.add_argument("--variants", type=int, default=300)
    ap.add_argument("--pairs", type=int, default=3000)
    ap.add_argument("--bams", type=int, default=1)
    ap.add_argument("--switch", type=float, default=0.002)
    ap.add_argument("--max_block_size", type=int, default=15)
    ap.add_argument("--keep", default="")
    a = ap.parse_args()
    ok = True
    for seed in a.seeds:
        tmp = a.keep or tempfile.mkdtemp(prefix="pin_")
        os.makedirs(tmp, exist_ok=True)
        vcf, sams = make_case(tmp, seed, a.variants, a.pairs, a.bams, switch_per_base=a.switch)
        t0 = time.time()
        extra = ["--max_block_size", str(a.max_block_size)]
        ref = reference_outputs(vcf, sams, "S1", os.path.join(tmp, "ref%d" % seed), extra_args=extra)
        t1 = time.time()
        got, res = port_outputs(vcf, sams, "S1", max_block_size=a.max_block_size)
        t2 = time.time()
        bad = compare.diff_outputs(ref, got)
        nblk = len(res.final_blocks)
        print("seed %d: reference %.1fs port %.1fs tuples %d edges %d dropped %d blocks %d %s -> %s" % (
            seed, t1 - t0, t2 - t1, res.total_tuples, len(res.edges), res.dropped, nblk, dict(port.STATS),
            "PARITY" if not bad else "MISMATCH"))
        for b in bad:
            print("   " + b)
            ok = False
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
