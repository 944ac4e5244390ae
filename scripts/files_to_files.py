#!/usr/bin/env python
"""Files to files at scale: the drop-in command line on a BGZF BAM + VCF against the unmodified
reference (compiled as its README prescribes, --threads = host cores) on the SAM-text twin of the SAME records, same box,
same run; the six output files are diffed (oracle.compare).  One JSON line on stdout.

    python scripts/files_to_files.py --pairs 10000000 --variants 400000 [--ref_timeout 1800] [--no_reference]

The sample is generated on the GPU (same generator and shape as bench.py's configs[1] workload), written once as a BAM
(product input) and as per-contig SAM text restricted to records that overlap a het site -- what the reference's two
`samtools view` stages (contig split, -L BED) would hand its mapper, done beforehand and charged to nobody."""
import argparse
import contextlib
import io
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np      # noqa: E402
import torch            # noqa: E402


def write_vcf_fast(g, path):
    """synth.write_vcf for millions of sites (array -> text without per-variant objects), BGZF-compressed"""
    from phaser_b200 import bgzf
    names = [c[0] for c in g.contigs]
    vc = g.v_contig.cpu().numpy(); pos = g.v_pos.cpu().numpy()
    lut = {1: "A", 2: "C", 4: "G", 8: "T"}
    ref = [lut[int(x)] for x in g.v_ref.cpu().numpy().tolist()]; alt = [lut[int(x)] for x in g.v_alt.cpu().numpy().tolist()]
    ph = g.v_phased.cpu().numpy().tolist(); fa = g.v_gt_first_alt.cpu().numpy().tolist(); nm = g.v_named.cpu().numpy().tolist()
    af = g.v_af.cpu().numpy().tolist()
    head = ["##fileformat=VCFv4.2\n"] + ["##contig=<ID=%s,length=%d>\n" % c for c in g.contigs] + [
        '##INFO=<ID=AF,Number=A,Type=Float,Description="Allele frequency">\n',
        '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n', "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\n"]
    rows = ["%s\t%d\t%s\t%s\t%s\t100\tPASS\tAF=%.4f\tGT\t%s\n" % (
        names[c], p, ("rs%d" % (i + 1)) if n else ".", r, a, f, ("1|0" if x else "0|1") if h else "0/1")
        for i, (c, p, n, r, a, f, h, x) in enumerate(zip(vc.tolist(), pos.tolist(), nm, ref, alt, af, ph, fa))]
    data = ("".join(head) + "".join(rows)).encode()
    with open(path, "wb") as f:
        for b in bgzf.compress_all(data):
            f.write(b)
        f.write(bgzf.EOF_BLOCK)
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=10_000_000)
    ap.add_argument("--variants", type=int, default=400_000)
    ap.add_argument("--seed", type=int, default=2000)
    ap.add_argument("--ref_timeout", type=int, default=1800)
    ap.add_argument("--no_reference", action="store_true")
    ap.add_argument("--keep", default="")
    a = ap.parse_args()
    from phaser_b200 import synth, engine as eng, phaser as cli
    from oracle import compare
    from oracle.harness import run_reference as rr
    dev = torch.device("cuda", 0)
    tmp = a.keep or tempfile.mkdtemp(prefix="phz_f2f_")
    os.makedirs(tmp, exist_ok=True)
    out = {"pairs": a.pairs, "cores": os.cpu_count()}
    t0 = time.time()
    g = synth.make_genome(a.seed, a.variants, exonic_frac=0.10, device=dev, n_genes=max(2, int(a.variants * 0.10) // 8))
    parts = []; done = 0; chunk = 2_000_000
    while done < a.pairs:
        n = min(chunk, a.pairs - done)
        rec = synth.make_reads(g, a.seed * 1000 + done // chunk, n, chunk_pairs=chunk)
        rec["frag"] = rec["frag"] + done
        parts.append(synth.compact_raw(rec)); done += n
    rec = synth.concat_sorted(parts); del parts
    # records that overlap a het site (what `samtools view -L sites.bed` keeps): reference span from the CIGAR table
    span = (rec["opl"].to(torch.int64) * ((rec["ops"] == 0) | (rec["ops"] == 2) | (rec["ops"] == 3))).sum(1)
    key = rec["contig"].to(torch.int64) * (1 << 32) + rec["pos"].to(torch.int64)
    vkey = g.v_contig * (1 << 32) + g.v_pos
    lo = torch.searchsorted(vkey, key); hi = torch.searchsorted(vkey, key + span)
    overlaps = (hi > lo).cpu().numpy()
    host = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in rec.items()}
    del rec, span, key, lo, hi
    torch.cuda.empty_cache()
    V = int(g.v_pos.shape[0]); R = int(host["pos"].shape[0])
    out.update(het_snvs=V, records=R, records_overlapping_a_site=int(overlaps.sum()), generate_s=round(time.time() - t0, 1))
    t0 = time.time()
    vcf = write_vcf_fast(g, os.path.join(tmp, "sample.vcf.gz"))
    bam = eng.write_sam_native(host, g.contigs, os.path.join(tmp, "sample.bam"), bam_name="bam0", bam=True)
    out["write_inputs_s"] = round(time.time() - t0, 1); out["bam_bytes"] = os.path.getsize(bam); out["vcf_bytes"] = os.path.getsize(vcf)
    # ---- product: the drop-in command line, BAM + VCF in, six files (+ index) out
    E = eng.Engine(device=dev)
    argv = ["--vcf", vcf, "--bam", bam, "--sample", "S1", "--mapq", "255", "--baseq", "10", "--paired_end", "1",
            "--o", os.path.join(tmp, "ours"), "--threads", str(os.cpu_count() or 1)]
    runs = []
    for _ in range(2):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            cli.run(cli.build_parser().parse_args(argv), engine=E)
        runs.append((time.perf_counter() - t0, {k: round(v, 3) for k, v in cli.LAST_STAGE_SECONDS.items()}))
    best = min(runs, key=lambda r: r[0])
    out["product"] = {"seconds": best[0], "het_snvs_per_sec": V / best[0], "records_per_sec": R / best[0], "stage_seconds": best[1],
                      "all_runs_s": [round(r[0], 3) for r in runs], "input": "BGZF BAM + BGZF VCF", "backend": E.backend}
    # ---- reference: unmodified, compiled, --threads = cores, SAM text twin (pre-split per contig, -L filter applied)
    if not a.no_reference and rr.compiled_available():
        t0 = time.time()
        split = os.path.join(tmp, "per_contig"); os.makedirs(split, exist_ok=True)
        cont = host["contig"].numpy()
        sel_all = np.flatnonzero(overlaps)
        for ci, (name, _len) in enumerate(g.contigs):
            sel = sel_all[cont[sel_all] == ci]
            if sel.shape[0] == 0:
                continue
            sub = {k: (v[torch.from_numpy(sel)] if torch.is_tensor(v) else v) for k, v in host.items()}
            eng.write_sam_native(sub, g.contigs, os.path.join(split, name + ".sam"), bam_name="bam0")
        open(os.path.join(tmp, "twin.bam"), "w").write("@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % c for c in g.contigs))
        out["write_reference_twin_s"] = round(time.time() - t0, 1)
        gz = os.path.join(tmp, "ref_in.vcf.gz")
        shutil.copy(vcf, gz)
        t0 = time.perf_counter()
        try:
            r = rr.run_reference(gz, [os.path.join(tmp, "twin.bam")], os.path.join(tmp, "ref"), "S1", threads=os.cpu_count() or 1,
                                 compiled=True, fast_shim_dir=split, timeout=a.ref_timeout)
            dt = time.perf_counter() - t0
            out["reference"] = {"seconds": dt, "het_snvs_per_sec": V / dt, "returncode": r["returncode"], "threads": os.cpu_count(),
                                "kind": "unmodified reference, Cython-compiled as its README prescribes (oracle/_ref)",
                                "log_tail": r["log"][-600:]}
            if r["returncode"] == 0:
                names = {"allelic_counts": "allelic_counts.txt", "allele_config": "allele_config.txt", "haplotypes": "haplotypes.txt",
                         "haplotypic_counts": "haplotypic_counts.txt", "variant_connections": "variant_connections.txt", "vcf": "vcf.gz"}
                ref = {k: rr.read_text(r[s]) for k, s in names.items() if s in r}
                got = {k: rr.read_text(os.path.join(tmp, "ours." + s)) for k, s in names.items()}
                # the BAM display name is the only intended difference: ours reads sample.bam, the reference twin.bam
                got["haplotypic_counts"] = got["haplotypic_counts"].replace("\tsample\t", "\ttwin\t")
                bad = compare.diff_outputs(ref, got)
                out["parity"] = {"identical": not bad, "differences": [b[:300] for b in bad[:4]],
                                 "rows": {k: ref[k].count("\n") for k in ref}}
                out["speedup_files_to_files"] = dt / best[0]
        except Exception as e:      # noqa: BLE001 -- a timeout is a result, not a failure of this script
            out["reference"] = {"error": repr(e)[:300], "seconds_at_least": time.perf_counter() - t0}
    print(json.dumps(out))
    if not a.keep:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
