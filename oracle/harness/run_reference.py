"""Run the UNMODIFIED reference (secastel/phaser, /root/reference/phaser/phaser.py) as an oracle.

TEST INFRASTRUCTURE ONLY.  Nothing under phaser_b200/ may import this module.  It exists to
(1) pin oracle/port.py against the real reference and (2) generate the golden fixtures under
tests/golden/ (see tests/golden/make_golden.py).  It needs /root/reference, which exists only in
the build container -- never on the GPU box.

How: the reference imports `pysam` without using it (phaser/phaser.py:15)
and reaches samtools/bgzip/tabix/bedtools/bcftools only through bash pipelines
(phaser/phaser.py:97-101, 1346, 1851), so a stub module on PYTHONPATH plus the PATH shims in
oracle/harness/bin make it run end to end on SAM *text* passed as --bam.
"""
import gzip
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_DIR = os.environ.get("PHASER_REFERENCE_DIR", "/root/reference/phaser")

OUTPUT_SUFFIXES = [
    "haplotypic_counts.txt", "haplotypes.txt", "allele_config.txt",
    "allelic_counts.txt", "variant_connections.txt", "vcf.gz",
]


COMPILED_DIR = os.path.join(os.path.dirname(HERE), "_ref")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "phaser.py"))


def compiled_available():
    """oracle/_ref: the reference compiled as-is by oracle/build_ref.py (travels to the GPU box)."""
    return all(os.path.isfile(os.path.join(COMPILED_DIR, f)) for f in ("phaser.so", "read_variant_map.so", "run_phaser.py"))


def split_sam_per_contig(sam_path, out_dir, vcf_gz=None):
    """Pre-split a SAM text file per contig for the fast shim (header repeated in every part).  With
    `vcf_gz`, records overlapping no VCF position are dropped here, outside any timing -- what the
    reference's `samtools view -L bed` stage (phaser.py:1346) does in C."""
    import bisect
    os.makedirs(out_dir, exist_ok=True)
    sites = None
    if vcf_gz:
        sites = {}
        with gzip.open(vcf_gz, "rt") as f:
            for line in f:
                if line[0] != "#":
                    c = line.split("\t", 2)
                    sites.setdefault(c[0], []).append(int(c[1]))
        for k in sites:
            sites[k].sort()
    header = []; outs = {}
    with open(sam_path) as f:
        for line in f:
            if line[0] == "@":
                header.append(line); continue
            c = line.split("\t", 6)
            chrom = c[2]
            if sites is not None:
                n = 0; span = 0
                for ch in c[5]:
                    if ch.isdigit():
                        n = n * 10 + ord(ch) - 48
                    else:
                        if ch in "MDN=X":
                            span += n
                        n = 0
                lo = int(c[3]); s = sites.get(chrom)
                if not s:
                    continue
                j = bisect.bisect_left(s, lo)
                if j >= len(s) or s[j] >= lo + max(span, 1):
                    continue
            o = outs.get(chrom)
            if o is None:
                o = outs[chrom] = open(os.path.join(out_dir, chrom + ".sam"), "w")
                o.writelines(header)
            o.write(line)
    for o in outs.values():
        o.close()
    return out_dir


def _env(hashseed):
    env = dict(os.environ)
    env["PATH"] = os.path.join(HERE, "bin") + os.pathsep + env.get("PATH", "")
    env["PYTHONPATH"] = os.path.join(HERE, "stub") + os.pathsep + env.get("PYTHONPATH", "")
    env["PYTHONHASHSEED"] = str(hashseed)
    return env


def run_mapper(sam_path, variant_table, out_path, baseq=10, isize_cutoff=0, hashseed=0):
    """L1 oracle: reference call_read_variant_map.py on SAM text (phaser/call_read_variant_map.py:7-19)."""
    cmd = [sys.executable, os.path.join(REFERENCE_DIR, "call_read_variant_map.py"),
           "--variant_table", variant_table, "--o", out_path, "--baseq", str(baseq),
           "--splice", "1", "--isize_cutoff", str(isize_cutoff)]
    with open(sam_path, "rb") as fin:
        subprocess.run(cmd, stdin=fin, stdout=subprocess.DEVNULL, check=True, env=_env(hashseed))
    return out_path


def run_reference(vcf_gz, bams, out_prefix, sample, mapq="255", baseq=10, paired_end="1",
                  extra_args=(), hashseed=0, threads=1, quiet=True, timeout=None, compiled=False, fast_shim_dir=None):
    """L2/L3 oracle: the whole reference phaser.py (phaser/phaser.py:26-178).

    `bams` are SAM text files (coordinate sorted, with @SQ lines).  Empty `.bai` / `.tbi` files are
    created next to the inputs because the reference only checks that they exist
    (phaser/phaser.py:124, 190).  Returns {suffix: path}.
    """
    if compiled:
        if not compiled_available():
            raise RuntimeError("compiled reference not found in %s (run oracle/build_ref.py)" % COMPILED_DIR)
        script = os.path.join(COMPILED_DIR, "run_phaser.py")
    else:
        if not reference_available():
            raise RuntimeError("reference not found at %s" % REFERENCE_DIR)
        script = os.path.join(REFERENCE_DIR, "phaser.py")
    if isinstance(bams, str):
        bams = [bams]
    for b in bams:
        if not os.path.exists(b + ".bai"):
            open(b + ".bai", "w").close()
    if not os.path.exists(vcf_gz + ".tbi") and not os.path.exists(vcf_gz + ".csi"):
        open(vcf_gz + ".tbi", "w").close()
    cmd = [sys.executable, script,
           "--vcf", vcf_gz, "--bam", ",".join(bams), "--sample", sample,
           "--mapq", str(mapq), "--baseq", str(baseq), "--paired_end", str(paired_end),
           "--o", out_prefix, "--threads", str(threads)] + [str(a) for a in extra_args]
    env = _env(hashseed)
    if fast_shim_dir:
        env["PHZ_FAST_SHIM_DIR"] = fast_shim_dir
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    log = res.stdout.decode("utf-8", "replace")
    if not quiet:
        sys.stdout.write(log)
    out = {"log": log, "returncode": res.returncode}
    for suf in OUTPUT_SUFFIXES:
        p = out_prefix + "." + suf
        if os.path.exists(p):
            out[suf] = p
    return out


def read_text(path):
    if path.endswith(".gz"):
        with gzip.open(path, "rt") as f:
            return f.read()
    with open(path) as f:
        return f.read()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--vcf", required=True)
    ap.add_argument("--bam", required=True)
    ap.add_argument("--sample", required=True)
    ap.add_argument("--o", required=True)
    ap.add_argument("--mapq", default="255")
    ap.add_argument("--paired_end", default="1")
    ap.add_argument("rest", nargs="*")
    a = ap.parse_args()
    r = run_reference(a.vcf, a.bam.split(","), a.o, a.sample, a.mapq, 10, a.paired_end, a.rest, quiet=False)
    sys.exit(r["returncode"])
