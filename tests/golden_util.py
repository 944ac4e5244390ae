"""Load the committed golden cases (tests/golden/cases, produced by tests/golden/make_golden.py)."""
import json
import os

CASES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cases")
FILES = {"allelic_counts": "ref.allelic_counts.txt", "allele_config": "ref.allele_config.txt",
         "haplotypes": "ref.haplotypes.txt", "haplotypic_counts": "ref.haplotypic_counts.txt",
         "variant_connections": "ref.variant_connections.txt", "vcf": "ref.vcf"}


def case_names():
    return sorted(d for d in os.listdir(CASES) if os.path.isdir(os.path.join(CASES, d)))


def load_case(name):
    d = os.path.join(CASES, name)
    with open(os.path.join(d, "case.json")) as f:
        meta = json.load(f)
    ref = {k: open(os.path.join(d, fn)).read() for k, fn in FILES.items()}
    for k in ("network_links", "network_nodes"):        # --output_network cases only
        fn = os.path.join(d, "ref." + k.replace("_", ".") + ".txt")
        if os.path.exists(fn):
            ref[k] = open(fn).read()
    src = os.path.join(CASES, meta["inputs"]) if "inputs" in meta else d
    sams = [os.path.join(src, b) for b in meta["bams"]]
    mapper = {b: open(os.path.join(d, "ref.mapper.%s.tsv" % b.replace("/", "_"))).read() for b in meta["bams"]}
    return dict(dir=d, vcf=os.path.join(src, "in.vcf.gz"), sams=sams, meta=meta, ref=ref, mapper=mapper)


def args_to_kw(args):
    """reference CLI flags of a case -> keyword arguments understood by tests.util helpers"""
    kw = {}
    it = iter(args)
    for a in it:
        v = next(it)
        if a == "--as_q_cutoff":
            kw["as_q_cutoff"] = float(v)
        elif a == "--isize":
            kw["isize"] = [float(x) for x in v.split(",")]
        elif a == "--max_block_size":
            kw["max_block_size"] = int(v)
        elif a == "--haplo_count_bam_exclude":
            kw["exclude"] = [int(x) - 1 for x in v.split(",")]
        elif a in ("--gw_phase_method", "--gw_phase_vcf", "--unphased_vars", "--unique_ids", "--pass_only", "--remove_dups",
                   "--include_indels", "--output_read_ids"):
            kw[a[2:]] = int(v)
        elif a in ("--gw_phase_vcf_min_confidence", "--cc_threshold"):
            kw[a[2:]] = float(v)
        elif a in ("--id_separator", "--output_network", "--chr"):
            kw[a[2:]] = v
        elif a in ("--blacklist", "--haplo_count_blacklist"):
            kw[a[2:]] = os.path.join(os.path.dirname(CASES), v)
        else:
            raise KeyError(a)
    return kw
