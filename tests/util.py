"""Shared helpers of the test-suite: build seeded cases, run the product path, run the oracle."""
import ctypes
import gzip
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from phaser_b200 import engine as eng            # noqa: E402
from phaser_b200 import pipeline, writer, vcfio, samio, synth   # noqa: E402
from oracle import port, compare                 # noqa: E402

HOSTSIM_SO = os.path.join(ROOT, "tests", "hostsim", "_phz_hostsim.so")


def build_hostsim():
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    deps = [src] + [os.path.join(ROOT, "phaser_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "phaser_b200", "csrc"))]
    if os.path.exists(HOSTSIM_SO) and all(os.path.getmtime(HOSTSIM_SO) >= os.path.getmtime(d) for d in deps):
        return HOSTSIM_SO
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-shared", "-fPIC", "-o", HOSTSIM_SO, src, "-lz", "-lpthread"])
    return HOSTSIM_SO


def hostsim_engine():
    lib = eng._declare(ctypes.CDLL(build_hostsim()))
    return eng.Engine(device="cpu", lib=lib)


def gpu_engine():
    return eng.Engine(device="cuda:0")


def make_case(tmp, seed, n_variants=300, n_pairs=3000, n_bams=1, contigs=None, dup_frac=0.05, **read_kw):
    contigs = contigs or [("21", 300000), ("22", 200000)]
    g = synth.make_genome(seed, n_variants, contigs=contigs, n_genes=max(2, n_variants // 8))
    vcf = synth.write_vcf(g, os.path.join(str(tmp), "s%d.vcf.gz" % seed))
    sams = []
    for b in range(n_bams):
        rec = synth.make_reads(g, seed * 100 + b, n_pairs, dup_frac=dup_frac, **read_kw)
        sams.append(synth.write_sam(rec, g, os.path.join(str(tmp), "s%d_b%d.bam" % (seed, b)), bam_name="b%d" % b))
    return vcf, sams


def bam_display_names(paths):
    base = [os.path.basename(p).replace(".bam", "") for p in paths]
    out = []; counter = {}
    for b in base:
        if base.count(b) > 1:
            counter[b] = counter.get(b, 0) + 1
            out.append(b + "." + str(counter[b]))
        else:
            out.append(b)
    return out


def load_inputs(vcf_gz, sams, sample="S1", mapq="255", paired_end="1", remove_dups=1, pass_only=1, id_separator="_",
                gw_phase_method=0, blacklist="", haplo_count_blacklist="", include_indels=0, chr=""):
    col = vcfio.sample_column_map(vcf_gz)[sample]
    vt, st = vcfio.parse_vcf(vcf_gz, col, pass_only=pass_only, id_separator=id_separator, gw_phase_method=gw_phase_method,
                             blacklist=blacklist, haplo_count_blacklist=haplo_count_blacklist, include_indels=include_indels,
                             chrom_of_interest=chr)
    fd = samio.FragmentDictionary()
    mq = [int(x) for x in str(mapq).split(",")]; pe = [int(x) for x in str(paired_end).split(",")]
    if len(mq) == 1:
        mq = mq * len(sams)
    if len(pe) == 1:
        pe = pe * len(sams)
    batches = [samio.parse_sam(s, vt.contigs, fd, bool(remove_dups), bool(pe[i]), mq[i]) for i, s in enumerate(sams)]
    return vt, st, batches, col, fd


def product_outputs(engine, vcf_gz, sams, sample="S1", mapq="255", paired_end="1", max_block_size=15,
                    as_q_cutoff=0.05, cc_threshold=0.01, exclude=(), isize=(0.0,), baseq=10, unphased_vars=1,
                    gw_phase_vcf=0, gw_phase_method=0, gw_phase_vcf_min_confidence=0.90, unique_ids=0, pass_only=1,
                    remove_dups=1, id_separator="_", blacklist="", haplo_count_blacklist="", include_indels=0,
                    output_read_ids=0, output_network="", chr=""):
    vt, st, batches, col, fd = load_inputs(vcf_gz, sams, sample, mapq, paired_end, remove_dups, pass_only, id_separator,
                                           gw_phase_method, blacklist, haplo_count_blacklist, include_indels, chr)
    P = pipeline.PhaseParams(baseq=baseq, isize=list(isize), as_q_cutoff=as_q_cutoff, cc_threshold=cc_threshold,
                             max_block_size=max_block_size, haplo_count_bam_exclude=list(exclude),
                             want_read_ids=(output_read_ids == 1), want_kept_tuples=(output_network != ""))
    dev = [engine.upload_reads(b) for b in batches]
    res = pipeline.run_path(engine, vt, dev, P, n_fragments=len(fd.names))
    o = writer.Outputs(res, vt, bam_display_names(sams), P, unphased_vars=unphased_vars, gw_phase_method=gw_phase_method,
                       unique_ids=unique_ids, read_names=fd.names if output_read_ids == 1 else None,
                       output_network=output_network)
    ac = o.allelic_counts(); vc = o.variant_connections()
    hp, hc, cfg = o.block_tables()
    with gzip.open(vcf_gz, "rt") as f:
        vcf_text, _, _ = o.vcf_text(f.readlines(), col, id_separator=id_separator, gw_phase_vcf=gw_phase_vcf,
                                    min_conf=gw_phase_vcf_min_confidence, chrom_of_interest=chr)
    out = dict(allelic_counts=ac, allele_config=cfg, haplotypes=hp, haplotypic_counts=hc, variant_connections=vc, vcf=vcf_text)
    if o.network is not None:
        out["network_links"], out["network_nodes"] = o.network
    return out, res, (vt, batches)


def oracle_outputs(vcf_gz, sams, sample="S1", mapq="255", paired_end="1", pass_only=1, remove_dups=1, exclude=None,
                   blacklist="", haplo_count_blacklist="", include_indels=0, **kw):
    chrom = kw.pop("chr", "")
    vt, st, batches, col, fd = load_inputs(vcf_gz, sams, sample, mapq, paired_end, remove_dups, pass_only,
                                           kw.get("id_separator", "_"), kw.get("gw_phase_method", 0), blacklist,
                                           haplo_count_blacklist, include_indels, chrom)
    if exclude is not None:
        kw["haplo_count_bam_exclude"] = list(exclude)
    if kw.get("output_read_ids", 0) == 1:
        kw["read_names"] = fd.names
    P = port.Params(bam_names=bam_display_names(sams), **kw)
    res = port.run(vt, batches, P)
    with gzip.open(vcf_gz, "rt") as f:
        vcf_text, _, _ = port.write_vcf_text(res, vt, f.readlines(), col, P, chrom_of_interest=chrom)
    out = dict(allelic_counts=res.allelic_counts, allele_config=res.allele_config, haplotypes=res.haplotypes,
               haplotypic_counts=res.haplotypic_counts, variant_connections=res.variant_connections, vcf=vcf_text)
    if hasattr(res, "network_links"):
        out["network_links"] = res.network_links; out["network_nodes"] = res.network_nodes
    return out, res


def compare_tuples(engine, vt, batch, baseq=10, isize=0.0):
    """K1 parity: device candidate tuples (minus the 'nothing printed' class) == oracle mapper tuples."""
    engine.set_variants(vt)
    d = engine.upload_reads(batch)
    engine.map_reads(d, baseq, isize)
    rec = engine.download("t_rec"); var = engine.download("t_var"); misc = engine.download("t_misc")
    cls = misc & 3
    keep = cls != 3
    got = list(zip(rec[keep].tolist(), ((misc[keep] >> 8) & 0xFF).tolist(), var[keep].tolist(), cls[keep].tolist(),
                   ((misc[keep] >> 2) & 1).tolist(), ((misc[keep] >> 4) & 15).tolist(),
                   (misc[keep] >> 16).astype(np.uint16).view(np.int16).tolist()))
    exp = []
    for (r, si, v, s, a) in port.map_reads(batch, vt, baseq, isize):
        info = port.variant_info(vt, v)
        c = info["alleles"].index(s) if s in info["alleles"] else 2
        exp.append((r, min(si, 255), v, c, 1 if len(s) > 1 else 0, port.BASES.index(s[0]), a))
    return got, exp


def packed_vs_plain(engine, vt, batch, n_contigs):
    """K1 through the packed transport form must emit exactly the tuples of the plain arrays."""
    from phaser_b200 import engine as eng
    engine.set_variants(vt)
    engine.map_reads(engine.upload_reads(batch), 10, 0.0)
    plain = [engine.download(k).copy() for k in ("t_rec", "t_var", "t_misc")]
    packed = eng.pack_reads(batch, n_contigs, threads=3, lib=engine.lib)
    n = engine.map_reads_packed(packed, 10, 0.0)
    # what the device expanded must be the original arrays, bit for bit
    nb = int(batch.qual.shape[0])
    for name, orig in (("st_pos", batch.pos), ("st_tlen", batch.tlen), ("st_as", batch.aln_score), ("st_cig", batch.cigar),
                       ("st_coff", batch.cigar_off), ("st_soff", batch.seq_off), ("st_qual", batch.qual), ("st_frag", batch.frag)):
        back = engine.download(name)
        assert np.array_equal(back.view(np.asarray(orig).dtype), np.asarray(orig)), name
    seq = engine.download("st_seq"); o = np.asarray(batch.seq)
    assert np.array_equal(seq[:nb // 2], o[:nb // 2]) and (nb % 2 == 0 or (seq[nb // 2] >> 4) == (o[nb // 2] >> 4)), "st_seq"
    got = [engine.download(k).copy() for k in ("t_rec", "t_var", "t_misc")]
    assert n == plain[0].shape[0]
    for a, b in zip(plain, got):
        assert np.array_equal(a, b)
    return packed
