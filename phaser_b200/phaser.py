#!/usr/bin/env python
"""phaser.py -- drop-in command line for the read -> variant -> haplotype path of phASER on B200.

Same flags, file formats and fatal-error behaviour as the reference CLI (phaser/phaser.py:26-178);
the work between "het sites loaded" and "files written" runs on the GPU through the C ABI
(include/phz.h).  Differences a user can see:
  * no samtools / bgzip / tabix / bedtools / bcftools are needed (BAM or SAM text is read directly);
  * not yet supported, rejected with a FATAL ERROR instead of being silently ignored:
    --process_slow 1 (per-contig noise changes the results in the reference and is broken on python 3);
  * fields the reference prints in CPython-set order come out in a canonical order
    (SURVEY.md section 8c).
"""
import argparse
import datetime
import gzip
import os
import sys
import time

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from phaser_b200 import vcfio, samio, pipeline, writer, bgzf, tabix    # noqa: E402
from phaser_b200.vcfio import PhaserFatal                        # noqa: E402

VERSION = "1.2.0"


def build_parser():
    p = argparse.ArgumentParser()
    # required (phaser.py:30-36)
    p.add_argument("--bam", required=False, default='', help="Indexed BAMs (comma separated) containing aligned reads")
    p.add_argument("--vcf", required=True, default='', help="VCF for the sample, must be gzipped and tabix indexed.")
    p.add_argument("--sample", required=False, default='', help="Sample name in VCF")
    p.add_argument("--mapq", required=True, help="Minimum MAPQ for reads to be used for phasing (comma separated list allowed).")
    p.add_argument("--baseq", type=int, required=True, help="Minimum baseq for bases to be used for phasing")
    p.add_argument("--paired_end", required=True, help="Sequencing data comes from a paired end assay (0,1; comma separated list allowed).")
    p.add_argument("--o", required=True, help="Out prefix")
    # optional (phaser.py:40-53)
    p.add_argument("--python_string", default="python3")
    p.add_argument("--haplo_count_bam_exclude", default="")
    p.add_argument("--haplo_count_blacklist", default="")
    p.add_argument("--cc_threshold", type=float, default=0.01)
    p.add_argument("--isize", default="0")
    p.add_argument("--as_q_cutoff", type=float, default=0.05)
    p.add_argument("--blacklist", default="")
    p.add_argument("--write_vcf", type=int, default=1)
    p.add_argument("--include_indels", type=int, default=0)
    p.add_argument("--output_read_ids", type=int, default=0)
    p.add_argument("--remove_dups", type=int, default=1)
    p.add_argument("--pass_only", type=int, default=1)
    p.add_argument("--unphased_vars", type=int, default=1)
    p.add_argument("--chr_prefix", type=str, default="")
    # genome wide phasing (phaser.py:56-59)
    p.add_argument("--gw_phase_method", type=int, default=0)
    p.add_argument("--gw_af_field", default="AF")
    p.add_argument("--gw_phase_vcf", type=int, default=0)
    p.add_argument("--gw_phase_vcf_min_confidence", type=float, default=0.90)
    # performance (phaser.py:62-65)
    p.add_argument("--threads", type=int, default=1, help="host threads for BAM decompression")
    p.add_argument("--max_block_size", type=int, default=15)
    p.add_argument("--temp_dir", default="")
    p.add_argument("--max_items_per_thread", type=int, default=100000)
    # debug / reporting (phaser.py:68-78)
    p.add_argument("--show_warning", type=int, default=0)
    p.add_argument("--debug", type=int, default=0)
    p.add_argument("--chr", default="")
    p.add_argument("--unique_ids", type=int, default=0)
    p.add_argument("--id_separator", default="_")
    p.add_argument("--output_network", default="")
    p.add_argument("--process_slow", type=int, default=0, required=False)
    # this implementation only
    p.add_argument("--device", default="cuda:0", help="CUDA device to run on")
    p.add_argument("--soa_cache", default=os.environ.get("PHZ_SOA_CACHE", ""),
                   help="directory of the SoA cache (parse once): the first BAM's packed arrays are kept there and mapped back "
                        "in by later runs on the same BAM, contigs and read filters")
    return p


def say(text=""):
    print(text)
    sys.stdout.flush()


def fatal_error(text):
    """phaser.py:2032-2034"""
    say("     FATAL ERROR: " + text)
    sys.exit(1)


def bam_display_names(bam_list):
    """phaser.py:469-480"""
    file_names = [os.path.basename(x).replace(".bam", "") for x in bam_list]
    out, counter = [], {}
    for x in file_names:
        if file_names.count(x) > 1:
            counter[x] = counter.get(x, 0) + 1
            out.append(x + "." + str(counter[x]))
        else:
            out.append(x)
    return out


def per_bam_lists(args, bam_list):
    """phaser.py:482-513"""
    mapq = args.mapq.split(",")
    if len(mapq) == 1 and len(bam_list) > 1:
        mapq = mapq * len(bam_list)
    elif len(mapq) != len(bam_list):
        fatal_error("Number of mapq values and input BAMs does not match. Supply either one mapq to be used for all BAMs or one mapq per input BAM.")
    isize = args.isize.split(",")
    if len(isize) == 1 and len(bam_list) > 1:
        isize = isize * len(bam_list)
    elif len(mapq) != len(isize):
        fatal_error("Number of isize values and input BAMs does not match. Supply either one isize to be used for all BAMs or one isize per input BAM.")
    isize = list(map(float, isize))
    paired = args.paired_end.split(",")
    if len(paired) == 1 and len(bam_list) > 1:
        paired = paired * len(bam_list)
    elif len(paired) != len(bam_list):
        fatal_error("Number of paired_end values and input BAMs does not match. Supply either one paired_end to be used for all BAMs or one paired_end per input BAM.")
    return [int(x) for x in mapq], isize, [int(x) for x in paired]


_TRACE = bool(os.environ.get("PHZ_TRACE"))
LAST_STAGE_SECONDS = {}          # stage -> seconds of the last run() in this process (bench.py reports them)


def _trace(t0, what):
    dt = time.perf_counter() - t0
    LAST_STAGE_SECONDS[what] = LAST_STAGE_SECONDS.get(what, 0.0) + dt
    if _TRACE:
        print("[phaser.py] %-28s %8.1f ms" % (what, dt * 1e3), file=sys.stderr)
    return time.perf_counter()


def _warm_imports():
    """scipy.stats takes 1-2 s to import and is first needed after the graph stage (critical values): load it on a
    thread while the VCF and the BAMs are being read."""
    import threading

    def work():
        try:
            import scipy.stats  # noqa: F401
        except Exception:
            pass
    threading.Thread(target=work, daemon=True).start()


def run(args, engine=None):
    _warm_imports()
    say("")
    say("##################################################")
    say("              Welcome to phASER v%s" % VERSION)
    say("  Author: Stephane Castel (stephanecastel@gmail.com)")
    say("  Updated by: Bishwa K. Giri (bkgiri@uncg.edu)")
    say("  B200-native read->variant->haplotype path (phaser_b200)")
    say("##################################################")
    say("")
    for flag, bad, why in (("--process_slow 1", args.process_slow == 1, "per-contig mode changes results in the reference"),):
        if bad:
            fatal_error("%s is not supported by the B200 path yet (%s)." % (flag, why))
    if args.id_separator == ":" or args.id_separator == "":
        fatal_error("ID separator must not be ':' or blank. Please choose another separator that is not found in the contig names.")
    if os.path.isfile(args.vcf) is False:
        fatal_error("VCF file does not exist.")
    if args.vcf.endswith(".gz") is False and args.vcf.endswith(".bgz") is False:
        fatal_error("VCF must be gzipped.")
    for xfile in (args.blacklist, args.haplo_count_blacklist):
        if xfile != "" and os.path.isfile(xfile) is False:
            fatal_error("File: %s not found." % xfile)
    bam_list = [b for b in args.bam.split(",")]
    for b in bam_list:
        if b != "" and os.path.isfile(b) is False:
            fatal_error("File: %s not found." % b)
    # the VCF is inflated and split into lines ONCE by the native library; the site table is parsed from that text on
    # --threads host threads and the output VCF is written from it (blacklists still take the general Python parser)
    from phaser_b200.engine import load_library
    lib = engine.lib if engine is not None else load_library()
    nv = None
    try:
        nv = vcfio.NativeVcf(args.vcf, lib, threads=max(1, args.threads))
        cols = nv.sample_column_map()
    except PhaserFatal:
        nv = None
        cols = vcfio.sample_column_map(args.vcf)
    if args.sample not in cols:
        fatal_error("Sample '%s' not found in the input VCF file." % args.sample)
    sample_column = cols[args.sample]
    if args.haplo_count_bam_exclude != "":
        exclude = [x - 1 for x in map(int, args.haplo_count_bam_exclude.split(","))]
    else:
        exclude = []
    start_time = time.time()
    say('STARTED "Read backed phasing and ASE/haplotype analyses" ... ')
    say("    DATE, TIME : %s" % datetime.datetime.now().strftime('%Y-%m-%d, %H:%M:%S'))
    say("#1. Loading heterozygous variants into intervals...")
    LAST_STAGE_SECONDS.clear()
    _t = time.perf_counter()
    vt = None
    native_table = False
    if nv is not None and args.blacklist == "" and args.haplo_count_blacklist == "":
        try:
            vt, st = vcfio.parse_vcf_native(nv, sample_column, pass_only=args.pass_only, chrom_of_interest=args.chr,
                                            chr_prefix=args.chr_prefix, id_separator=args.id_separator,
                                            include_indels=args.include_indels, gw_phase_method=args.gw_phase_method,
                                            gw_af_field=args.gw_af_field)
            native_table = True
        except PhaserFatal as e:
            if not ("malformed" in str(e) or "carriage" in str(e) or "POS is not" in str(e)):
                fatal_error(str(e))
            vt = None          # odd input: the general parser decides (and fails the way the reference does)
    if vt is None:
        try:
            vt, st = vcfio.parse_vcf(args.vcf, sample_column, pass_only=args.pass_only, chrom_of_interest=args.chr,
                                     chr_prefix=args.chr_prefix, id_separator=args.id_separator,
                                     include_indels=args.include_indels, gw_phase_method=args.gw_phase_method,
                                     gw_af_field=args.gw_af_field, blacklist=args.blacklist,
                                     haplo_count_blacklist=args.haplo_count_blacklist)
        except PhaserFatal as e:
            fatal_error(str(e))
    say("          %d heterozygous sites being used for phasing (%d filtered, %d indels excluded, %d unphased)" % (
        st.het_count, st.filter_count, st.indels_excluded, st.unphased_count))
    say()
    if st.het_count == 0:
        fatal_error("No heterozygous sites that passed all filters were included in the analysis, phASER cannot continue. Check blacklist and pass_only arguments.")
    say("#2. Retrieving reads that overlap heterozygous sites...")
    _t = _trace(_t, "parse_vcf")
    mapq, isize, paired = per_bam_lists(args, bam_list)
    bam_names = bam_display_names(bam_list)
    if engine is None:
        from phaser_b200.engine import Engine
        engine = Engine(device=args.device)
    from phaser_b200.engine import NativeFragmentDictionary, read_alignments_native, PhzError, pack_reads
    fd = NativeFragmentDictionary(engine.lib)
    batches = []
    n_frag_cached = 0
    for i, bam in enumerate(bam_list):
        say("     file: %s" % bam)
        say("          minimum mapq: %s" % mapq[i])
        rb = None; ckey = None
        if args.soa_cache and i == 0:          # parse once: the first BAM's arrays may already sit in the cache
            from phaser_b200 import soacache
            ckey = soacache.cache_key(bam, vt.contigs, args.remove_dups == 1, paired[i] == 1, mapq[i])
            rb = soacache.load(args.soa_cache, ckey, fd, need_names=(len(bam_list) > 1 or args.output_read_ids == 1),
                               threads=max(1, args.threads))
            if rb is not None:
                n_frag_cached = rb.n_fragments
                say("          (packed arrays from the SoA cache)")
        if rb is None:
            try:        # native reader: BGZF blocks inflated on --threads host threads, records decoded straight into SoA
                rb = read_alignments_native(bam, vt.contigs, fd, remove_dups=(args.remove_dups == 1), proper_pair=(paired[i] == 1),
                                            min_mapq=mapq[i], threads=max(1, args.threads), lib=engine.lib)
            except PhzError as e:
                fatal_error(str(e))
            if ckey is not None:
                soacache.save(args.soa_cache, ckey, rb, fd)
        _t = _trace(_t, "read_alignments_native")
        # One-shot run: the plain arrays go up as they are.  The packed transport form (engine.pack_reads) pays off
        # when host buffers are copied more than once or prefetched (sample loops, bench.py); packing once for a
        # single copy costs more host time than it saves on the bus.  PHZ_PACK=1 forces it.
        if os.environ.get("PHZ_PACK"):
            try:
                batches.append(pack_reads(rb, len(vt.contigs), threads=max(1, args.threads), lib=engine.lib))
            except PhzError:            # a record beyond 65535 CIGAR ops / bases: plain arrays
                batches.append(engine.upload_reads(rb))
        else:
            batches.append(engine.upload_reads(rb))
        del rb
        _t = _trace(_t, "pack / upload")
    P = pipeline.PhaseParams(baseq=args.baseq, isize=isize, as_q_cutoff=args.as_q_cutoff, cc_threshold=args.cc_threshold,
                             max_block_size=args.max_block_size, haplo_count_bam_exclude=exclude,
                             want_read_ids=(args.output_read_ids == 1), want_kept_tuples=(args.output_network != ""))
    try:
        res = pipeline.run_path(engine, vt, batches, P, n_fragments=max(len(fd), n_frag_cached), reuse_result_buffer=True)
    except PhaserFatal as e:
        fatal_error(str(e))
    _t = _trace(_t, "run_path (device)")
    for i, bam in enumerate(bam_list):
        if res.as_cutoff[i] is not None:
            say("          using alignment score cutoff of %d" % res.as_cutoff[i])
        say("          retrieved %d reads" % res.tuples_per_bam[i])
    say("#3. Identifying connected variants...")
    say("     sequencing noise level estimated at %f" % res.noise_e)
    out = writer.Outputs(res, vt, bam_names, P, unphased_vars=args.unphased_vars, gw_phase_method=args.gw_phase_method,
                         unique_ids=args.unique_ids, read_names=fd.names if args.output_read_ids == 1 else None,
                         output_network=args.output_network, lib=engine.lib, threads=max(1, args.threads))
    out.prefetch_meta()
    with open(args.o + ".variant_connections.txt", "w") as f:
        f.write(out.variant_connections())
    say("     %d variant connections dropped because of conflicting configurations (threshold = %f)" % (
        res.counters["dropped"], args.cc_threshold))
    with open(args.o + ".allelic_counts.txt", "w") as f:
        f.write(out.allelic_counts())
    say("     %d variants covered by at least 1 read" % out.covered_count)
    say("#4. Identifying haplotype blocks...")
    say("#5. Phasing blocks...")
    say("#6. Outputting haplotypes...")
    hp, hc, cfg = out.block_tables()
    with open(args.o + ".haplotypes.txt", "w") as f:
        f.write(hp)
    with open(args.o + ".haplotypic_counts.txt", "w") as f:
        f.write(hc)
    with open(args.o + ".allele_config.txt", "w") as f:
        f.write(cfg)
    if out.network is not None:                      # phaser.py:1128-1157
        with open(args.o + ".network.links.txt", "w") as f:
            f.write(out.network[0])
        with open(args.o + ".network.nodes.txt", "w") as f:
            f.write(out.network[1])
    unphased_phased = phase_corrected = 0
    _t = _trace(_t, "text tables")
    if args.write_vcf == 1:
        say("#7. Outputting phased VCF...")
        text = None
        if native_table:
            try:
                text, unphased_phased, phase_corrected, records = out.vcf_native(
                    nv, gw_phase_vcf=args.gw_phase_vcf, min_conf=args.gw_phase_vcf_min_confidence, chrom_of_interest=args.chr,
                    id_separator=args.id_separator, chr_prefix=args.chr_prefix)
            except writer.PhzUnsupported:
                text = None
        if text is None:
            with gzip.open(args.vcf, "rt") as f:
                text, unphased_phased, phase_corrected = out.vcf_text(
                    f, sample_column, id_separator=args.id_separator, gw_phase_vcf=args.gw_phase_vcf,
                    min_conf=args.gw_phase_vcf_min_confidence, chrom_of_interest=args.chr)
            records = getattr(out, "vcf_records", None)
        # what `bgzip -f` + `tabix -f -p vcf [--csi]` write (phaser.py:1847-1853); --csi iff the input VCF has a .csi (:131)
        if text is not None and native_table and records is not None and len(records) == 4:
            text = None
            out.vcf_save_native(nv, args.o + ".vcf.gz", csi=os.path.isfile(args.vcf + ".csi"))
        else:
            tabix.write_vcf_with_index(args.o + ".vcf.gz", text, csi=os.path.isfile(args.vcf + ".csi"), records=records)
            text = None
    _t = _trace(_t, "vcf out")
    total_time = time.time() - start_time
    say('')
    say("     COMPLETED using %d reads in %d seconds on %s" % (sum(res.tuples_per_bam), total_time, engine.backend))
    n_phased = len(out.all_variants)
    say("     PHASED  %d of %d all variants (= %f) with at least one other variant" % (
        n_phased, st.het_count, float(n_phased) / float(st.het_count)))
    if args.write_vcf == 1:
        if st.unphased_count > 0:
            say("     GENOME WIDE PHASED  %d of %d unphased variants (= %f)" % (
                unphased_phased, st.unphased_count, float(unphased_phased) / float(st.unphased_count)))
        say("     GENOME WIDE PHASE CORRECTED  %d of %d variants (= %f)" % (
            phase_corrected, st.het_count, float(phase_corrected) / float(st.het_count)))
    say('The End.')
    return res


def main(argv=None):
    args = build_parser().parse_args(argv)
    run(args)


if __name__ == "__main__":
    main()
