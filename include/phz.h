/* phz.h -- C ABI of the B200-native read -> variant -> haplotype path of phASER.
 *
 * This is the drop-in boundary.  The reference (secastel/phaser) has no FFI of its own: it is pure
 * Python whose hot path is reached through (S1) the module function
 * read_variant_map.do_read_variant_map (phaser/read_variant_map.py:3) -- which its README tells
 * users to replace by a compiled read_variant_map.so (phaser/README.md:20-25) -- and (S2) the stage
 * functions of phaser/phaser.py that process_vcf maps over contigs / pairs / blocks
 * (phaser/phaser.py:442, 533, 556, 650, 680, 784, 808).  Each entry point below names the reference
 * code it replaces.  INTEGRATION.md shows the ctypes stubs a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only.  `d_` pointers are DEVICE pointers (sm_100a build) --
 * the caller owns them and keeps them alive until the stage that reads them has returned; `h_`
 * pointers are host pointers.  Every function returns 0 on success and a negative code on error;
 * phz_last_error() gives the message.  One context = one GPU = one CUDA stream; no internal threads.
 */
#ifndef PHZ_H
#define PHZ_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct phz_ctx phz_ctx;

#define PHZ_AS_BINS 65536          /* alignment-score histogram: bin = AS + 32768 */
#define PHZ_AS_NONE INT32_MIN      /* as_cutoff value meaning "no cutoff" */

/* Packed SoA of one BAM's records after the samtools-stage filters (phaser.py:505-513, 1346),
 * grouped by contig in VCF order.  Replaces the SAM text stream read by read_variant_map.py:25-35. */
typedef struct phz_reads {
  int64_t n_records;
  int64_t n_cigar_ops;             /* length of cigar[] */
  int64_t n_bases;                 /* length of qual[]; seq[] holds (n_bases+1)/2 bytes */
  const int64_t* h_contig_rec_off; /* HOST, n_contigs+1: records of contig c are [off[c], off[c+1]) */
  const int32_t* pos;              /* 1-based leftmost position (SAM POS) */
  const int32_t* tlen;             /* SAM TLEN */
  const int16_t* aln_score;        /* AS:i, -32768 when the tag is absent */
  const uint32_t* frag;            /* fragment id: one per distinct QNAME */
  const uint32_t* cigar_off;       /* n_records+1 */
  const uint32_t* cigar;           /* BAM encoding: len<<4 | op (MIDNSHP=X) */
  const uint64_t* seq_off;         /* n_records+1, in bases */
  const uint8_t* seq;              /* 4-bit BAM base codes, even base index = high nibble */
  const uint8_t* qual;             /* phred, one byte per base */
} phz_reads;

const char* phz_last_error(void);
/* "cuda-sm_100a" for the product library; the host logic-test double reports "hostsim". */
const char* phz_backend_name(void);

/* device: CUDA ordinal; stream: a cudaStream_t (0 = the legacy default stream). */
phz_ctx* phz_create(int device, void* stream);
void phz_destroy(phz_ctx* ctx);
int phz_sync(phz_ctx* ctx);

/* Het-site table of the sample, sorted by (contig, VCF order).  Replaces the per-contig mapping
 * tables written by generate_mapping_table (phaser.py:1355-1413) and parsed by `variant`
 * (read_variant_map.py:126-138).  a0/a1: 4-bit base codes of the two alleles the sample carries,
 * in allele-index order (phaser.py:1431-1435); 0xFF = cannot equal a read base. */
int phz_set_variants(phz_ctx* ctx, int n_contigs, const int64_t* h_contig_var_off, const int32_t* d_pos,
                     const uint8_t* d_a0, const uint8_t* d_a1, int64_t n_variants);

/* Optional, after phz_set_variants: --include_indels 1 (phaser.py:1398-1408).  Sites whose REF is longer than one
 * base or whose alleles are multi-base strings carry a0 = a1 = 0xFE in phz_set_variants and are resolved by
 * identify_allele's general rule (read_variant_map.py:236-258) against these tables: d_ref_len[V] = len(REF);
 * the two allele strings of site j are d_al_codes[d_al_off[2j] .. d_al_off[2j+1]) and [d_al_off[2j+1] ..
 * d_al_off[2j+2]), one 4-bit base code per character (0xFF = matches no read base).  NULL clears. */
int phz_set_indel_alleles(phz_ctx* ctx, const int32_t* d_ref_len, const uint32_t* d_al_off, const uint8_t* d_al_codes);

/* Optional, after phz_set_variants: one byte per het site, 1 = the site lies in a --haplo_count_blacklist
 * interval and is left out of the per-BAM haplotypic counts and read lists (phaser.py:1070, 1189).  NULL clears. */
int phz_set_haplo_blacklist(phz_ctx* ctx, const uint8_t* d_flags);

/* K1.  Replaces do_read_variant_map (read_variant_map.py:3-124) incl. split_read (:165-234) and
 * identify_allele (:236-258) for one BAM: emits, in (record, segment, variant) order, one tuple per
 * het SNV a spliced segment of a record covers.  All pointers of `reads` except h_contig_rec_off
 * are device pointers.  Tuples stay inside the context (arrays "t_rec", "t_var", "t_misc"). */
int phz_map_reads(phz_ctx* ctx, const phz_reads* reads, int baseq, double isize_cutoff, int64_t* n_candidates);
/* Same with HOST arrays (pinned memory recommended): copies them to the device on the context's
 * stream first.  This is the end-to-end entry point bench.py times. */
int phz_map_reads_host(phz_ctx* ctx, const phz_reads* host_reads, int baseq, double isize_cutoff, int64_t* n_candidates);

/* Lossless transport form of phz_reads for the host -> device copy (the copy is PCIe-bound: 146 bytes per 2x76 bp
 * record as plain SoA, ~41 packed).  Every field takes the narrowest coding that loses nothing:
 *   offsets      -> per-record counts (u8 or u16; scanned on the device), or nothing when all reads are equally long
 *   pos          -> u16 difference to the previous record; 65535 = look the (signed 32-bit) difference up in the
 *                   exception list (contig starts, long gaps); an inclusive scan on the device rebuilds pos
 *   tlen         -> i16; -32768 = look the value up in the exception list (pairs spanning long introns)
 *   aln_score    -> u8 index into the table of distinct scores when there are at most 256, else i16
 *   cigar        -> u16 index into the table of distinct len<<4|op words when there are at most 65536, else u32
 *   bases        -> 2 bits (A C G T = 0..3) plus a sparse exception list for every other 4-bit code
 *   qualities    -> index into the BAM's own table of distinct phred values (1, 2, 4 or 8 bits, whatever it needs)
 * Everything is expanded again on the device into the phz_reads layout K1 reads; nothing is thresholded or dropped. */
typedef struct phz_packed_reads {
  int64_t n_records;
  int64_t n_cigar_ops;
  int64_t n_bases;
  const int64_t* h_contig_rec_off; /* n_contigs+1 */
  const uint16_t* pos_delta;       /* pos[r] - pos[r-1] (pos[-1] = 0), 65535 = exception */
  int64_t n_pos_exc;
  const uint32_t* pos_exc_index;   /* ascending record index */
  const int32_t* pos_exc_delta;
  const int16_t* tlen16;           /* -32768 = exception */
  int64_t n_tlen_exc;
  const uint32_t* tlen_exc_index;
  const int32_t* tlen_exc_value;
  int32_t as_bits;                 /* 8: as_data = u8 indices into as_table; 16: as_data = i16 scores */
  const void* as_data;
  int16_t as_table[256];
  const uint32_t* frag;
  int32_t n_cigar_bits;            /* 8 or 16 */
  const void* n_cigar;             /* CIGAR ops per record */
  int32_t l_seq_const;             /* >= 0: every record has this many bases and l_seq is NULL */
  const uint16_t* l_seq;           /* bases per record */
  int32_t cigar_bits;              /* 16: cigar = u16 indices into cigar_table; 32: cigar = the words */
  const void* cigar;
  int32_t n_cigar_table;
  const uint32_t* cigar_table;
  const uint8_t* seq2;             /* (n_bases+3)/4 bytes; base i at bits 2*(i&3) of byte i>>2 */
  int64_t n_exceptions;            /* bases whose code is not A/C/G/T (N, IUPAC, '='), ascending base index */
  const uint64_t* exc_index;
  const uint8_t* exc_code;         /* the original 4-bit code */
  int32_t qual_bits;               /* 1, 2, 4 or 8 */
  uint8_t qual_table[256];         /* index -> phred */
  const uint8_t* qualp;            /* (n_bases*qual_bits+7)/8 bytes; base i at bit i*qual_bits, little-endian */
  /* fragment ids.  frag_bits 32: `frag` holds them.  frag_bits 16 (`frag` is NULL): ids numbered by first appearance
   * (what the ingest assigns) are implicit -- bit r of frag_first set = record r opens the next new id, i.e.
   * frag_base + (set bits before r); a record with a clear bit refers back to an open fragment (its mate, a
   * secondary line) with a 16-bit distance: id = frag_base + (set bits before r) - frag_back[k], k = the record's rank
   * among the clear bits.  Any record whose id does not follow from this is listed in the exception arrays (ascending
   * record index), which win.  Lossless for ANY id sequence; the packer takes it when it is the smaller form. */
  int32_t frag_bits;
  uint32_t frag_base;
  const uint32_t* frag_first;      /* (n_records+31)/32 words */
  int64_t n_frag_back;             /* records with a clear bit */
  const uint16_t* frag_back;
  int64_t n_frag_exc;
  const uint32_t* frag_exc_index;
  const uint32_t* frag_exc_value;
} phz_packed_reads;
typedef struct phz_packed_host phz_packed_host;
/* Packs HOST arrays on n_threads host threads.  page_locked != 0: the output lives in page-locked memory (worth its
 * allocation cost when the buffers are copied more than once or the copy must overlap kernels; a one-shot run is
 * faster from pageable memory).  NULL + phz_last_error() when a record has more than 65535 CIGAR ops or bases: use
 * phz_map_reads_host for such data. */
phz_packed_host* phz_pack_reads(const phz_reads* host_reads, int n_contigs, int n_threads, int page_locked);
int phz_packed_view(phz_packed_host* p, phz_packed_reads* out);
int64_t phz_packed_bytes(phz_packed_host* p);
void phz_packed_free(phz_packed_host* p);
/* phz_map_reads with the packed HOST form: copies it to the device on the context's stream, expands it there and
 * runs K1.  The end-to-end entry point bench.py times; the command line takes it with PHZ_PACK=1 (a one-shot run is
 * faster with plain arrays: packing costs more host time than it saves on the bus). */
int phz_map_reads_packed(phz_ctx* ctx, const phz_packed_reads* packed, int baseq, double isize_cutoff, int64_t* n_candidates);
/* Optional: starts the copy of a sample's packed buffers on the context's copy stream and returns at once.  A later
 * phz_map_reads_packed with the same buffers waits for that copy instead of issuing its own, so a loop over samples
 * (the GTEx-style batch) hides the path of sample i under the copy of sample i+1.  The context has two transport
 * slots: at most one prefetch may be outstanding while another sample is being mapped.  The host buffers must stay
 * untouched until the matching phz_map_reads_packed has returned. */
int phz_prefetch_packed(phz_ctx* ctx, const phz_packed_reads* packed);

/* Host -> device copy of a PAGEABLE host array through two page-locked staging buffers filled by n_threads host threads:
 * what the command line uses to put the ingest's arrays (150 bytes per record) on the device at PCIe speed. */
int phz_upload(phz_ctx* ctx, void* d_dst, const void* h_src, int64_t bytes, int n_threads);

/* Exact histogram of the alignment scores of the tuples the reference mapper would print; the host
 * derives numpy.percentile from it (phaser.py:545-553).  d_hist: PHZ_AS_BINS uint64 on the device. */
int phz_as_histogram(phz_ctx* ctx, uint64_t* d_hist);

/* Applies `int(AS) >= as_cutoff` (phaser.py:1304) and appends the surviving tuples of BAM
 * `bam_index` to the run-wide store, i.e. process_mapping_result + the merge loops
 * (phaser.py:1287-1328, 558-581).  d_frag: the BAM's fragment-id array (NULL after phz_map_reads_host). */
int phz_commit_bam(phz_ctx* ctx, int bam_index, int32_t as_cutoff, const uint32_t* d_frag, int64_t* n_kept);

/* Per-variant read lists and the noise sums (phaser.py:610-632): h_noise[2] receives
 * (base_match_count, base_mismatch_count).  Separate from phz_build_graph so that the host can compute the
 * critical values (which need only the noise level) while the graph is being built. */
int phz_variant_stats(phz_ctx* ctx, uint64_t* h_noise);
/* The same in two halves, so that the host does not wait between the commits and the graph stage: the first half queues the
 * work and the copy of the two sums to a page-locked slot, the second waits for exactly that copy -- from any host thread,
 * e.g. the one that computes the critical values while the caller's thread queues phz_build_graph. */
int phz_variant_stats_async(phz_ctx* ctx);
int phz_noise_wait(phz_ctx* ctx, uint64_t* h_noise);
/* Contig-sharded runs: the two sums must be added over the ranks before anybody uses them (phaser.py:610-631 is global).
 * phz_variant_stats_device leaves this rank's sums in the caller's DEVICE buffer (uint64[2]; no wait, no host copy), the
 * caller all-reduces that buffer in stream order (NCCL on the context's stream: the collective runs between this stage's
 * kernels and the graph stage's, never beside them), and phz_noise_publish queues its copy to the slot phz_noise_wait
 * reads.  No host thread takes part in the exchange. */
int phz_variant_stats_device(phz_ctx* ctx, uint64_t* d_noise);
int phz_noise_publish(phz_ctx* ctx, const uint64_t* d_noise);

/* Unique read sets, generate_connectivity_map (phaser.py:1265-1285), pair enumeration (:667-678) and the
 * count part of test_variant_connection (:1594-1642). */
int phz_build_graph(phz_ctx* ctx, uint64_t n_fragments, uint64_t bam_exclude_mask, int64_t* n_edges,
                    uint32_t* max_c_total);

/* Edge drop (phaser.py:696-707) through the integer critical values h_kstar[c_total] computed on the
 * host with scipy (so the binomial test of phaser.py:1649 agrees bit for bit), build_haplotypes
 * (:1861-1882), phase_v3 (:2107-2324), block statistics and haplotypic counts (:876-931, 1048-1095).
 * status_flags: bit0 = a sub-block larger than 24 variants needed exhaustive phasing (unsupported),
 * bit1 = split_by_weak cannot reach max_block_size (the reference would not terminate), bit2 = a c_total had no
 * critical value. */
/* Optional, before phz_phase: critical values of the (few, distinct) c_total values beyond the dense table, as two
 * host arrays with ascending totals.  With it h_kstar only has to cover the small totals; a c_total found in neither
 * raises status bit 2.  Used once, by the next phz_phase. */
int phz_set_big_critical_values(phz_ctx* ctx, const uint32_t* h_n, const uint32_t* h_k, int64_t count);
int phz_phase(phz_ctx* ctx, const uint32_t* h_kstar, int64_t kstar_len, int max_block_size,
              uint64_t bam_exclude_mask, int64_t* n_final_blocks, int* status_flags);

/* The per-variant read lists behind the aReads/bReads columns (phaser.py:1105-1115). */
int phz_read_lists(phz_ctx* ctx, uint64_t bam_exclude_mask, int64_t* n_entries);

/* Result / intermediate arrays by name (DESIGN.md lists them). */
int phz_array(phz_ctx* ctx, const char* name, const void** d_ptr, int64_t* count, int* elem_bytes);
int phz_download(phz_ctx* ctx, const char* name, void* h_dst, int64_t dst_bytes);
/* Same without the final wait: enqueues the copy on the context's stream (h_dst should be page-locked, otherwise
 * the copy is synchronous anyway); phz_sync() makes the bytes visible.  Lets the caller fetch all result arrays
 * of a run with one wait. */
int phz_download_async(phz_ctx* ctx, const char* name, void* h_dst, int64_t dst_bytes);
/* Same, but `dst` may be a DEVICE pointer as well (the copy direction follows the pointer): packs result arrays into
 * one device buffer that a collective then moves to the rank that writes the files (the contig-sharded run's gather of
 * formatted rows / per-variant annotations; phaser.py:863-867 needs them in one place). */
int phz_copy_array(phz_ctx* ctx, const char* name, void* dst, int64_t dst_bytes);
/* Merge step of the contig-sharded run on the rank that writes the files (the per-variant read lists
 * of phaser.py:1105-1115 brought into the run-wide block order of phaser.py:863-867).  The ranks' read lists arrive
 * concatenated (n entries); consecutive entries of one row form a run.  Run k covers entries [d_run_first[k],
 * d_run_first[k+1]) (d_run_first has n_runs + 1 elements, the last one = n), goes to d_run_dest[k] .. in the merged
 * order, takes row id d_run_row[k], and its site ids (local to the sending rank) are turned into the sample's ids
 * through d_site_map[local id + d_run_site_base[k]].  All pointers are DEVICE pointers of this context's device; one
 * kernel on the context's stream, no wait. */
int phz_expand_runs(phz_ctx* ctx, int64_t n_runs, const int64_t* d_run_first, const int64_t* d_run_dest,
                    const int64_t* d_run_row, const int64_t* d_run_site_base, int64_t n, const uint32_t* d_site_local,
                    const uint32_t* d_frag, const int64_t* d_site_map, uint32_t* d_out_row, uint32_t* d_out_site,
                    uint32_t* d_out_frag);
/* counters[16]: n_tuples, entries, groups, pairs, distinct pairs, edges, dropped, members, blocks,
 * hard blocks, final blocks, read-list entries, n_candidates, n_bams, fragment runs re-sorted in place by the
 * graph stage, 1 if that stage fell back to the full-key sort */
int phz_counters(phz_ctx* ctx, int64_t* counters);
/* Tuning / A-B switches.  "k1_mode": 3 = tile kernel (het-site slab staged in shared memory by a TMA
 * bulk copy, CTA scan, one atomic cursor per tile, dense coalesced emission) + streaming permute into
 * canonical order (default); 2 = fused single pass with decoupled look-back; 1 = windowed two-pass
 * (count, scan, emit); 0 = generic two-pass (searches in global memory).  All four emit identical tuples. */
int phz_set_option(phz_ctx* ctx, const char* name, int64_t value);
/* Reads a switch back, or one of the read-only facts about the last run: "graph_ranked_in_commit" (1 when the graph
 * stage used the in-fragment ranks the commits took, see "n_fragments"), "pair_table_slots" (after any growth). */
int phz_get_option(phz_ctx* ctx, const char* name, int64_t* value);
/* CUDA-event timing of the K1 passes of the LAST phz_map_reads call on the context's stream:
 * ms[0] = count pass, ms[1] = scan + size readback, ms[2] = emit pass (-1 when profiling is off). */
int phz_set_profiling(phz_ctx* ctx, int level);   /* 0 off, 1 K1 pass events, 2 + named stage marks */
int phz_map_times(phz_ctx* ctx, float* ms);
/* "stage<TAB>milliseconds" lines for the stages run since the last report (profiling on); clears them. */
int phz_stage_report(phz_ctx* ctx, char* buf, int64_t buf_len);
/* kernels of this library launched so far / library (CUB) passes launched so far */
int phz_launch_counts(phz_ctx* ctx, uint64_t* own, uint64_t* library);
/* blocking waits of the host on the context's stream so far (device counters read back, result downloads, phz_sync) */
int phz_sync_count(phz_ctx* ctx, uint64_t* n);

/* ---- feature-level haplotypic counts: the join and the distinct-read counting of
 * phaser_gene_ae.py (phaser_gene_ae/phaser_gene_ae.py:95-101, 172-219).  All pointers are DEVICE pointers.  Rows =
 * the rows of haplotypic_counts.txt of ONE run (all BAMs); row_contig = index into the feature contigs, -1 when the
 * row takes no part (totalCount <= 0 or contig without features).  Read ids are renumbered densely per (row,
 * haplotype): the lists of variant v are ids[id_off[2v] .. id_off[2v+1]) (haplotype A) and ids[id_off[2v+1] ..
 * id_off[2v+2]) (B); row_ids_a/b = number of distinct ids of the row per haplotype.  Features are sorted by (contig,
 * start); f_maxstop = running maximum of f_stop inside the contig; f_contig_off has n_contigs+1 entries.
 * Result arrays (phz_array / phz_download): "ae_row", "ae_feat" (index into the sorted features), "ae_a", "ae_b" =
 * distinct reads per haplotype among the row's variants inside the feature, one entry per (row, overlapping feature)
 * in row order. */
typedef struct phz_ae_input {
  int64_t n_rows;
  const int32_t* row_contig;
  const int32_t* row_start;        /* 1-based, as printed */
  const int32_t* row_stop;
  const uint32_t* row_a;           /* aCount / bCount columns */
  const uint32_t* row_b;
  const uint32_t* row_ids_a;
  const uint32_t* row_ids_b;
  const uint32_t* var_off;         /* n_rows+1 */
  const int32_t* var_pos;          /* position field of the variant id */
  const uint32_t* id_off;          /* 2*n_vars+1 */
  const uint32_t* ids;
  int64_t n_features;
  const int32_t* f_start;          /* BED, 0-based */
  const int32_t* f_stop;
  const int32_t* f_maxstop;
  const int64_t* f_contig_off;
} phz_ae_input;
int phz_gene_ae_pairs(phz_ctx* ctx, const phz_ae_input* in, int64_t* n_pairs);

/* ---- native host ingest (no GPU involved): BAM (BGZF, parallel inflate) or SAM text -> phz_reads with
 * HOST pointers.  Replaces `samtools view -h BAM chr: | samtools view -Sh [-F 0x400] [-f 2] -q MAPQ`
 * (phaser.py:1346) and the mapper's per-line parsing (read_variant_map.py:25-64).  Records come out
 * grouped by contig in the order of `contigs` (the VCF's), file order inside a contig.  A fragment
 * dictionary gives one id per distinct QNAME and is shared by all BAMs of a run. */
typedef struct phz_fragdict phz_fragdict;
typedef struct phz_host_reads phz_host_reads;
phz_fragdict* phz_fragdict_create(void);
void phz_fragdict_destroy(phz_fragdict* d);
int64_t phz_fragdict_size(phz_fragdict* d);
int64_t phz_fragdict_name(phz_fragdict* d, int64_t id, char* buf, int64_t buflen);
/* bulk export / import of the dictionary in id order (names back to back, n+1 offsets): the SoA cache file keeps it
 * beside the arrays so that a later run skips the ingest ("parse once").  Import needs an empty dictionary. */
int64_t phz_fragdict_blob_bytes(phz_fragdict* d);
int phz_fragdict_export(phz_fragdict* d, char* blob, int64_t* off);
int phz_fragdict_import(phz_fragdict* d, const char* blob, const int64_t* off, int64_t n, int n_threads);
phz_host_reads* phz_read_alignments(const char* path, const char* const* contigs, int n_contigs, phz_fragdict* d,
                                    int remove_dups, int proper_pair, int min_mapq, int n_threads);
int phz_host_reads_view(phz_host_reads* r, phz_reads* out, int* sorted_by_coordinate);
void phz_host_reads_free(phz_host_reads* r);
/* ---- native host VCF ingest (no GPU involved).  phz_vcf_open inflates the file (gzip or BGZF, BGZF blocks in
 * parallel) and keeps the text; phz_vcf_parse applies, for one sample column, the reference's het-site filter:
 * `gunzip -c VCF | cut -f 1-9,<col> | grep -v '0|0\|1|1'` (phaser.py:205-225), the per-line het / PASS test
 * (:396-434) and the mapping-table rules incl. the indel exclusion (:1355-1413).  The table's pointers stay valid
 * until the next phz_vcf_parse or phz_vcf_close on the handle.  Text fields (ids, rsids, alleles, GT strings) are NOT
 * copied: var_line_off / var_line_len locate each site's line in the text returned by phz_vcf_text. */
typedef struct phz_vcf phz_vcf;
typedef struct phz_vcf_table {
  int64_t n_variants;
  int32_t n_contigs;
  const int64_t* contig_var_off;   /* n_contigs+1; contigs in order of first appearance (phaser.py:410-411, 437-438) */
  const int32_t* pos;              /* VCF POS */
  const uint8_t* a0;               /* 4-bit base codes of the sample's two alleles (allele-index order), 0xFF = no single */
  const uint8_t* a1;               /*   base, 0xFE = multi-base site resolved through phz_set_indel_alleles */
  const int32_t* ref_len;          /* len(REF) */
  const int64_t* var_line_off;     /* byte offset of the site's line in the text */
  const int32_t* var_line_len;     /* its length without the newline */
  const char* contig_names;        /* n_contigs NUL-terminated names, back to back (as written in the VCF) */
  int32_t n_seen;                  /* chromosomes that reached the contig-name check (phaser.py:404-408), in order */
  const char* seen_names;
  int64_t stats[4];                /* het sites used, filtered (not PASS), indels excluded, unphased among the kept */
} phz_vcf_table;
phz_vcf* phz_vcf_open(const char* path, int n_threads);
void phz_vcf_close(phz_vcf* v);
int phz_vcf_text(phz_vcf* v, const char** text, int64_t* n_bytes, int64_t* n_lines, int* has_carriage_returns);
/* first line containing "#CHR" (sample_column_map, phaser.py:2326-2342); off = -1 when there is none */
int phz_vcf_chrom_line(phz_vcf* v, int64_t* off, int64_t* len);
int phz_vcf_parse(phz_vcf* v, int sample_column, int pass_only, const char* chrom_of_interest, int include_indels,
                  int n_threads, phz_vcf_table* out);

/* Output VCF (write_vcf, phaser.py:1661-1845) from the text phz_vcf_open holds and the sites phz_vcf_parse numbered.
 * Per site: the final block it sits in (or -1), which allele index haplotype A carries, and the genome-wide phase of its
 * two alleles; per block: members (PB = their rsids), printed index (PI / PS), confidence string (PC), whether that
 * confidence reaches --gw_phase_vcf_min_confidence, maf string (PM).  ids_match = 0 when --chr_prefix renames the
 * contigs: the reference then matches no line (its ids carry the prefix, the lines do not).  Returns the text (valid
 * until the next write / close) and counts[2] = (unphased sites phased genome-wide, phases corrected). */
typedef struct phz_vcf_annot {
  int32_t gw_phase_vcf;            /* 0, 1 or 2 */
  int32_t ids_match;
  const char* chrom_of_interest;   /* "" = all */
  int64_t n_variants;
  const int32_t* v_block;          /* [V] row of the blk_* arrays, -1 = in no phased block */
  const uint8_t* v_hap;            /* [V] allele index on haplotype A */
  const int8_t* v_gw;              /* [V][2] genome-wide phase of allele 0 / 1: 0, 1, or -1 = none */
  int64_t n_blocks;
  const int64_t* blk_first;        /* [B] first member in blk_members */
  const int64_t* blk_len;          /* [B] */
  const int32_t* blk_members;      /* site indices, position order */
  const int32_t* blk_index;        /* [B] */
  const uint8_t* blk_confident;    /* [B] */
  const char* blk_stat;            /* B NUL-terminated strings, back to back */
  const char* blk_maf;             /* B NUL-terminated strings, back to back */
  const char* id_separator;        /* --id_separator and --chr_prefix: a member without an rsid is named by its id */
  const char* chr_prefix;
} phz_vcf_annot;
int phz_vcf_write(phz_vcf* v, const phz_vcf_annot* annot, int n_threads, const char** text, int64_t* n_bytes, int64_t* counts);
/* data lines of the last phz_vcf_write in file order: chromosome (index into n_names NUL-terminated names), 0-based
 * reference span [beg, end) as tabix indexes it, byte offset of the line in the text */
int phz_vcf_records(phz_vcf* v, int64_t* n, const int32_t** chrom, const int64_t** beg, const int64_t** end,
                    const int64_t** text_off, const char** names, int32_t* n_names);

/* `bgzip -f` + `tabix -f -p vcf [--csi]` of the text of the last phz_vcf_write (phaser.py:1847-1853): writes path and
 * path + ".tbi" (or ".csi"); BGZF blocks deflated on n_threads threads. */
int phz_vcf_save(phz_vcf* v, const char* path_vcf_gz, int csi, int n_threads);

/* Text-side view of `n` het sites (indices into the table of the last phz_vcf_parse) for the table writers -- what
 * generate_variant_dict keeps per variant (phaser/phaser.py:1418-1462).  One line per site, tab-separated: POS, ID, REF, ALT
 * as in the file, the two alleles the sample's genotype names in allele-index order, the two alleles in genotype order when
 * the genotype is phased ("-", "-" otherwise).  A site whose genotype is not two different single digits gets the line "?"
 * (the caller reads that line itself).  The text stays valid until the next call on the same thread. */
int phz_vcf_site_text(phz_vcf* vcf, const int64_t* sites, int64_t n, int n_threads, const char** text, int64_t* n_bytes);
/* aReads / bReads columns of haplotypic_counts.txt (phaser.py:1105-1115) for the requested rows: rl_* are the triples of
 * phz_read_lists on the HOST (sorted by row); row r is named by row_key[r] (the packed block / BAM / haplotype key of
 * "rl_row") and prints the variants row_vars[row_var_off[r] .. row_var_off[r+1]) in that order.  Reads are numbered by
 * first occurrence inside the row.  text / row_text_off (n_rows+1 offsets) stay valid until the next call on the thread. */
int phz_format_read_lists(int64_t n, const uint32_t* rl_row, const uint32_t* rl_var, const uint32_t* rl_frag, int64_t n_rows,
                          const uint32_t* row_key, const int64_t* row_var_off, const uint32_t* row_vars, int n_threads,
                          const char** text, const int64_t** row_text_off);

/* SAM text twin of generated records: input for the reference baseline (test / bench infrastructure). */
int phz_write_sam(const char* path, const char* const* contig_names, const int64_t* contig_lengths, int n_contigs, int64_t n,
                  const int64_t* contig, const int64_t* pos, const int64_t* tlen, const int64_t* flag, const int64_t* mapq,
                  const int64_t* aln, const int64_t* frag, const int64_t* ops, const int64_t* opl, int n_ops,
                  const uint8_t* bases, const uint8_t* qual, int read_len, const char* bam_name);

/* BAM (BGZF) twin of the same records: the product's input in the files-to-files benchmark (test / bench infrastructure). */
int phz_write_bam(const char* path, const char* const* contig_names, const int64_t* contig_lengths, int n_contigs, int64_t n,
                  const int64_t* contig, const int64_t* pos, const int64_t* tlen, const int64_t* flag, const int64_t* mapq,
                  const int64_t* aln, const int64_t* frag, const int64_t* ops, const int64_t* opl, int n_ops,
                  const uint8_t* bases, const uint8_t* qual, int read_len, const char* bam_name, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
