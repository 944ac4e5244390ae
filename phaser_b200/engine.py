"""ctypes binding of the C ABI (include/phz.h) + device-memory plumbing through torch.

The product library is phaser_b200/_phz.so (built by __graft_entry__.build() with nvcc for sm_100a).
There is NO CPU path: without that library, or without a CUDA device, Engine() raises.  The `lib`
argument exists for the build container's logic tests, which pass the host-simulation double built
from the same pipeline source (tests/hostsim); product code never passes it.
"""
import ctypes
import os
from ctypes import c_int, c_int32, c_int64, c_uint32, c_uint64, c_double, c_void_p, c_char_p, POINTER, byref

import numpy as np
import torch

from .layout import ReadBatch, VariantTable, indel_tables

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHZ_LIB") or os.path.join(HERE, "_phz.so")      # PHZ_LIB: tuning builds of the same sources
AS_BINS = 65536
AS_NONE = -(2 ** 31)


class PhzError(RuntimeError):
    pass


class phz_reads(ctypes.Structure):
    _fields_ = [("n_records", c_int64), ("n_cigar_ops", c_int64), ("n_bases", c_int64),
                ("h_contig_rec_off", c_void_p), ("pos", c_void_p), ("tlen", c_void_p), ("aln_score", c_void_p),
                ("frag", c_void_p), ("cigar_off", c_void_p), ("cigar", c_void_p), ("seq_off", c_void_p),
                ("seq", c_void_p), ("qual", c_void_p)]


class phz_packed_reads(ctypes.Structure):
    _fields_ = [("n_records", c_int64), ("n_cigar_ops", c_int64), ("n_bases", c_int64), ("h_contig_rec_off", c_void_p),
                ("pos_delta", c_void_p), ("n_pos_exc", c_int64), ("pos_exc_index", c_void_p), ("pos_exc_delta", c_void_p),
                ("tlen16", c_void_p), ("n_tlen_exc", c_int64), ("tlen_exc_index", c_void_p), ("tlen_exc_value", c_void_p),
                ("as_bits", c_int32), ("as_data", c_void_p), ("as_table", ctypes.c_int16 * 256), ("frag", c_void_p),
                ("n_cigar_bits", c_int32), ("n_cigar", c_void_p), ("l_seq_const", c_int32), ("l_seq", c_void_p),
                ("cigar_bits", c_int32), ("cigar", c_void_p), ("n_cigar_table", c_int32), ("cigar_table", c_void_p),
                ("seq2", c_void_p), ("n_exceptions", c_int64), ("exc_index", c_void_p), ("exc_code", c_void_p),
                ("qual_bits", c_int32), ("qual_table", ctypes.c_uint8 * 256), ("qualp", c_void_p),
                ("frag_bits", c_int32), ("frag_base", ctypes.c_uint32), ("frag_first", c_void_p), ("n_frag_back", c_int64),
                ("frag_back", c_void_p), ("n_frag_exc", c_int64), ("frag_exc_index", c_void_p), ("frag_exc_value", c_void_p)]


class phz_vcf_table(ctypes.Structure):
    _fields_ = [("n_variants", c_int64), ("n_contigs", c_int32), ("contig_var_off", c_void_p), ("pos", c_void_p), ("a0", c_void_p),
                ("a1", c_void_p), ("ref_len", c_void_p), ("var_line_off", c_void_p), ("var_line_len", c_void_p),
                ("contig_names", c_void_p), ("n_seen", c_int32), ("seen_names", c_void_p), ("stats", c_int64 * 4)]


class phz_vcf_annot(ctypes.Structure):
    _fields_ = [("gw_phase_vcf", c_int32), ("ids_match", c_int32), ("chrom_of_interest", c_char_p), ("n_variants", c_int64),
                ("v_block", c_void_p), ("v_hap", c_void_p), ("v_gw", c_void_p), ("n_blocks", c_int64), ("blk_first", c_void_p),
                ("blk_len", c_void_p), ("blk_members", c_void_p), ("blk_index", c_void_p), ("blk_confident", c_void_p),
                ("blk_stat", c_char_p), ("blk_maf", c_char_p), ("id_separator", c_char_p), ("chr_prefix", c_char_p)]


class phz_ae_input(ctypes.Structure):
    _fields_ = [("n_rows", c_int64), ("row_contig", c_void_p), ("row_start", c_void_p), ("row_stop", c_void_p),
                ("row_a", c_void_p), ("row_b", c_void_p), ("row_ids_a", c_void_p), ("row_ids_b", c_void_p),
                ("var_off", c_void_p), ("var_pos", c_void_p), ("id_off", c_void_p), ("ids", c_void_p),
                ("n_features", c_int64), ("f_start", c_void_p), ("f_stop", c_void_p), ("f_maxstop", c_void_p),
                ("f_contig_off", c_void_p)]


EXPORTS = ["phz_last_error", "phz_backend_name", "phz_create", "phz_destroy", "phz_sync", "phz_set_variants",
           "phz_map_reads", "phz_map_reads_host", "phz_as_histogram", "phz_commit_bam", "phz_variant_stats", "phz_build_graph",
           "phz_phase", "phz_read_lists", "phz_array", "phz_download", "phz_download_async", "phz_counters", "phz_launch_counts",
           "phz_set_profiling", "phz_map_times", "phz_stage_report", "phz_set_option", "phz_get_option",
           "phz_fragdict_create", "phz_fragdict_destroy", "phz_fragdict_size", "phz_fragdict_name", "phz_read_alignments",
           "phz_host_reads_view", "phz_host_reads_free", "phz_set_haplo_blacklist", "phz_write_sam",
           "phz_set_indel_alleles", "phz_pack_reads", "phz_packed_view", "phz_packed_bytes", "phz_packed_free",
           "phz_map_reads_packed", "phz_prefetch_packed", "phz_gene_ae_pairs", "phz_set_big_critical_values",
           "phz_copy_array", "phz_expand_runs", "phz_variant_stats_async", "phz_noise_wait", "phz_variant_stats_device", "phz_noise_publish", "phz_vcf_open", "phz_vcf_close", "phz_vcf_text", "phz_vcf_chrom_line", "phz_vcf_parse",
           "phz_vcf_write", "phz_vcf_records", "phz_write_bam", "phz_vcf_save", "phz_format_read_lists", "phz_vcf_site_text", "phz_upload", "phz_sync_count",
           "phz_fragdict_blob_bytes", "phz_fragdict_export", "phz_fragdict_import"]


def _declare(lib):
    lib.phz_last_error.restype = c_char_p
    lib.phz_backend_name.restype = c_char_p
    lib.phz_create.restype = c_void_p
    lib.phz_create.argtypes = [c_int, c_void_p]
    lib.phz_destroy.argtypes = [c_void_p]
    lib.phz_sync.argtypes = [c_void_p]
    lib.phz_set_variants.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64]
    lib.phz_map_reads.argtypes = [c_void_p, POINTER(phz_reads), c_int, c_double, POINTER(c_int64)]
    lib.phz_map_reads_host.argtypes = [c_void_p, POINTER(phz_reads), c_int, c_double, POINTER(c_int64)]
    lib.phz_as_histogram.argtypes = [c_void_p, c_void_p]
    lib.phz_commit_bam.argtypes = [c_void_p, c_int, c_int32, c_void_p, POINTER(c_int64)]
    lib.phz_variant_stats.argtypes = [c_void_p, POINTER(c_uint64)]
    lib.phz_variant_stats_async.argtypes = [c_void_p]
    lib.phz_noise_wait.argtypes = [c_void_p, POINTER(c_uint64)]
    lib.phz_variant_stats_device.argtypes = [c_void_p, c_void_p]
    lib.phz_noise_publish.argtypes = [c_void_p, c_void_p]
    lib.phz_build_graph.argtypes = [c_void_p, c_uint64, c_uint64, POINTER(c_int64), POINTER(c_uint32)]
    lib.phz_phase.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_uint64, POINTER(c_int64), POINTER(c_int)]
    lib.phz_read_lists.argtypes = [c_void_p, c_uint64, POINTER(c_int64)]
    lib.phz_array.argtypes = [c_void_p, c_char_p, POINTER(c_void_p), POINTER(c_int64), POINTER(c_int)]
    lib.phz_download.argtypes = [c_void_p, c_char_p, c_void_p, c_int64]
    lib.phz_download_async.argtypes = [c_void_p, c_char_p, c_void_p, c_int64]
    lib.phz_copy_array.argtypes = [c_void_p, c_char_p, c_void_p, c_int64]
    lib.phz_expand_runs.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p]
    lib.phz_vcf_open.restype = c_void_p
    lib.phz_vcf_open.argtypes = [c_char_p, c_int]
    lib.phz_vcf_close.argtypes = [c_void_p]
    lib.phz_vcf_text.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64), POINTER(c_int)]
    lib.phz_vcf_chrom_line.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64)]
    lib.phz_vcf_parse.argtypes = [c_void_p, c_int, c_int, c_char_p, c_int, c_int, POINTER(phz_vcf_table)]
    lib.phz_vcf_write.argtypes = [c_void_p, POINTER(phz_vcf_annot), c_int, POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64)]
    lib.phz_vcf_records.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                    POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int32)]
    lib.phz_vcf_save.argtypes = [c_void_p, c_char_p, c_int, c_int]
    lib.phz_upload.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int]
    lib.phz_sync_count.argtypes = [c_void_p, POINTER(c_uint64)]
    lib.phz_vcf_site_text.argtypes = [c_void_p, c_void_p, c_int64, c_int, POINTER(c_char_p), POINTER(c_int64)]
    lib.phz_format_read_lists.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int,
                                          POINTER(c_void_p), POINTER(c_void_p)]
    lib.phz_counters.argtypes = [c_void_p, POINTER(c_int64)]
    lib.phz_launch_counts.argtypes = [c_void_p, POINTER(c_uint64), POINTER(c_uint64)]
    lib.phz_set_profiling.argtypes = [c_void_p, c_int]
    lib.phz_map_times.argtypes = [c_void_p, POINTER(ctypes.c_float)]
    lib.phz_stage_report.argtypes = [c_void_p, c_char_p, c_int64]
    lib.phz_set_option.argtypes = [c_void_p, c_char_p, c_int64]
    lib.phz_get_option.argtypes = [c_void_p, c_char_p, POINTER(c_int64)]
    lib.phz_fragdict_create.restype = c_void_p
    lib.phz_fragdict_destroy.argtypes = [c_void_p]
    lib.phz_fragdict_size.restype = c_int64
    lib.phz_fragdict_size.argtypes = [c_void_p]
    lib.phz_fragdict_name.restype = c_int64
    lib.phz_fragdict_name.argtypes = [c_void_p, c_int64, c_char_p, c_int64]
    lib.phz_fragdict_blob_bytes.restype = c_int64
    lib.phz_fragdict_blob_bytes.argtypes = [c_void_p]
    lib.phz_fragdict_export.argtypes = [c_void_p, c_void_p, c_void_p]
    lib.phz_fragdict_import.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int]
    lib.phz_read_alignments.restype = c_void_p
    lib.phz_read_alignments.argtypes = [c_char_p, POINTER(c_char_p), c_int, c_void_p, c_int, c_int, c_int, c_int]
    lib.phz_host_reads_view.argtypes = [c_void_p, POINTER(phz_reads), POINTER(c_int)]
    lib.phz_host_reads_free.argtypes = [c_void_p]
    lib.phz_set_haplo_blacklist.argtypes = [c_void_p, c_void_p]
    lib.phz_set_indel_alleles.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    lib.phz_pack_reads.restype = c_void_p
    lib.phz_pack_reads.argtypes = [POINTER(phz_reads), c_int, c_int, c_int]
    lib.phz_packed_view.argtypes = [c_void_p, POINTER(phz_packed_reads)]
    lib.phz_packed_bytes.restype = c_int64
    lib.phz_packed_bytes.argtypes = [c_void_p]
    lib.phz_packed_free.argtypes = [c_void_p]
    lib.phz_map_reads_packed.argtypes = [c_void_p, POINTER(phz_packed_reads), c_int, c_double, POINTER(c_int64)]
    lib.phz_set_big_critical_values.argtypes = [c_void_p, c_void_p, c_void_p, c_int64]
    lib.phz_prefetch_packed.argtypes = [c_void_p, POINTER(phz_packed_reads)]
    lib.phz_gene_ae_pairs.argtypes = [c_void_p, POINTER(phz_ae_input), POINTER(c_int64)]
    return lib


def load_library(path=LIB_PATH):
    if not os.path.isfile(path):
        raise PhzError("CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a).  phaser_b200 has no CPU path." % path)
    return _declare(ctypes.CDLL(path))


_DT = {1: np.uint8, 2: np.int16, 4: np.uint32, 8: np.uint64}
COUNTER_NAMES = ["n_tuples", "entries", "groups", "pairs", "distinct_pairs", "edges", "dropped", "members", "blocks",
                 "hard_blocks", "final_blocks", "read_list_entries", "n_candidates", "n_bams", "frag_runs_resorted",
                 "full_sort_fallback"]


def _as_torch(a: np.ndarray, device, pin=False):
    """numpy -> torch on `device`; unsigned 32/64-bit arrays travel as same-width signed views."""
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    elif a.dtype == np.uint64:
        a = a.view(np.int64)
    import warnings
    with warnings.catch_warnings():          # memory-mapped cache arrays are read-only: the tensor is only ever copied from
        warnings.simplefilter("ignore", UserWarning)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if pin and torch.cuda.is_available():
        t = t.pin_memory()
    if device is None:
        return t
    if torch.device(device).type == "cpu":
        return t.clone()          # never alias caller memory (the native reader's buffers are freed with the batch)
    return t.to(device, non_blocking=False)


class NativeFragmentDictionary:
    """QNAME -> fragment id kept by the native reader (one namespace for all BAMs of a run)."""

    def __init__(self, lib=None):
        self.lib = lib if lib is not None else load_library()
        self.h = self.lib.phz_fragdict_create()

    def __len__(self):
        return int(self.lib.phz_fragdict_size(self.h))

    def export(self):
        """(names back to back as uint8, int64 offsets[n+1]) in id order"""
        n = len(self)
        blob = np.empty(max(1, int(self.lib.phz_fragdict_blob_bytes(self.h))), np.uint8); off = np.zeros(n + 1, np.int64)
        if self.lib.phz_fragdict_export(self.h, blob.ctypes.data, off.ctypes.data) != 0:
            raise PhzError(self.lib.phz_last_error().decode())
        return blob[:int(off[n])], off

    def load(self, blob, off, threads=0):
        """fill the (empty) dictionary with names in id order"""
        blob = np.ascontiguousarray(blob, np.uint8); off = np.ascontiguousarray(off, np.int64)
        if self.lib.phz_fragdict_import(self.h, blob.ctypes.data, off.ctypes.data, int(off.shape[0] - 1),
                                        int(threads or (os.cpu_count() or 1))) != 0:
            raise PhzError(self.lib.phz_last_error().decode())

    @property
    def names(self):
        buf = ctypes.create_string_buffer(4096)
        out = []
        for i in range(len(self)):
            n = self.lib.phz_fragdict_name(self.h, i, buf, len(buf))
            out.append(buf.raw[:n].decode())
        return out

    def __del__(self):
        try:
            if self.h:
                self.lib.phz_fragdict_destroy(self.h); self.h = None
        except Exception:
            pass


def write_sam_native(rec, contigs, path, bam_name="bam0", lib=None, bam=False, threads=0):
    """synth.make_reads records -> SAM text through the native writer (50x faster than the Python loop); bam=True writes
    the BGZF-compressed BAM twin of the same records instead (phz_write_bam)."""
    lib = lib if lib is not None else load_library()
    g = lambda k, dt=np.int64: np.ascontiguousarray(rec[k].cpu().numpy().astype(dt))
    c, pos, tl, fl, mq, aln, fr = (g(k) for k in ("contig", "pos", "tlen", "flag", "mapq", "aln", "frag"))
    ops, opl = g("ops"), g("opl"); bases = g("bases", np.uint8); qual = g("qual", np.uint8)
    names = (c_char_p * len(contigs))(*[x[0].encode() for x in contigs])
    lens = np.asarray([x[1] for x in contigs], np.int64)
    p = lambda a: a.ctypes.data_as(c_void_p)
    lib.phz_write_sam.argtypes = [c_char_p, POINTER(c_char_p), c_void_p, c_int, c_int64] + [c_void_p] * 9 + [c_int, c_void_p, c_void_p, c_int, c_char_p]
    lib.phz_write_bam.argtypes = lib.phz_write_sam.argtypes + [c_int]
    if bam:
        rc = lib.phz_write_bam(path.encode(), names, p(lens), len(contigs), pos.shape[0], p(c), p(pos), p(tl), p(fl), p(mq), p(aln),
                               p(fr), p(ops), p(opl), ops.shape[1], p(bases), p(qual), bases.shape[1], bam_name.encode(),
                               int(threads or (os.cpu_count() or 1)))
    else:
        rc = lib.phz_write_sam(path.encode(), names, p(lens), len(contigs), pos.shape[0], p(c), p(pos), p(tl), p(fl), p(mq), p(aln),
                               p(fr), p(ops), p(opl), ops.shape[1], p(bases), p(qual), bases.shape[1], bam_name.encode())
    if rc != 0:
        raise PhzError(lib.phz_last_error().decode())
    return path


class PackedReads:
    """One BAM in the packed transport form (include/phz.h: phz_packed_reads) in host memory owned by the native
    library.  Built once at ingest; phz_map_reads_packed copies it to the device and expands it there."""

    def __init__(self, lib, handle):
        self.lib = lib; self.h = handle
        self.view = phz_packed_reads()
        if lib.phz_packed_view(handle, byref(self.view)) != 0:
            raise PhzError(lib.phz_last_error().decode())
        self.nbytes = int(lib.phz_packed_bytes(handle))
        self.n_records = int(self.view.n_records)
        v = self.view
        self.qual_bits = int(v.qual_bits)
        self.n_exceptions = int(v.n_exceptions)
        self.coding = dict(qual_bits=int(v.qual_bits), base_exceptions=int(v.n_exceptions), as_bits=int(v.as_bits),
                           n_cigar_bits=int(v.n_cigar_bits), l_seq_const=int(v.l_seq_const), cigar_bits=int(v.cigar_bits),
                           cigar_table=int(v.n_cigar_table), pos_exceptions=int(v.n_pos_exc), tlen_exceptions=int(v.n_tlen_exc),
                           frag_bits=int(v.frag_bits), frag_exceptions=int(v.n_frag_exc))

    def __del__(self):
        try:
            if self.h:
                self.lib.phz_packed_free(self.h); self.h = None
        except Exception:
            pass


def pack_reads(reads, n_contigs, threads=0, lib=None, page_locked=False) -> PackedReads:
    """`reads`: ReadBatch or dict of host arrays (numpy / CPU tensors) in the phz_reads layout.  `page_locked`: put the
    packed buffers in page-locked memory (pays off when they are copied repeatedly; a one-shot run skips the cost)."""
    lib = lib if lib is not None else load_library()
    if isinstance(reads, ReadBatch):
        reads = dict(contig_rec_off=np.ascontiguousarray(reads.contig_rec_off, dtype=np.int64), pos=reads.pos, tlen=reads.tlen,
                     aln_score=reads.aln_score, frag=reads.frag, cigar_off=reads.cigar_off, cigar=reads.cigar,
                     seq_off=reads.seq_off, seq=reads.seq, qual=reads.qual)
    reads = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in reads.items()}
    r = Engine._reads_struct(reads)
    h = lib.phz_pack_reads(byref(r), int(n_contigs), int(threads or (os.cpu_count() or 1)), int(bool(page_locked)))
    if not h:
        raise PhzError(lib.phz_last_error().decode())
    return PackedReads(lib, h)


class _NativeReads:
    def __init__(self, lib, h):
        self.lib = lib; self.h = h

    def __del__(self):
        try:
            if self.h:
                self.lib.phz_host_reads_free(self.h); self.h = None
        except Exception:
            pass


def read_alignments_native(path, contigs, fragdict: NativeFragmentDictionary, remove_dups=True, proper_pair=True,
                           min_mapq=0, threads=0, lib=None) -> ReadBatch:
    """BAM (BGZF) or SAM text -> ReadBatch through the native reader (phz_read_alignments).  The arrays are
    views of native memory owned by the returned batch."""
    lib = lib if lib is not None else fragdict.lib
    arr = (c_char_p * len(contigs))(*[c.encode() for c in contigs])
    h = lib.phz_read_alignments(path.encode(), arr, len(contigs), fragdict.h, int(remove_dups), int(proper_pair),
                                int(min_mapq), int(threads or (os.cpu_count() or 1)))
    if not h:
        raise PhzError(lib.phz_last_error().decode())
    owner = _NativeReads(lib, h)
    v = phz_reads(); srt = c_int(0)
    if lib.phz_host_reads_view(h, byref(v), byref(srt)) != 0:
        raise PhzError(lib.phz_last_error().decode())
    if not srt.value:
        raise PhzError("%s: records are not sorted by coordinate" % path)

    def view(ptr, n, dt):
        if n == 0 or not ptr:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(ctypes.cast(ptr, POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,))

    R = v.n_records; nc = len(contigs)
    rb = ReadBatch(nc, view(v.h_contig_rec_off, nc + 1, np.int64), view(v.pos, R, np.int32), view(v.tlen, R, np.int32),
                   view(v.aln_score, R, np.int16), view(v.frag, R, np.uint32), view(v.cigar_off, R + 1, np.uint32),
                   view(v.cigar, v.n_cigar_ops, np.uint32), view(v.seq_off, R + 1, np.uint64),
                   view(v.seq, (v.n_bases + 1) // 2, np.uint8), view(v.qual, v.n_bases, np.uint8), None)
    rb._owner = owner
    return rb


class Engine:
    def __init__(self, device="cuda:0", lib=None):
        self.lib = lib if lib is not None else load_library()
        self.backend = self.lib.phz_backend_name().decode()
        self.device = torch.device(device)
        if self.backend != "hostsim":
            if self.device.type != "cuda" or not torch.cuda.is_available():
                raise PhzError("phaser_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
            torch.cuda.set_device(self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            index = self.device.index or 0
        else:
            if lib is None:
                raise PhzError("the host-simulation backend is a test double and must be passed explicitly")
            self.device = torch.device("cpu")
            stream = 0
            index = 0
        self.ctx = self.lib.phz_create(index, c_void_p(stream))
        if not self.ctx:
            raise PhzError(self.lib.phz_last_error().decode())
        self._keep = {}
        self.n_contigs = 0
        self._cur = None

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.phz_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PhzError(self.lib.phz_last_error().decode())

    # ------------------------------------------------------------------ stages
    def set_variants(self, vt: VariantTable):
        d = self.device
        if self._keep.get("vt_id") != id(vt):          # the table stays resident across calls
            self._keep["v"] = (_as_torch(vt.pos, d), _as_torch(vt.a0, d), _as_torch(vt.a1, d))
            self._keep["vt_id"] = id(vt); self._keep["vt"] = vt
            it = indel_tables(vt)                       # --include_indels: multi-base sites
            self._keep["indel"] = None if it is None else tuple(_as_torch(a, d) for a in it)
        off = np.ascontiguousarray(vt.contig_var_off, np.int64)
        self._keep["voff"] = off
        self.n_contigs = len(vt.contigs)
        p, a0, a1 = self._keep["v"]
        self._check(self.lib.phz_set_variants(self.ctx, self.n_contigs, off.ctypes.data, p.data_ptr(), a0.data_ptr(),
                                              a1.data_ptr(), vt.n_variants))
        it = self._keep.get("indel")
        if it is not None:
            self._check(self.lib.phz_set_indel_alleles(self.ctx, it[0].data_ptr(), it[1].data_ptr(), it[2].data_ptr()))
        bl = getattr(vt, "haplo_blacklisted", None)
        if bl is not None and bl.any():
            self._keep["vblack"] = _as_torch(np.ascontiguousarray(bl, np.uint8), d)
            self._check(self.lib.phz_set_haplo_blacklist(self.ctx, self._keep["vblack"].data_ptr()))

    def _upload(self, a: np.ndarray, threads=8):
        """numpy array (pageable host memory) -> device tensor through the library's page-locked staging ring"""
        if a.dtype == np.uint32:
            a = a.view(np.int32)
        elif a.dtype == np.uint64:
            a = a.view(np.int64)
        a = np.ascontiguousarray(a)
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, device=self.device)
        if a.nbytes:
            self._check(self.lib.phz_upload(self.ctx, t.data_ptr(), a.ctypes.data, a.nbytes, int(threads)))
        return t

    def upload_reads(self, batch: ReadBatch, staged=None):
        """ReadBatch (numpy) -> dict of device tensors (the 'inputs resident in HBM' form).  `staged` (default: on a
        CUDA engine for batches above 64 MB): the big arrays go through phz_upload instead of pageable copies."""
        d = self.device
        nbytes = batch.qual.nbytes + batch.seq.nbytes
        if staged is None:
            staged = self.backend != "hostsim" and nbytes > (64 << 20)
        if staged:
            up = self._upload
            return dict(contig_rec_off=np.array(batch.contig_rec_off, dtype=np.int64, copy=True),
                        pos=up(batch.pos), tlen=up(batch.tlen), aln_score=up(batch.aln_score), frag=up(batch.frag),
                        cigar_off=up(batch.cigar_off), cigar=up(batch.cigar), seq_off=up(batch.seq_off), seq=up(batch.seq),
                        qual=up(batch.qual))
        return dict(contig_rec_off=np.array(batch.contig_rec_off, dtype=np.int64, copy=True),
                    pos=_as_torch(batch.pos, d), tlen=_as_torch(batch.tlen, d), aln_score=_as_torch(batch.aln_score, d),
                    frag=_as_torch(batch.frag, d), cigar_off=_as_torch(batch.cigar_off, d), cigar=_as_torch(batch.cigar, d),
                    seq_off=_as_torch(batch.seq_off, d), seq=_as_torch(batch.seq, d), qual=_as_torch(batch.qual, d))

    @staticmethod
    def _reads_struct(t):
        def ptr(x):
            return x.data_ptr() if torch.is_tensor(x) else x.ctypes.data
        r = phz_reads()
        r.n_records = int(t["pos"].shape[0]); r.n_cigar_ops = int(t["cigar"].shape[0]); r.n_bases = int(t["qual"].shape[0])
        r.h_contig_rec_off = t["contig_rec_off"].ctypes.data
        for k in ("pos", "tlen", "aln_score", "frag", "cigar_off", "cigar", "seq_off", "seq", "qual"):
            setattr(r, k, ptr(t[k]))
        return r

    def map_reads(self, dev_reads, baseq, isize_cutoff):
        """dev_reads: dict from upload_reads (device tensors)."""
        self._cur = dev_reads
        r = self._reads_struct(dev_reads)
        n = c_int64(0)
        self._check(self.lib.phz_map_reads(self.ctx, byref(r), int(baseq), float(isize_cutoff), byref(n)))
        return n.value

    def map_reads_host(self, host_reads, baseq, isize_cutoff):
        """host_reads: dict of (pinned) CPU tensors / numpy arrays; H2D copies happen inside the call."""
        self._cur = None
        r = self._reads_struct(host_reads)
        n = c_int64(0)
        self._check(self.lib.phz_map_reads_host(self.ctx, byref(r), int(baseq), float(isize_cutoff), byref(n)))
        return n.value

    def map_reads_packed(self, packed: PackedReads, baseq, isize_cutoff):
        """packed host form -> device (copies + expansion inside the call) -> K1"""
        self._cur = None
        n = c_int64(0)
        self._check(self.lib.phz_map_reads_packed(self.ctx, byref(packed.view), int(baseq), float(isize_cutoff), byref(n)))
        return n.value

    def prefetch_packed(self, packed: PackedReads):
        """Start copying `packed` to the device in the background; the next map_reads_packed(packed) uses that copy."""
        self._check(self.lib.phz_prefetch_packed(self.ctx, byref(packed.view)))

    def as_histogram(self):
        h = torch.zeros(AS_BINS, dtype=torch.int64, device=self.device)
        self._check(self.lib.phz_as_histogram(self.ctx, h.data_ptr()))
        return h

    def commit_bam(self, bam_index, as_cutoff=None):
        n = c_int64(0)
        frag = self._cur["frag"].data_ptr() if self._cur is not None else None
        self._check(self.lib.phz_commit_bam(self.ctx, int(bam_index), AS_NONE if as_cutoff is None else int(as_cutoff),
                                            frag, byref(n)))
        return n.value

    def variant_stats(self):
        noise = (c_uint64 * 2)()
        self._check(self.lib.phz_variant_stats(self.ctx, noise))
        return int(noise[0]), int(noise[1])

    def variant_stats_async(self):
        """first half of variant_stats: queues the work, does not wait (noise_wait returns the two sums)"""
        self._check(self.lib.phz_variant_stats_async(self.ctx))

    def variant_stats_device(self, t):
        """variant_stats with the two sums left in `t` (int64[2] on the engine's device): no wait, no host copy"""
        assert t.dtype == torch.int64 and t.numel() == 2 and t.is_contiguous() and t.device.type == self.device.type
        self._check(self.lib.phz_variant_stats_device(self.ctx, t.data_ptr()))

    def noise_publish(self, t):
        """queues the copy of `t` (the sums, e.g. after an all-reduce) to the slot noise_wait reads"""
        self._check(self.lib.phz_noise_publish(self.ctx, t.data_ptr()))

    def noise_wait(self):
        noise = (c_uint64 * 2)()
        self._check(self.lib.phz_noise_wait(self.ctx, noise))
        return int(noise[0]), int(noise[1])

    def build_graph(self, n_fragments, exclude_mask=0):
        e = c_int64(0); mt = c_uint32(0)
        self._check(self.lib.phz_build_graph(self.ctx, int(n_fragments), int(exclude_mask), byref(e), byref(mt)))
        return e.value, mt.value

    def phase(self, kstar: np.ndarray, max_block_size, exclude_mask=0, big=None):
        """`kstar`: dense critical values for c_total < len(kstar); `big` = (totals ascending, values) for the rest."""
        if big is not None and len(big[0]):
            bn = np.ascontiguousarray(big[0], np.uint32); bk = np.ascontiguousarray(big[1], np.uint32)
            self._check(self.lib.phz_set_big_critical_values(self.ctx, bn.ctypes.data, bk.ctypes.data, bn.shape[0]))
        k = np.ascontiguousarray(kstar, np.uint32)
        nf = c_int64(0); fl = c_int(0)
        self._check(self.lib.phz_phase(self.ctx, k.ctypes.data, k.shape[0], int(max_block_size), int(exclude_mask),
                                       byref(nf), byref(fl)))
        return nf.value, fl.value

    def read_lists(self, exclude_mask=0):
        n = c_int64(0)
        self._check(self.lib.phz_read_lists(self.ctx, int(exclude_mask), byref(n)))
        return n.value

    # ------------------------------------------------------------------ results
    def download(self, name, dtype=None):
        p = c_void_p(); n = c_int64(0); eb = c_int(0)
        self._check(self.lib.phz_array(self.ctx, name.encode(), byref(p), byref(n), byref(eb)))
        out = np.empty(n.value, dtype or _DT[eb.value])
        self._check(self.lib.phz_download(self.ctx, name.encode(), out.ctypes.data, out.nbytes))
        return out

    def download_many(self, names):
        """All `names` with ONE wait: copies are enqueued into a page-locked result buffer owned by the engine
        (grow-only), then the stream is synchronised once.  The returned arrays are VIEWS of that buffer: they are
        overwritten by the next download_many on this engine."""
        info = []
        total = 0
        for name in names:
            p = c_void_p(); n = c_int64(0); eb = c_int(0)
            self._check(self.lib.phz_array(self.ctx, name.encode(), byref(p), byref(n), byref(eb)))
            off = (total + 63) // 64 * 64
            info.append((name, off, n.value, eb.value))
            total = off + n.value * eb.value
        buf = getattr(self, "_result_buf", None)
        if buf is None or buf.numel() < total:
            buf = torch.empty(int(total * 1.25) + 4096, dtype=torch.uint8)
            if self.device.type == "cuda":
                try:
                    buf = buf.pin_memory()
                except RuntimeError:
                    pass
            self._result_buf = buf
        base = buf.data_ptr()
        host = buf.numpy()
        out = {}
        for name, off, n, eb in info:
            self._check(self.lib.phz_download_async(self.ctx, name.encode(), base + off, n * eb))
            out[name] = host[off:off + n * eb].view(_DT[eb])
        self.sync()
        return out

    def pack_arrays(self, names):
        """All `names` back to back (64-byte aligned) in ONE buffer that lives where the engine's arrays live (device
        memory in the product) -- the send buffer of the sharded run's result gather.  Returns (uint8 tensor,
        [(name, byte offset, count, element bytes)]).  The copies are enqueued on the engine's stream, no wait."""
        info = []
        total = 0
        for name in names:
            p = c_void_p(); n = c_int64(0); eb = c_int(0)
            self._check(self.lib.phz_array(self.ctx, name.encode(), byref(p), byref(n), byref(eb)))
            off = (total + 63) // 64 * 64
            info.append((name, off, n.value, eb.value))
            total = off + n.value * eb.value
        buf = torch.empty(max(total, 64), dtype=torch.uint8, device=self.device)
        base = buf.data_ptr()
        for name, off, n, eb in info:
            self._check(self.lib.phz_copy_array(self.ctx, name.encode(), base + off, n * eb))
        return buf, info

    def expand_runs(self, run_first, run_dest, run_row, run_site_base, site_local, frag, site_map, out_row, out_site, out_frag):
        """phz_expand_runs on tensors that live on the engine's device (int64 run tables, int32 entry arrays)."""
        for t in (run_first, run_dest, run_row, run_site_base, site_map):
            assert t.dtype == torch.int64 and t.is_contiguous() and t.device.type == self.device.type
        for t in (site_local, frag, out_row, out_site, out_frag):
            assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == self.device.type
        self._check(self.lib.phz_expand_runs(self.ctx, int(run_row.shape[0]), run_first.data_ptr(), run_dest.data_ptr(),
                                             run_row.data_ptr(), run_site_base.data_ptr(), int(site_local.shape[0]),
                                             site_local.data_ptr(), frag.data_ptr(), site_map.data_ptr(), out_row.data_ptr(),
                                             out_site.data_ptr(), out_frag.data_ptr()))

    def gene_ae_pairs(self, rows, feats):
        """phaser_gene_ae join + distinct-read counts (phz_gene_ae_pairs).  `rows` / `feats`: phaser_gene_ae.Rows /
        Features.  Returns (row, sorted-feature index, aCount, bCount) per (row, overlapping feature)."""
        d = self.device
        keep = []

        def up(a, dt):
            t = _as_torch(np.ascontiguousarray(np.asarray(a, dt)), d)
            keep.append(t)
            return t.data_ptr()
        s = phz_ae_input()
        s.n_rows = rows.n
        s.row_contig = up(rows.contig, np.int32); s.row_start = up(rows.start, np.int32); s.row_stop = up(rows.stop, np.int32)
        s.row_a = up(rows.a, np.uint32); s.row_b = up(rows.b, np.uint32)
        s.row_ids_a = up(rows.ida, np.uint32); s.row_ids_b = up(rows.idb, np.uint32)
        s.var_off = up(rows.var_off, np.uint32); s.var_pos = up(rows.var_pos, np.int32)
        s.id_off = up(rows.id_off, np.uint32); s.ids = up(rows.ids, np.uint32)
        s.n_features = int(feats.f_start.shape[0])
        s.f_start = up(feats.f_start, np.int32); s.f_stop = up(feats.f_stop, np.int32); s.f_maxstop = up(feats.f_maxstop, np.int32)
        s.f_contig_off = up(feats.f_contig_off, np.int64)
        n = c_int64(0)
        self._check(self.lib.phz_gene_ae_pairs(self.ctx, byref(s), byref(n)))
        out = tuple(self.download(k) for k in ("ae_row", "ae_feat", "ae_a", "ae_b"))
        del keep
        return out

    def counters(self):
        c = (c_int64 * 16)()
        self._check(self.lib.phz_counters(self.ctx, c))
        return {k: int(c[i]) for i, k in enumerate(COUNTER_NAMES)}

    def launch_counts(self):
        a = c_uint64(0); b = c_uint64(0)
        self._check(self.lib.phz_launch_counts(self.ctx, byref(a), byref(b)))
        return a.value, b.value

    def sync_count(self):
        n = c_uint64(0)
        self._check(self.lib.phz_sync_count(self.ctx, byref(n)))
        return n.value

    def set_profiling(self, level=1):
        """0 off, 1 CUDA events around the K1 passes, 2 additionally named stage marks."""
        self._check(self.lib.phz_set_profiling(self.ctx, int(level)))

    def map_times(self):
        """(count pass, scan + readback, emit pass) in ms for the last map_reads call; CUDA events."""
        ms = (ctypes.c_float * 3)()
        self._check(self.lib.phz_map_times(self.ctx, ms))
        return float(ms[0]), float(ms[1]), float(ms[2])

    def set_option(self, name, value):
        self._check(self.lib.phz_set_option(self.ctx, name.encode(), int(value)))

    def get_option(self, name):
        v = c_int64()
        self._check(self.lib.phz_get_option(self.ctx, name.encode(), byref(v)))
        return int(v.value)

    def stage_report(self):
        """{stage: ms} accumulated since the last call (CUDA events; needs set_profiling(True))."""
        buf = ctypes.create_string_buffer(1 << 16)
        self._check(self.lib.phz_stage_report(self.ctx, buf, len(buf)))
        out = {}
        for ln in buf.value.decode().splitlines():
            k, v = ln.split("\t")
            if k.endswith(".end"):
                k = "host_gap_after." + k[:-4]      # time until the next API call: host-side work
            out[k] = out.get(k, 0.0) + float(v)
        return out

    def sync(self):
        self._check(self.lib.phz_sync(self.ctx))
