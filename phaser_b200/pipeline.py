"""Host driver of the device pipeline: the stages of process_vcf between "het sites loaded" and
"write the files" (phaser/phaser.py:463-831), with the three tiny floating-point pieces kept on the
host so that they agree bit for bit with the reference's numpy / scipy calls:

  * alignment-score cutoff   numpy.percentile semantics from an exact histogram  (phaser.py:545-553)
  * noise level              one float64 division                                  (phaser.py:631)
  * conflicting-config test  integer critical values from scipy's binom            (phaser.py:1649, 696)

Multi-GPU: contigs are sharded over ranks (shard.py); `comm` carries the three exact reductions
(histogram, noise sums, max c_total).  Everything else is rank-local.
"""
import math
import os
import sys
import threading
import time
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .engine import Engine, AS_BINS, PackedReads
from .layout import ReadBatch, VariantTable, AS_MISSING
from .vcfio import PhaserFatal


@dataclass
class PhaseParams:
    baseq: int = 10
    isize: List[float] = field(default_factory=lambda: [0.0])
    as_q_cutoff: float = 0.05
    cc_threshold: float = 0.01
    max_block_size: int = 15
    haplo_count_bam_exclude: List[int] = field(default_factory=list)   # 0-based
    want_read_lists: bool = True
    want_read_ids: bool = False        # --output_read_ids 1: kept tuples of un-blocked variants come back too
    want_kept_tuples: bool = False     # --output_network: all kept (fragment, variant, class|bam) tuples come back

    def exclude_mask(self):
        m = 0
        for b in self.haplo_count_bam_exclude:
            m |= 1 << b
        return m


_TRACE = bool(os.environ.get("PHZ_TRACE"))
PRECOMPUTED_TOTALS = 2048      # c_total values whose critical value is computed while the graph is being built


class NullComm:
    """Single-GPU stand-in for the cross-rank reductions."""
    world_size = 1
    rank = 0

    def allreduce_sum(self, t):
        return t

    def allreduce_max_int(self, x):
        return x

    def allreduce_sum_ints(self, xs):
        return list(xs)


def percentile_from_histogram(hist: np.ndarray, q_fraction: float):
    """numpy.percentile(values, q_fraction*100) (default 'linear' method) where `values` are the
    integers whose exact histogram is `hist` (bin = value + 32768).  phaser.py:551."""
    n = int(hist.sum())
    if n == 0:
        return None
    q = np.true_divide(np.float64(q_fraction * 100), 100)
    virtual = (n - 1) * q
    prev = int(np.floor(virtual))
    gamma = np.float64(virtual - prev)
    nxt = min(prev + 1, n - 1)
    prev = max(min(prev, n - 1), 0)
    cum = np.cumsum(hist)
    a = np.float64(int(np.searchsorted(cum, prev + 1, side="left")) - 32768)
    b = np.float64(int(np.searchsorted(cum, nxt + 1, side="left")) - 32768)
    diff = b - a
    out = a + diff * gamma
    if gamma >= 0.5:
        out = b - diff * (1 - gamma)
    return float(out)


def noise_level(match: int, mismatch: int) -> float:
    """phaser.py:631"""
    return float(mismatch) / (float(match + mismatch) * 2)


_POOL_THREADS = 2
_POOL = None


def _pool():
    """resident helper thread of run_path (critical values under the graph stage; splitting the large totals over
    several threads was measured slower: the numpy calls around the ufunc serialise on the GIL)"""
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=_POOL_THREADS, thread_name_prefix="phz-host")
    return _POOL


def critical_values(max_total: int, noise_e: float, cc_threshold: float, totals=None) -> np.ndarray:
    """kstar[n] = min{k : binom.cdf(k, n, p) >= cc_threshold}, p = 1-(6e+10e^2)  (phaser.py:1649, 696).
    An edge with 0 < c_supporting < c_total is dropped iff c_supporting < kstar[c_total]; uses the same
    scipy function as the reference so the comparison agrees with it exactly.  With `totals` (the
    c_total values that actually occur) only those entries of the table are computed."""
    p = 1 - ((6 * noise_e) + (10 * math.pow(noise_e, 2)))
    if totals is not None:
        t = np.asarray(totals)
        n = np.flatnonzero(np.bincount(t.astype(np.int64), minlength=1)) if t.shape[0] else np.zeros(0, np.int64)
        table = np.zeros(max_total + 1, np.uint32)
        if n.shape[0]:
            table[n] = _critical_values_at(n.astype(np.int64), p, cc_threshold)
        return table
    n = np.arange(0, max_total + 1, dtype=np.int64)
    return _critical_values_at(n, p, cc_threshold)


def _binom_cdf(k, n, p):
    """binom.cdf for 0 <= k <= n: the same Boost-backed ufunc scipy's public method ends in
    (scipy/stats/_discrete_distns.py: binom._cdf), without the per-call argument plumbing."""
    from scipy.stats import binom
    try:
        out = binom._cdf(k.astype(np.float64), n, p)
        return np.where(k >= n, 1.0, out)
    except Exception:
        return binom.cdf(k, n, p)


def _scipy_ppf(q, n, p):
    from scipy.stats import binom
    try:
        return binom._ppf(q, n, p)
    except Exception:
        return binom.ppf(q, n, p)


def _critical_values_at(n, p, cc_threshold):
    """min{k : binom.cdf(k, n, p) >= cc_threshold} for every n.  Starts from the Cornish-Fisher quantile (exact for
    ~98 % of the totals, one below otherwise) instead of scipy's inverse (an iterative root search per element), then
    walks each element to the k with cdf(k) >= threshold > cdf(k-1) using the SAME cdf the reference's test uses,
    re-evaluating only the elements that moved: the result does not depend on the starting point."""
    from scipy.special import ndtri
    n = np.asarray(n, np.int64)
    if n.shape[0] == 0:
        return np.zeros(0, np.uint32)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = float(ndtri(cc_threshold)) if 0.0 < cc_threshold < 1.0 else 0.0
        sd = np.sqrt(n * p * (1 - p))
        w = z + (z * z - 1) * ((1 - 2 * p) / sd) / 6
        k = np.ceil(n * p + w * sd - 0.5)
    bad = ~np.isfinite(k)
    if bad.any():                      # degenerate spread (p = 0 or 1, n = 0): scipy's own inverse as the start
        k[bad] = _scipy_ppf(cc_threshold, n[bad], p)
    k = np.clip(np.nan_to_num(k, nan=0.0, posinf=0.0, neginf=0.0).astype(np.int64), 0, n)
    act = np.arange(n.shape[0])
    for it in range(1 << 20):
        if it == 8:                    # the closed form was far off for these (tiny n): restart them from the inverse
            k[act] = np.clip(np.nan_to_num(_scipy_ppf(cc_threshold, n[act], p), nan=0.0).astype(np.int64), 0, n[act])
        ka = k[act]; na = n[act]
        low = (_binom_cdf(ka, na, p) < cc_threshold) & (ka < na)             # k too small
        ka = np.where(low, ka + 1, ka)
        km = np.maximum(ka - 1, 0)
        high = ~low & (ka > 0) & (_binom_cdf(km, na, p) >= cc_threshold)      # k-1 already passes
        ka = np.where(high, km, ka)
        k[act] = ka
        moved = low | high
        if not moved.any():
            break
        act = act[moved]
    # cdf(n; n, p) = 1 >= threshold always, so k <= n
    return k.astype(np.uint32)


def edge_pvalues(sup: np.ndarray, tot: np.ndarray, noise_e: float):
    """conflicting_config_p per edge as the reference prints it (phaser.py:1645-1652): python int 0 / 1 or
    numpy float64 from binom.cdf."""
    from scipy.stats import binom
    p = 1 - ((6 * noise_e) + (10 * math.pow(noise_e, 2)))
    out = np.empty(sup.shape[0], dtype=object)
    test = (sup > 0) & (tot > sup)
    out[sup == 0] = 0
    out[(sup > 0) & ~test] = 1
    if test.any():
        pairs = np.stack([sup[test].astype(np.int64), tot[test].astype(np.int64)], 1)
        uniq, inv = np.unique(pairs, axis=0, return_inverse=True)
        vals = binom.cdf(uniq[:, 0], uniq[:, 1], p)
        res = vals[inv.reshape(-1)]
        idx = np.nonzero(test)[0]
        for i, v in zip(idx.tolist(), res):
            out[i] = v
    return out


@dataclass
class PhaseResult:
    n_bams: int
    as_cutoff: List[Optional[float]]
    tuples_per_bam: List[int]
    candidates_per_bam: List[int]
    noise_e: float
    match: int
    mismatch: int
    counters: dict
    status_flags: int
    arrays: dict

    def __getattr__(self, k):
        a = self.__dict__.get("arrays", {})
        if k in a:
            return a[k]
        raise AttributeError(k)


RESULT_ARRAYS = ["vfirst", "ncls", "setsize", "vb_cnt", "ed_a", "ed_b", "ed_sup", "ed_tot", "ed_cfg", "ed_keep",
                 "members", "fb_first", "fb_len", "fb_sup", "fb_tot", "fb_cnt", "fb_bcnt", "v_final", "v_hap"]


def run_path(engine: Engine, vt: VariantTable, batches, params: PhaseParams, n_fragments: int, comm=None,
             host_inputs=False, download=True, reuse_result_buffer=False) -> PhaseResult:
    """`batches`: per BAM a dict of device tensors (Engine.upload_reads), a PackedReads (host, packed transport
    form: phz_map_reads_packed copies and expands it) or, with host_inputs, a dict of plain host arrays
    (phz_map_reads_host copies them).  `reuse_result_buffer`: the result arrays are views of the engine's
    page-locked result buffer (one wait for all of them, no pageable copies) and stay valid until the next run on
    this engine; otherwise they are private copies."""
    comm = comm or NullComm()
    nb = len(batches)
    isz = list(params.isize) * nb if len(params.isize) == 1 else list(params.isize)
    excl = params.exclude_mask()
    engine.set_variants(vt)
    engine.set_option("n_fragments", int(n_fragments))      # lets the commits rank the tuples inside their fragments
    cutoffs, kept, cands = [], [], []
    for b, reads in enumerate(batches):
        if isinstance(reads, PackedReads):          # packed transport form in host memory
            n_cand = engine.map_reads_packed(reads, params.baseq, isz[b])
        elif host_inputs:
            n_cand = engine.map_reads_host(reads, params.baseq, isz[b])
        else:
            n_cand = engine.map_reads(reads, params.baseq, isz[b])
        cands.append(n_cand)
        cutoff = None
        if params.as_q_cutoff > 0:
            hist = comm.allreduce_sum(engine.as_histogram()).cpu().numpy()
            n_missing = int(hist[AS_MISSING + 32768]); hist[AS_MISSING + 32768] = 0
            cutoff = percentile_from_histogram(hist, params.as_q_cutoff)     # None: no AS values at all
            if cutoff is not None and n_missing > 0:
                raise PhaserFatal("%d mapped reads carry no AS:i tag but an alignment-score cutoff is active "
                                  "(the reference fails on int('') at phaser.py:1304); use --as_q_cutoff 0" % n_missing)
        cutoffs.append(cutoff)
        kept.append(engine.commit_bam(b, None if cutoff is None else int(math.ceil(cutoff))))
    # The two noise sums are not needed by THIS thread before the graph stage is queued: variant_stats only queues its work
    # and the copy of the sums; the helper thread below waits for that copy (an event), so the device never idles here.
    # Several ranks: the sums are added over the ranks ON THE DEVICE, in stream order (comm.allreduce_sum_device: the
    # collective runs between this stage's kernels and the graph stage's; issued from the helper thread it had to find
    # room beside the graph kernels and came back a millisecond late).
    overlap_reduce = comm.world_size > 1 and getattr(comm, "timers", None) is None
    stream_reduce = comm.world_size > 1 and hasattr(comm, "allreduce_sum_device")
    async_noise = comm.world_size == 1 or overlap_reduce or stream_reduce
    if stream_reduce:
        import torch
        sums = torch.zeros(2, dtype=torch.int64, device=engine.device)
        engine.variant_stats_device(sums); comm.allreduce_sum_device(sums); engine.noise_publish(sums)
        overlap_reduce = False; match = mism = None
    elif async_noise:
        engine.variant_stats_async(); match = mism = None
    else:
        match, mism = engine.variant_stats()
    # The critical values of the small totals (the bulk of the edges) need only the noise level: a host thread computes
    # them while the GPU builds the graph (scipy's ufuncs and the C calls release the GIL).  With several ranks the same
    # thread first sums the two noise counters over the ranks, so that exchange is hidden under the graph stage as well
    # (no other collective is issued meanwhile: the order of collectives stays the same on every rank).
    pre = {}
    if not async_noise:
        match, mism = comm.allreduce_sum_ints([match, mism])

    def noise_and_critical_values(local=(match, mism)):
        try:
            w0 = time.perf_counter() if _TRACE else 0.0
            if async_noise:
                local = engine.noise_wait()
            w1 = time.perf_counter() if _TRACE else 0.0
            m_, x_ = comm.allreduce_sum_ints(list(local)) if overlap_reduce else local
            w2 = time.perf_counter() if _TRACE else 0.0
            pre["counts"] = (m_, x_)
            if m_ > 0:
                pre["k"] = critical_values(PRECOMPUTED_TOTALS, noise_level(m_, x_), params.cc_threshold)
            if _TRACE:
                print("[run_path helper] wait for the noise sums %.2f ms, all-reduce %.2f ms, critical values %.2f ms" % (
                    (w1 - w0) * 1e3, (w2 - w1) * 1e3, (time.perf_counter() - w2) * 1e3), file=sys.stderr)
        except BaseException as e:          # noqa: BLE001 -- re-raised on the main thread
            pre["error"] = e
    if comm.world_size == 1:
        wait_for_helper = _pool().submit(noise_and_critical_values).result      # a resident thread: no start-up cost on the critical path
    else:       # its own thread: the helper may take part in a collective, and logical shards (threads) must not queue behind each other
        helper = threading.Thread(target=noise_and_critical_values); helper.start()
        wait_for_helper = helper.join
    _t = time.perf_counter() if _TRACE else 0.0
    engine.set_option("big_total_threshold", PRECOMPUTED_TOTALS)
    try:
        n_edges, max_tot = engine.build_graph(n_fragments, excl)
    finally:
        _t1 = time.perf_counter() if _TRACE else 0.0
        wait_for_helper()
    if "error" in pre:
        raise pre["error"]
    match, mism = pre["counts"]
    if match == 0:
        raise PhaserFatal("No reads could be matched to variants. Please double check your settings and input files. "
                          "Common reasons for this occurring include: 1) MAPQ or BASEQ set too conservatively 2) BAM "
                          "and VCF have different chromosome names (IE 'chr1' vs '1').")
    noise_e = noise_level(match, mism)
    _t2 = time.perf_counter() if _TRACE else 0.0
    m = min(int(max_tot), PRECOMPUTED_TOTALS)
    kstar = pre["k"][:m + 1]
    n_big = 0; big_pair = None
    if max_tot > PRECOMPUTED_TOTALS and n_edges > 0:
        # the device kept the (few) totals above the precomputed range on a side list; their critical values go back as
        # a sparse (total, value) list, so nothing of size max_tot is ever built or copied
        big = np.unique(engine.download("big_tot")).astype(np.int64)
        n_big = int(big.shape[0])
        p = 1 - ((6 * noise_e) + (10 * math.pow(noise_e, 2)))
        big_pair = (big.astype(np.uint32), _critical_values_at(big, p, params.cc_threshold))
    if _TRACE:
        _t3 = time.perf_counter()
        print("[run_path] build_graph %.2f ms, join wait %.2f ms, kstar assembly %.2f ms (max_tot %d, %d distinct totals above %d)" % (
            (_t1 - _t) * 1e3, (_t2 - _t1) * 1e3, (_t3 - _t2) * 1e3, max_tot, n_big, PRECOMPUTED_TOTALS), file=sys.stderr)
    nf, flags = engine.phase(kstar, params.max_block_size, excl, big=big_pair)
    if flags & 4:
        raise PhaserFatal("internal: an edge total has no critical value")
    if flags & 2:
        raise PhaserFatal("a haplotype block cannot be split down to --max_block_size (the reference does not "
                          "terminate on this input)")
    if flags & 1:
        raise PhaserFatal("a sub-block of more than 24 variants needs exhaustive phasing; lower --max_block_size")
    arrays = {}
    if params.want_read_lists:
        engine.read_lists(excl)
    if download:
        names = RESULT_ARRAYS + (["rl_frag", "rl_var", "rl_row"] if params.want_read_lists else [])
        if reuse_result_buffer:
            arrays.update(engine.download_many(names))
        else:
            for name in names:
                arrays[name] = engine.download(name)
        if params.want_kept_tuples:
            for name in ("g_var", "g_cb", "g_frag"):
                arrays[name] = engine.download(name)
        if params.want_read_ids:
            # read_ids of singleton rows (phaser.py:1194-1218): the kept allele tuples of variants outside every block
            gv = engine.download("g_var"); gc = engine.download("g_cb")
            sel = np.nonzero((arrays["v_final"][gv] == 0xFFFFFFFF) & ((gc & 3) < 2))[0]
            arrays["sg_var"] = gv[sel]; arrays["sg_cb"] = gc[sel]; arrays["sg_frag"] = engine.download("g_frag")[sel]
    return PhaseResult(nb, cutoffs, kept, cands, noise_e, match, mism, engine.counters(), flags, arrays)
