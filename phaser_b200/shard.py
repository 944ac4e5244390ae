"""Multi-GPU: contigs are sharded over ranks, one process per GPU.

The reference itself maps, connects and builds blocks per contig (phaser/phaser.py:442, 533, 556, 650,
784; pairs are same-contig only, :1278-1280), so a contig never needs data of another one.  What IS
global, and therefore exchanged exactly between ranks, is small:
  * per BAM, the alignment-score histogram   -> all-reduce(sum)  -> same AS cutoff everywhere (phaser.py:545-553)
  * the two noise counters                   -> all-reduce(sum)  -> same noise_e / critical values (phaser.py:610-631)
  * the result arrays                        -> gather to rank 0 (one packed device buffer per rank, grouped
                                                send/recv), which renumbers blocks in the global output order
                                                (phaser.py:863-867) and writes the files.
Reads and tuples never leave their GPU.  Backend: NCCL on GPUs, gloo in the CPU tests.
"""
import os
import sys
import time
from typing import List

import numpy as np
import torch
import torch.distributed as dist

from .layout import ReadBatch, VariantTable
from .pipeline import PhaseParams, PhaseResult, run_path, RESULT_ARRAYS

NONE32 = 0xFFFFFFFF
_DT = {1: np.uint8, 2: np.int16, 4: np.uint32, 8: np.uint64}
N_COUNTERS = 16


def plan_shards(weights: List[int], world: int) -> List[List[int]]:
    """Longest-processing-time bin packing of contigs (weight = record count) onto `world` ranks."""
    order = sorted(range(len(weights)), key=lambda c: (-weights[c], c))
    loads = [0] * world
    out = [[] for _ in range(world)]
    for c in order:
        r = min(range(world), key=lambda i: (loads[i], i))
        out[r].append(c); loads[r] += weights[c]
    return [sorted(x) for x in out]


class DistComm:
    """The exact cross-rank reductions of pipeline.run_path over torch.distributed.  With `timers` (a dict) every
    collective is bracketed by device synchronisation and its wall time accumulated under its name."""

    def __init__(self, device, timers=None):
        self.rank = dist.get_rank(); self.world_size = dist.get_world_size(); self.device = torch.device(device)
        self.timers = timers
        self._side = None

    def _sync(self):
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def _timed(self, name, fn):
        if self.timers is None:
            return fn()
        self._sync(); t0 = time.perf_counter()
        out = fn()
        self._sync()
        self.timers[name] = self.timers.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return out

    def allreduce_sum(self, t):
        t = t.to(self.device)
        self._timed("allreduce_as_histogram_ms", lambda: dist.all_reduce(t, op=dist.ReduceOp.SUM))
        return t

    def allreduce_sum_device(self, t):
        """in-place sum over the ranks of a tensor that lives on the engine's device, in STREAM order: with NCCL the host does
        not wait and the collective runs where it is queued (between the kernels before and after it)"""
        if self.timers is not None:
            self._timed("allreduce_noise_ms", lambda: dist.all_reduce(t, op=dist.ReduceOp.SUM))
        elif t.device.type == self.device.type:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        else:           # gloo beside a CUDA engine (tests on a one-GPU box): through the collective's own device
            u = t.to(self.device); dist.all_reduce(u, op=dist.ReduceOp.SUM); t.copy_(u)
        return t

    def allreduce_sum_ints(self, xs):
        import threading
        if self.device.type == "cuda" and threading.current_thread() is not threading.main_thread():
            # called from the critical-value thread while the main thread queues the graph stage: a stream of its own, so
            # the exchange is ordered after nothing but its own upload (on the engine's stream it would queue behind
            # every graph kernel launched so far and the thread would wait for all of them)
            # -- and a HIGH-PRIORITY one: the collective's few CTAs must not wait until the graph kernels that fill every SM
            # have drained (measured at N = 2..8: the helper thread got its sums ~0.7 ms late and the critical values with them)
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device, priority=-1)
            with torch.cuda.stream(self._side):
                t = torch.tensor(list(xs), dtype=torch.int64, device=self.device)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                return [int(x) for x in t.cpu().tolist()]
        t = torch.tensor(list(xs), dtype=torch.int64, device=self.device)
        self._timed("allreduce_noise_ms", lambda: dist.all_reduce(t, op=dist.ReduceOp.SUM))
        return [int(x) for x in t.cpu().tolist()]

    def allreduce_max_int(self, x):
        t = torch.tensor([x], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())


def sub_variant_table(vt: VariantTable, contigs: List[int]):
    """Variant table restricted to `contigs` (kept in order) + the global ids of its variants."""
    idx = np.concatenate([np.arange(vt.contig_var_off[c], vt.contig_var_off[c + 1]) for c in contigs]) if contigs else \
        np.zeros(0, np.int64)
    off = np.zeros(len(contigs) + 1, np.int64)
    off[1:] = np.cumsum([vt.contig_var_off[c + 1] - vt.contig_var_off[c] for c in contigs])
    pick = (lambda lst: [lst[i] for i in idx.tolist()]) if vt.ids else (lambda lst: [])
    sub = VariantTable([vt.contigs[c] for c in contigs], off, vt.pos[idx], vt.a0[idx], vt.a1[idx], vt.ref_len[idx],
                       pick(vt.ids), pick(vt.rsids), pick(vt.all_alleles), pick(vt.gt), pick(vt.maf))
    if vt.haplo_blacklisted is not None:
        sub.haplo_blacklisted = np.ascontiguousarray(vt.haplo_blacklisted[idx])
    return sub, idx.astype(np.int64)


def sub_read_batch(rb: ReadBatch, contigs: List[int]) -> ReadBatch:
    """Records of `contigs` only (contig order kept); offsets rebased."""
    recs = np.concatenate([np.arange(rb.contig_rec_off[c], rb.contig_rec_off[c + 1]) for c in contigs]) if contigs else \
        np.zeros(0, np.int64)
    off = np.zeros(len(contigs) + 1, np.int64)
    off[1:] = np.cumsum([rb.contig_rec_off[c + 1] - rb.contig_rec_off[c] for c in contigs])
    ncig = (rb.cigar_off[1:].astype(np.int64) - rb.cigar_off[:-1].astype(np.int64))[recs]
    nbase = (rb.seq_off[1:].astype(np.int64) - rb.seq_off[:-1].astype(np.int64))[recs]
    cig_off = np.zeros(recs.shape[0] + 1, np.int64); cig_off[1:] = np.cumsum(ncig)
    seq_off = np.zeros(recs.shape[0] + 1, np.int64); seq_off[1:] = np.cumsum(nbase)

    def gather(starts, lens, total):
        if total == 0:
            return np.zeros(0, np.int64)
        rep = np.repeat(np.arange(recs.shape[0]), lens)
        first = np.cumsum(lens) - lens
        return starts[rep] + (np.arange(total) - first[rep])

    ci = gather(rb.cigar_off[:-1].astype(np.int64)[recs], ncig, int(cig_off[-1]))
    bi = gather(rb.seq_off[:-1].astype(np.int64)[recs], nbase, int(seq_off[-1]))
    codes = np.where(bi & 1, rb.seq[bi >> 1] & 15, rb.seq[bi >> 1] >> 4).astype(np.uint8)
    if codes.shape[0] & 1:
        codes = np.concatenate([codes, np.zeros(1, np.uint8)])
    seq = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)
    return ReadBatch(len(contigs), off, rb.pos[recs], rb.tlen[recs], rb.aln_score[recs], rb.frag[recs],
                     cig_off.astype(np.uint32), rb.cigar[ci], seq_off.astype(np.uint64), seq, rb.qual[bi], rb.qnames)


def sub_reads_tensors(reads: dict, contigs: List[int], dense_frag=False) -> dict:
    """sub_read_batch for the tensor form of a BAM (Engine.upload_reads: dict of torch tensors on any device +
    host `contig_rec_off`): the records of `contigs`, contig order kept, offsets rebased.  The records of a contig
    are one contiguous range of every array, so this is slicing and concatenation on the device the data is on.
    `dense_frag`: renumber the fragment ids of the shard densely (ascending global id = order of first appearance, so
    neighbours stay neighbours) and return the local -> global table as "frag_map": the graph stage's fragment table
    then has one slot per fragment of the SHARD instead of one per fragment of the sample."""
    cro = np.asarray(reads["contig_rec_off"], np.int64)
    pos = reads["pos"]; dev = pos.device
    coff = reads["cigar_off"]; soff = reads["seq_off"]
    rng = [(int(cro[c]), int(cro[c + 1])) for c in contigs]
    off = np.zeros(len(contigs) + 1, np.int64)
    off[1:] = np.cumsum([b - a for a, b in rng])

    def cut(t):
        return torch.cat([t[a:b] for a, b in rng]) if rng else t[:0]

    out = dict(contig_rec_off=off, pos=cut(pos), tlen=cut(reads["tlen"]), aln_score=cut(reads["aln_score"]), frag=cut(reads["frag"]))
    if dense_frag:
        uniq, inv = torch.unique(out["frag"], return_inverse=True)
        out["frag"] = inv.to(out["frag"].dtype); out["frag_map"] = uniq
    cig_parts, qual_parts, seq_parts, co_parts, so_parts = [], [], [], [], []
    cbase = 0; sbase = 0
    aligned = True
    for a, b in rng:
        c0 = int(coff[a].item()) & 0xFFFFFFFF; c1 = int(coff[b].item()) & 0xFFFFFFFF
        s0 = int(soff[a].item()); s1 = int(soff[b].item())
        cig_parts.append(reads["cigar"][c0:c1]); qual_parts.append(reads["qual"][s0:s1])
        co_parts.append(coff[a:b].to(torch.int64).bitwise_and(0xFFFFFFFF) - c0 + cbase)
        so_parts.append(soff[a:b].to(torch.int64) - s0 + sbase)
        if (s0 & 1) or (sbase & 1):
            aligned = False
        seq_parts.append((s0, s1))
        cbase += c1 - c0; sbase += s1 - s0
    end_c = torch.tensor([cbase], dtype=torch.int64, device=dev); end_s = torch.tensor([sbase], dtype=torch.int64, device=dev)
    out["cigar_off"] = torch.cat(co_parts + [end_c]).to(torch.int32)           # u32 values travel as int32 bit patterns
    out["seq_off"] = torch.cat(so_parts + [end_s]).to(soff.dtype)
    out["cigar"] = torch.cat(cig_parts) if cig_parts else reads["cigar"][:0]
    out["qual"] = torch.cat(qual_parts) if qual_parts else reads["qual"][:0]
    seq = reads["seq"]
    if aligned:      # every piece starts on a byte boundary of both the source and the destination
        out["seq"] = torch.cat([seq[s0 >> 1:(s1 + 1) >> 1] for s0, s1 in seq_parts]) if seq_parts else seq[:0]
    else:            # general case: through one code per base
        codes = []
        for s0, s1 in seq_parts:
            by = seq[s0 >> 1:(s1 + 1) >> 1]
            u = torch.stack([by >> 4, by & 15], 1).reshape(-1)
            codes.append(u[(s0 & 1):(s0 & 1) + (s1 - s0)])
        u = torch.cat(codes) if codes else seq[:0]
        if u.shape[0] & 1:
            u = torch.cat([u, torch.zeros(1, dtype=u.dtype, device=dev)])
        out["seq"] = ((u[0::2] << 4) | u[1::2]).contiguous()
    return out


# ------------------------------------------------------------------------------------------------ merge

def merge_results(parts, vt: VariantTable, n_bams: int) -> PhaseResult:
    """parts: per rank (PhaseResult, global variant ids of the rank's table, global contig ids) or None for a rank
    without contigs.  Returns one PhaseResult in global variant ids with blocks in the global output order
    (phaser.py:863-867).  Array work only: the cost does not depend on Python loops over blocks or variants."""
    V = vt.n_variants
    nb = n_bams
    nc = len(vt.contigs)
    vfirst = np.full(V, np.iinfo(np.uint64).max, np.uint64)
    ncls = np.zeros(V * 3, np.uint32); setsize = np.zeros(V * 3, np.uint32); vb = np.zeros(V * nb * 2, np.uint32)
    v_final = np.full(V, NONE32, np.uint32); v_hap = np.zeros(V, np.uint8)
    contig_of = np.repeat(np.arange(nc, dtype=np.int64), np.diff(np.asarray(vt.contig_var_off, np.int64)))
    live = [(ri, p) for ri, p in enumerate(parts) if p is not None and p[0] is not None]
    counters = {}
    first_bam_of_contig = np.full(nc, 1 << 30, np.int64)
    for ri, (res, gid, _c) in live:
        for k, v in res.counters.items():
            counters[k] = counters.get(k, 0) + v
        bam_start = np.concatenate([[0], np.cumsum(res.tuples_per_bam)]).astype(np.int64)
        lf = res.vfirst.astype(np.int64)
        seen = res.vfirst != NONE32
        bam_of = np.searchsorted(bam_start, lf, side="right") - 1
        # comparable across ranks: (BAM of the first tuple, contig, tuple order inside the rank)
        key = (bam_of.astype(np.uint64) << np.uint64(56)) | (contig_of[gid].astype(np.uint64) << np.uint64(40)) | lf.astype(np.uint64)
        vfirst[gid[seen]] = key[seen]
        for name, dst, w in (("ncls", ncls, 3), ("setsize", setsize, 3), ("vb_cnt", vb, nb * 2)):
            dst.reshape(V, w)[gid] = res.arrays[name].reshape(-1, w)
        v_hap[gid] = res.v_hap
        # contig order of first appearance (phaser.py:573-574): first BAM with a tuple on the contig, then VCF order
        np.minimum.at(first_bam_of_contig, contig_of[gid[seen]], bam_of[seen])
    first_bam_of_contig[first_bam_of_contig == (1 << 30)] = 0
    # ---- global block order: (first BAM of the contig, contig, local order); a contig lives on one rank
    k_bam, k_contig, k_local, k_rank = [], [], [], []
    for ri, (res, gid, _c) in live:
        nf = int(res.fb_first.shape[0])
        c = contig_of[gid[res.members[res.fb_first.astype(np.int64)].astype(np.int64)]] if nf else np.zeros(0, np.int64)
        k_bam.append(first_bam_of_contig[c]); k_contig.append(c); k_local.append(np.arange(nf, dtype=np.int64))
        k_rank.append(np.full(nf, ri, np.int64))
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    k_bam = cat(k_bam, np.int64); k_contig = cat(k_contig, np.int64); k_local = cat(k_local, np.int64); k_rank = cat(k_rank, np.int64)
    order = np.lexsort((k_local, k_contig, k_bam))          # global position -> index into the concatenation
    NF = order.shape[0]
    new_of_cat = np.empty(NF, np.int64); new_of_cat[order] = np.arange(NF)
    cat_len = cat([p[0].fb_len for _, p in live], np.int64)
    g_len = cat_len[order]
    g_first = np.zeros(NF, np.int64)
    if NF:
        g_first[1:] = np.cumsum(g_len)[:-1]
    members = np.zeros(int(g_len.sum()), np.uint32)
    fb_sup = np.zeros(NF, np.uint32); fb_tot = np.zeros(NF, np.uint32)
    fb_cnt = np.zeros((NF, 2), np.uint32); fb_bcnt = np.zeros((NF, nb, 2), np.uint32)
    ed = {k: [] for k in ("ed_a", "ed_b", "ed_sup", "ed_tot", "ed_cfg", "ed_keep")}
    rl = {k: [] for k in ("rl_frag", "rl_var", "rl_row")}
    sg = {k: [] for k in ("sg_var", "sg_cb", "sg_frag", "g_var", "g_cb", "g_frag")}
    bb = max(1, int(np.ceil(np.log2(max(nb, 2)))))
    base = 0
    for ri, (res, gid, _c) in live:
        nf = int(res.fb_first.shape[0])
        new_id = new_of_cat[base:base + nf]          # local final block -> global block index
        base += nf
        ln = res.fb_len.astype(np.int64); lfirst = res.fb_first.astype(np.int64)
        tot = int(ln.sum())
        if tot:
            rep = np.repeat(np.arange(nf), ln)
            within = np.arange(tot) - np.repeat(np.cumsum(ln) - ln, ln)
            members[g_first[new_id[rep]] + within] = gid[res.members[lfirst[rep] + within].astype(np.int64)]
        fb_sup[new_id] = res.fb_sup; fb_tot[new_id] = res.fb_tot
        fb_cnt[new_id] = res.fb_cnt.reshape(-1, 2); fb_bcnt[new_id] = res.fb_bcnt.reshape(-1, nb, 2)
        inb = res.v_final != NONE32
        v_final[gid[inb]] = new_id[res.v_final[inb].astype(np.int64)]
        ed["ed_a"].append(gid[res.ed_a.astype(np.int64)]); ed["ed_b"].append(gid[res.ed_b.astype(np.int64)])
        for k in ("ed_sup", "ed_tot", "ed_cfg", "ed_keep"):
            ed[k].append(res.arrays[k])
        if "sg_var" in res.arrays:
            sg["sg_var"].append(gid[res.sg_var.astype(np.int64)]); sg["sg_cb"].append(res.sg_cb); sg["sg_frag"].append(res.sg_frag)
        if "g_var" in res.arrays:
            sg["g_var"].append(gid[res.g_var.astype(np.int64)]); sg["g_cb"].append(res.g_cb); sg["g_frag"].append(res.g_frag)
        if "rl_row" in res.arrays and res.rl_row.shape[0]:
            row = res.rl_row.astype(np.int64)
            rl["rl_row"].append((new_id[row >> (bb + 1)] << (bb + 1)) | (row & ((1 << (bb + 1)) - 1)))
            rl["rl_var"].append(gid[res.rl_var.astype(np.int64)]); rl["rl_frag"].append(res.rl_frag)
    arrays = dict(vfirst=vfirst, ncls=ncls, setsize=setsize, vb_cnt=vb, v_final=v_final, v_hap=v_hap,
                  members=members, fb_first=g_first.astype(np.uint32), fb_len=g_len.astype(np.uint32),
                  fb_sup=fb_sup, fb_tot=fb_tot, fb_cnt=fb_cnt.reshape(-1), fb_bcnt=fb_bcnt.reshape(-1))
    for k, dt in (("ed_a", np.uint32), ("ed_b", np.uint32), ("ed_sup", np.uint32), ("ed_tot", np.uint32), ("ed_cfg", np.uint8), ("ed_keep", np.uint8)):
        arrays[k] = cat(ed[k], dt)
    if rl["rl_row"]:
        # rows of one block live on one rank, so a stable sort by row keeps variant / tuple order
        row = np.concatenate(rl["rl_row"]); o = np.argsort(row, kind="stable")
        arrays["rl_row"] = row[o].astype(np.uint32); arrays["rl_var"] = np.concatenate(rl["rl_var"])[o].astype(np.uint32)
        arrays["rl_frag"] = np.concatenate(rl["rl_frag"])[o].astype(np.uint32)
    else:
        arrays["rl_row"] = np.zeros(0, np.uint32); arrays["rl_var"] = np.zeros(0, np.uint32); arrays["rl_frag"] = np.zeros(0, np.uint32)
    for pre in ("sg_", "g_"):
        if sg[pre + "var"]:
            for k, dt in (("var", np.uint32), ("cb", np.uint8), ("frag", np.uint32)):
                arrays[pre + k] = cat(sg[pre + k], dt)
    first = live[0][1][0]
    return PhaseResult(nb, first.as_cutoff, [sum(p[0].tuples_per_bam[b] for _, p in live) for b in range(nb)],
                       [sum(p[0].candidates_per_bam[b] for _, p in live) for b in range(nb)],
                       first.noise_e, first.match, first.mismatch, counters, 0, arrays)


def results_equal(a: PhaseResult, b: PhaseResult):
    """Array-level equality of two results of the same inputs (e.g. merged shards vs. one GPU).  The edge table is
    compared as a set of rows (its order is the order of the variant pairs per rank) and `vfirst` through the order
    it defines (a tuple index on one GPU, a (BAM, contig, index) key after a merge).  Returns the names that differ."""
    bad = []
    for k in ("ncls", "setsize", "vb_cnt", "v_final", "v_hap", "fb_len", "fb_sup", "fb_tot",
              "fb_cnt", "fb_bcnt", "rl_row", "rl_var", "rl_frag"):
        if (k in a.arrays) != (k in b.arrays):
            bad.append(k)
        elif k in a.arrays and not np.array_equal(np.asarray(a.arrays[k]).astype(np.int64), np.asarray(b.arrays[k]).astype(np.int64)):
            bad.append(k)

    def block_members(r):        # `members` may hold unused slots between the final blocks (one GPU) or none (merged)
        ln = r.fb_len.astype(np.int64); tot = int(ln.sum())
        if tot == 0:
            return np.zeros(0, np.int64)
        within = np.arange(tot) - np.repeat(np.cumsum(ln) - ln, ln)
        return r.members[np.repeat(r.fb_first.astype(np.int64), ln) + within].astype(np.int64)
    if not np.array_equal(a.fb_len, b.fb_len) or not np.array_equal(block_members(a), block_members(b)):
        bad.append("members")

    def edge_rows(r):
        m = np.stack([r.ed_a.astype(np.int64), r.ed_b.astype(np.int64), r.ed_sup.astype(np.int64), r.ed_tot.astype(np.int64),
                      r.ed_cfg.astype(np.int64), r.ed_keep.astype(np.int64)], 1)
        return m[np.lexsort((m[:, 1], m[:, 0]))] if m.shape[0] else m
    ea, eb = edge_rows(a), edge_rows(b)
    if ea.shape != eb.shape or not np.array_equal(ea, eb):
        bad.append("edges")

    def order(r):
        vf = r.vfirst
        seen = np.nonzero(vf != np.iinfo(vf.dtype).max)[0]
        return seen[np.argsort(vf[seen], kind="stable")]
    if not np.array_equal(order(a), order(b)):
        bad.append("vfirst order")
    for k in ("n_tuples", "edges", "dropped", "members", "final_blocks", "read_list_entries"):
        if a.counters.get(k) != b.counters.get(k):
            bad.append("counter " + k)
    if a.noise_e != b.noise_e or list(a.as_cutoff) != list(b.as_cutoff):
        bad.append("noise / AS cutoff")
    return bad


# ------------------------------------------------------------------------------------------------ merge on the device

def _typed(buf, off, n, eb):
    """typed view of a packed byte buffer; unsigned 32-bit data travels as int32 bit patterns (torch has no uint32 maths)"""
    t = buf[off:off + n * eb]
    return t if eb == 1 else t.view({2: torch.int16, 4: torch.int32, 8: torch.int64}[eb])


def merge_results_device(recv, lay, heads, names, gids, plan, vt: VariantTable, n_bams: int, meta: PhaseResult,
                         contig_of, want_read_ids=False, want_kept_tuples=False, host_cache=None, engine=None) -> PhaseResult:
    """merge_results on the arrays as the gather left them in rank 0's memory (device tensors, or CPU tensors under
    gloo): same result, but every step is a tensor operation where the data already is, and the merged arrays reach
    the host in ONE copy.  recv[r]: packed byte buffer of rank r (None: no contigs); lay[r]: its layout; gids[r]:
    global variant ids of rank r's table (int64 tensor on the same device); contig_of: int64[V] on that device.
    The arrays of all ranks are concatenated per name first and every step then runs ONCE over the concatenation (a
    per-element rank index supplies what differs between ranks: id bases, tuples per BAM), so the number of tensor
    operations -- what the merge costs on a GPU -- does not grow with the number of ranks, and nothing is read back
    before the final copy (all sizes are known from the layouts)."""
    from .engine import COUNTER_NAMES
    dev = contig_of.device
    V = vt.n_variants; nb = n_bams; nc = len(vt.contigs); nn = len(names)
    M32 = 0xFFFFFFFF
    i64 = torch.int64
    live = [r for r in range(len(recv)) if recv[r] is not None]
    nl = len(live)
    trace = [] if os.environ.get("PHZ_MERGE_TRACE") else None

    def mark(what):
        if trace is not None:
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            trace.append((what, time.perf_counter()))
    mark("start")
    # ---- ONE page-locked landing buffer (grow-only, the arrays returned are views of it), filled in two goes: whatever
    # is finished before the read lists are merged leaves on a side stream and travels under that merge
    bound = 64 * 64 + 4096 + V * (8 + 12 + 12 + 8 * nb + 4 + 1) + sum(int(recv[r].numel()) for r in live)
    host = host_cache.get("buf") if host_cache is not None else None
    if host is None or host.numel() < bound:
        host = torch.empty(int(bound * 1.25), dtype=torch.uint8)
        if dev.type == "cuda":
            try:
                host = host.pin_memory()
            except RuntimeError:
                pass
        if host_cache is not None:
            host_cache["buf"] = host
    ship_state = {"total": 0, "place": [], "host": host, "done": set()}

    def ship(arrays, stream=None):
        for k, t in arrays.items():
            if k in ship_state["done"]:
                continue
            t = t.contiguous(); arrays[k] = t
            off = (ship_state["total"] + 63) // 64 * 64; nbytes = t.numel() * t.element_size()
            if off + nbytes > host.numel():
                raise RuntimeError("merge: landing buffer too small for " + k)
            ship_state["place"].append((k, off, nbytes)); ship_state["total"] = off + nbytes; ship_state["done"].add(k)
            if nbytes:
                src = t.view(torch.uint8).reshape(-1) if t.dtype != torch.uint8 else t.reshape(-1)
                if stream is not None:
                    with torch.cuda.stream(stream):
                        host[off:off + nbytes].copy_(src, non_blocking=True)
                else:
                    host[off:off + nbytes].copy_(src, non_blocking=True)

    def count(r, name):
        for nm, _off, n, _eb in lay[r][0]:
            if nm == name:
                return int(n)
        return 0

    def arr(r, name):
        for nm, off, n, eb in lay[r][0]:
            if nm == name:
                return _typed(recv[r], off, n, eb)
        raise KeyError(name)

    def has(name):
        return nl > 0 and all(any(nm == name for nm, _o, _n, _e in lay[r][0]) for r in live)

    def allr(name, dt=torch.int32):
        xs = [arr(r, name) for r in live]
        return torch.cat(xs) if xs else torch.zeros(0, dtype=dt, device=dev)

    def u(t):          # int32 bit pattern -> non-negative int64
        return t.to(i64) & M32

    def per_element(counts, values):
        """values[k] repeated counts[k] times (k = position in `live`), as an int64 tensor"""
        return torch.repeat_interleave(torch.as_tensor(np.asarray(values, np.int64), device=dev),
                                       torch.as_tensor(np.asarray(counts, np.int64), device=dev))

    def bases(counts):
        return np.concatenate([[0], np.cumsum(np.asarray(counts, np.int64))])[:-1]

    # ---- host side: counters and the per-rank sizes (from the small headers and the layouts; no device reads)
    counters = {}; tpb = [0] * nb; cpb = [0] * nb; t_rs = []
    for r in live:
        h = heads[r]
        for i, k in enumerate(COUNTER_NAMES):
            counters[k] = counters.get(k, 0) + int(h[2 * nn + i])
        t_r = [int(x) for x in h[2 * nn + N_COUNTERS:2 * nn + N_COUNTERS + nb]]
        t_rs.append(t_r)
        for b in range(nb):
            tpb[b] += t_r[b]; cpb[b] += int(h[2 * nn + N_COUNTERS + nb + b])
    nv = [int(gids[r].shape[0]) for r in live]; nfb = [count(r, "fb_first") for r in live]; nm = [count(r, "members") for r in live]
    ne = [count(r, "ed_a") for r in live]
    vbase = bases(nv); fbase = bases(nfb); mbase = bases(nm)
    gid = torch.cat([gids[r] for r in live]) if nl else torch.zeros(0, dtype=i64, device=dev)      # local (concatenated) -> global site
    cgid = contig_of[gid]

    # ---- per site
    BIG = torch.iinfo(i64).max
    vfirst = torch.full((V,), BIG, dtype=i64, device=dev)
    ncls = torch.zeros((V, 3), dtype=torch.int32, device=dev); setsize = torch.zeros((V, 3), dtype=torch.int32, device=dev)
    vb = torch.zeros((V, nb * 2), dtype=torch.int32, device=dev)
    v_final = torch.full((V,), -1, dtype=torch.int32, device=dev); v_hap = torch.zeros(V, dtype=torch.uint8, device=dev)
    first_bam = torch.full((nc,), 1 << 30, dtype=i64, device=dev)
    if nl:
        lf = u(allr("vfirst")); seen = lf != M32
        if nb == 1:
            bam_of = torch.zeros_like(lf)
        else:          # first tuple of BAM b on each rank: a tuple rank belongs to the last BAM that starts at or before it
            starts = np.asarray([np.cumsum(t)[:-1] for t in t_rs], np.int64).reshape(nl, nb - 1)
            st = torch.repeat_interleave(torch.as_tensor(starts, device=dev), torch.as_tensor(np.asarray(nv, np.int64), device=dev), dim=0)
            bam_of = (lf[:, None] >= st).sum(1)
        key = (bam_of << 56) | (cgid << 40) | lf
        vfirst[gid[seen]] = key[seen]
        ncls[gid] = allr("ncls").view(-1, 3); setsize[gid] = allr("setsize").view(-1, 3); vb[gid] = allr("vb_cnt").view(-1, nb * 2)
        v_hap[gid] = allr("v_hap", torch.uint8)
        first_bam.scatter_reduce_(0, cgid[seen], bam_of[seen], "amin")
    first_bam[first_bam == (1 << 30)] = 0
    mark("per site")
    # ---- global block order: (first BAM of the contig, contig, local order); ranks in `live` order, as concatenated
    NF = int(sum(nfb))
    if NF:
        mem_g = gid[u(allr("members")) + per_element(nm, vbase)]                 # members as global site ids (concatenated order)
        ff = u(allr("fb_first")) + per_element(nfb, mbase)                       # first member of every block in mem_g
        cat_len = u(allr("fb_len"))
        c = contig_of[mem_g[ff]]
        local = torch.arange(NF, dtype=i64, device=dev) - per_element(nfb, fbase)
        order = torch.argsort((first_bam[c] << 48) | (c << 32) | local)
        new_of_cat = torch.empty(NF, dtype=i64, device=dev); new_of_cat[order] = torch.arange(NF, dtype=i64, device=dev)
        g_len = cat_len[order]
        g_first = torch.cumsum(g_len, 0) - g_len
        # the final blocks need not tile the member arrays (members of dropped blocks): expand the runs from the block table
        # (the one size of the merge that is not in the layouts: repeat_interleave reads it back)
        rep = torch.repeat_interleave(torch.arange(NF, dtype=i64, device=dev), cat_len)
        within = torch.arange(rep.shape[0], dtype=i64, device=dev) - torch.repeat_interleave(torch.cumsum(cat_len, 0) - cat_len, cat_len)
        members = torch.zeros(rep.shape[0], dtype=torch.int32, device=dev)
        members[g_first[new_of_cat[rep]] + within] = mem_g[ff[rep] + within].to(torch.int32)
        fb_sup = torch.zeros(NF, dtype=torch.int32, device=dev); fb_tot = torch.zeros(NF, dtype=torch.int32, device=dev)
        fb_cnt = torch.zeros((NF, 2), dtype=torch.int32, device=dev); fb_bcnt = torch.zeros((NF, nb * 2), dtype=torch.int32, device=dev)
        fb_sup[new_of_cat] = allr("fb_sup"); fb_tot[new_of_cat] = allr("fb_tot")
        fb_cnt[new_of_cat] = allr("fb_cnt").view(-1, 2); fb_bcnt[new_of_cat] = allr("fb_bcnt").view(-1, nb * 2)
        vfl = allr("v_final"); inb = vfl != -1
        v_final[gid[inb]] = new_of_cat[(vfl.to(i64) + per_element(nv, fbase))[inb]].to(torch.int32)
    else:
        new_of_cat = torch.zeros(0, dtype=i64, device=dev)
        g_len = torch.zeros(0, dtype=i64, device=dev); g_first = torch.zeros(0, dtype=i64, device=dev)
        members = torch.zeros(0, dtype=torch.int32, device=dev)
        fb_sup = torch.zeros(0, dtype=torch.int32, device=dev); fb_tot = torch.zeros(0, dtype=torch.int32, device=dev)
        fb_cnt = torch.zeros((0, 2), dtype=torch.int32, device=dev); fb_bcnt = torch.zeros((0, nb * 2), dtype=torch.int32, device=dev)
    mark("blocks")
    # ---- edges: the ranks' tables one after the other, sites renumbered
    ed = {}
    ev = per_element(ne, vbase) if nl else None
    for k in ("ed_a", "ed_b"):
        ed[k] = gid[u(allr(k)) + ev].to(torch.int32) if nl else torch.zeros(0, dtype=torch.int32, device=dev)
    for k, dt in (("ed_sup", torch.int32), ("ed_tot", torch.int32), ("ed_cfg", torch.uint8), ("ed_keep", torch.uint8)):
        ed[k] = allr(k, dt)
    sg = {}
    if (want_read_ids or want_kept_tuples) and nl:
        ng = [count(r, "g_var") for r in live]
        gvl = u(allr("g_var")) + per_element(ng, vbase); gc = allr("g_cb", torch.uint8); gfr = allr("g_frag")
        if want_read_ids:
            sel = torch.nonzero((allr("v_final")[gvl] == -1) & ((gc & 3) < 2))[:, 0]
            sg["sg_var"] = gid[gvl[sel]].to(torch.int32); sg["sg_cb"] = gc[sel]; sg["sg_frag"] = gfr[sel]
        if want_kept_tuples:
            sg["g_var"] = gid[gvl].to(torch.int32); sg["g_cb"] = gc; sg["g_frag"] = gfr
    bb = max(1, int(np.ceil(np.log2(max(nb, 2)))))
    low = (1 << (bb + 1)) - 1
    out = dict(vfirst=vfirst, ncls=ncls.reshape(-1), setsize=setsize.reshape(-1), vb_cnt=vb.reshape(-1), v_final=v_final, v_hap=v_hap,
               members=members, fb_first=g_first.to(torch.int32), fb_len=g_len.to(torch.int32), fb_sup=fb_sup, fb_tot=fb_tot,
               fb_cnt=fb_cnt.reshape(-1), fb_bcnt=fb_bcnt.reshape(-1))
    out.update(ed)
    # ---- read lists: rows of one block live on one rank and every rank's rows are already sorted, so the merged order
    # follows from the run lengths alone (no sort of the entries): destination = start of the run in the merged order +
    # offset inside the run
    if dev.type == "cuda":
        side = host_cache.setdefault("side_stream", torch.cuda.Stream(device=dev)) if host_cache is not None else torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        ship(out, side)
    nrl = [count(r, "rl_row") for r in live] if has("rl_row") else []
    if sum(nrl):
        total = int(sum(nrl))
        row32 = allr("rl_row")
        # runs of equal row ids (a run never continues across two ranks' arrays): head flags on the 32-bit ids
        head = torch.ones(total, dtype=torch.bool, device=dev)
        head[1:] = row32[1:] != row32[:-1]
        rstart = torch.as_tensor(bases(nrl), device=dev)
        head[rstart[rstart < total]] = True
        lstart = torch.nonzero(head)[:, 0]                                               # where the run sits in the concatenation
        nrun = int(lstart.shape[0])
        first = torch.cat([lstart, torch.full((1,), total, dtype=i64, device=dev)])
        cnt = first[1:] - first[:-1]
        lrow = u(row32[lstart])
        rrun = torch.searchsorted(rstart, lstart, right=True) - 1                        # position of the run's rank in `live`
        fb_of_run = torch.as_tensor(fbase, device=dev)[rrun]
        nrow = (new_of_cat[(lrow >> (bb + 1)) + fb_of_run] << (bb + 1)) | (lrow & low)
        o = torch.argsort(nrow)
        start_sorted = torch.cumsum(cnt[o], 0) - cnt[o]
        start = torch.empty_like(start_sorted); start[o] = start_sorted
        rl_row = torch.empty(total, dtype=torch.int32, device=dev); rl_var = torch.empty(total, dtype=torch.int32, device=dev)
        rl_frag = torch.empty(total, dtype=torch.int32, device=dev)
        vb_of_run = torch.as_tensor(vbase, device=dev)[rrun]
        if engine is not None and engine.device.type == dev.type:
            # one kernel of the native library: entry -> run (bisection over the run table) -> place, ids renumbered
            engine.expand_runs(first.contiguous(), start.contiguous(), nrow.contiguous(), vb_of_run.contiguous(),
                               allr("rl_var").contiguous(), allr("rl_frag").contiguous(), gid.contiguous(), rl_row, rl_var, rl_frag)
        else:       # the same as tensor operations (arrays that do not live where the engine's do: gloo on a GPU box)
            dest = torch.repeat_interleave(start - lstart, cnt) + torch.arange(total, dtype=i64, device=dev)
            rl_row[dest] = torch.repeat_interleave(nrow, cnt).to(torch.int32)
            rl_var[dest] = gid[u(allr("rl_var")) + torch.repeat_interleave(vb_of_run, cnt)].to(torch.int32)
            rl_frag[dest] = allr("rl_frag")
        out["rl_row"] = rl_row; out["rl_var"] = rl_var; out["rl_frag"] = rl_frag
    else:
        for k in ("rl_row", "rl_var", "rl_frag"):
            out[k] = torch.zeros(0, dtype=torch.int32, device=dev)
    out.update(sg)
    mark("edges + read lists")
    # ---- the rest of the copy to the host (the arrays finished before the read lists are already on their way)
    ship(out)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    total = ship_state["total"]; place = ship_state["place"]; host = ship_state["host"]
    mark("copy to the host (%d bytes)" % total)
    if trace is not None:
        print("[merge] " + ", ".join("%s %.2f ms" % (w, (t - trace[i][1]) * 1e3) for i, (w, t) in enumerate(trace[1:])), file=sys.stderr)
    hn = host.numpy()
    npdt = {torch.int32: np.uint32, torch.uint8: np.uint8, torch.int64: np.int64}
    arrays = {k: hn[off:off + nbytes].view(npdt[out[k].dtype]) for (k, off, nbytes) in place}
    return PhaseResult(nb, meta.as_cutoff, tpb, cpb, meta.noise_e, meta.match, meta.mismatch, counters, 0, arrays)


# ------------------------------------------------------------------------------------------------ gather

def gather_names(params: PhaseParams):
    names = list(RESULT_ARRAYS)
    if params.want_read_lists:
        names += ["rl_frag", "rl_var", "rl_row"]
    if params.want_kept_tuples or params.want_read_ids:
        names += ["g_var", "g_cb", "g_frag"]
    return names


def gather_results(engine, res, names, n_bams, device, timers=None, to_host=True, failed=False, frag_map=None):
    """Every rank packs its result arrays into ONE buffer where they live (device memory: phz_copy_array), the byte
    counts travel in a small all-gather, and a grouped send/recv moves the buffers into rank 0's memory.  Rank 0
    returns, per rank, {name: numpy array} + (counters, tuples per BAM, candidates per BAM) read from one page-locked
    host copy (`to_host`) -- or the raw device buffers; the other ranks return None.  `res` is None on a rank that
    owns no contig."""
    world = dist.get_world_size(); rank = dist.get_rank()
    device = torch.device(device)

    def sync():
        if device.type == "cuda":
            torch.cuda.synchronize(device)

    if timers is not None:
        sync(); t0 = time.perf_counter()
    nn = len(names)
    head = torch.zeros(nn * 2 + N_COUNTERS + 2 * n_bams + 1, dtype=torch.int64)
    buf = None
    if res is not None:
        buf, info = engine.pack_arrays(names)
        if frag_map is not None:          # shard-local fragment ids -> the sample's ids, before the arrays leave the rank
            for nm, off, n, eb in info:
                if nm in ("rl_frag", "g_frag") and n:
                    view = buf[off:off + n * eb].view(torch.int32)
                    view.copy_(frag_map.to(view.device)[view.to(torch.int64) & 0xFFFFFFFF].to(torch.int32))
        if buf.device != device:          # collectives on another device than the engine's (gloo between CUDA engines)
            buf = buf.to(device)
        for i, (_nm, _off, n, eb) in enumerate(info):
            head[2 * i] = n; head[2 * i + 1] = eb
        cn = list(res.counters.values())
        head[2 * nn:2 * nn + len(cn)] = torch.tensor(cn, dtype=torch.int64)
        head[2 * nn + N_COUNTERS:2 * nn + N_COUNTERS + n_bams] = torch.tensor(res.tuples_per_bam, dtype=torch.int64)
        head[2 * nn + N_COUNTERS + n_bams:2 * nn + N_COUNTERS + 2 * n_bams] = torch.tensor(res.candidates_per_bam, dtype=torch.int64)
        head[-1] = 1
    if failed:
        head[-1] = 2
    heads = [torch.zeros_like(head, device=device) for _ in range(world)]
    dist.all_gather(heads, head.to(device))
    heads = torch.stack(heads).cpu().numpy()
    if (heads[:, -1] == 2).any():
        if failed:
            return None
        from .vcfio import PhaserFatal
        raise PhaserFatal("rank %s failed; see its message" % ",".join(str(r) for r in np.nonzero(heads[:, -1] == 2)[0].tolist()))

    def layout(h):
        out = []; total = 0
        for i in range(nn):
            n = int(h[2 * i]); eb = int(h[2 * i + 1])
            off = (total + 63) // 64 * 64
            out.append((names[i], off, n, eb)); total = off + n * eb
        return out, max(total, 64)

    lay = [layout(h) for h in heads]
    ops = []; recv = [None] * world
    if rank == 0:
        recv[0] = buf
        for r in range(1, world):
            if heads[r][-1] == 1:
                recv[r] = torch.empty(lay[r][1], dtype=torch.uint8, device=device)
                ops.append(dist.P2POp(dist.irecv, recv[r], r))
    elif res is not None:
        ops.append(dist.P2POp(dist.isend, buf, 0))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if timers is not None:
        sync(); timers["gather_results_ms"] = timers.get("gather_results_ms", 0.0) + (time.perf_counter() - t0) * 1e3
        timers["gather_bytes_into_rank0"] = int(sum(lay[r][1] for r in range(1, world) if heads[r][-1] == 1))
    if rank != 0:
        return None
    if not to_host:
        return recv, lay, heads
    # one page-locked host copy of everything, then views
    total = sum(lay[r][1] for r in range(world) if recv[r] is not None)
    host = getattr(engine, "_gather_host", None)
    if host is None or host.numel() < total:
        host = torch.empty(int(total * 1.25) + 4096, dtype=torch.uint8)
        if device.type == "cuda":
            try:
                host = host.pin_memory()
            except RuntimeError:
                pass
        engine._gather_host = host
    o = 0; offs = []
    for r in range(world):
        offs.append(o)
        if recv[r] is not None:
            host[o:o + lay[r][1]].copy_(recv[r][:lay[r][1]], non_blocking=True)
            o += lay[r][1]
    sync()
    hn = host.numpy()
    out = []
    for r in range(world):
        if recv[r] is None:
            out.append(None); continue
        h = heads[r]
        arrays = {nm: hn[offs[r] + off:offs[r] + off + n * eb].view(_DT[eb]) for nm, off, n, eb in lay[r][0]}
        from .engine import COUNTER_NAMES
        counters = {k: int(h[2 * nn + i]) for i, k in enumerate(COUNTER_NAMES)}
        tpb = [int(x) for x in h[2 * nn + N_COUNTERS:2 * nn + N_COUNTERS + n_bams]]
        cpb = [int(x) for x in h[2 * nn + N_COUNTERS + n_bams:2 * nn + N_COUNTERS + 2 * n_bams]]
        out.append((arrays, counters, tpb, cpb))
    return out


def _idle_rank(params: PhaseParams, n_bams: int, comm):
    """A rank that owns no contig still takes part in every collective of run_path, with zeros, and reaches the same
    verdicts (cutoffs, noise level, fatal errors) as the ranks that do."""
    from .engine import AS_BINS
    from .layout import AS_MISSING
    from .pipeline import percentile_from_histogram, noise_level
    from .vcfio import PhaserFatal
    cutoffs = []
    for _ in range(n_bams):
        cutoff = None
        if params.as_q_cutoff > 0:
            hist = comm.allreduce_sum(torch.zeros(AS_BINS, dtype=torch.int64)).cpu().numpy()
            n_missing = int(hist[AS_MISSING + 32768]); hist[AS_MISSING + 32768] = 0
            cutoff = percentile_from_histogram(hist, params.as_q_cutoff)
            if cutoff is not None and n_missing > 0:
                raise PhaserFatal("%d mapped reads carry no AS:i tag but an alignment-score cutoff is active" % n_missing)
        cutoffs.append(cutoff)
    match, mism = comm.allreduce_sum_ints([0, 0])
    if match == 0:
        raise PhaserFatal("No reads could be matched to variants.")
    return PhaseResult(n_bams, cutoffs, [0] * n_bams, [0] * n_bams, noise_level(match, mism), match, mism, {}, 0, {})


class ShardedRun:
    """One sample over the ranks of the process group, sharded by contig.  Built once per sample (plan, this rank's
    variant table and ids); `step(batches)` runs the path on this rank's contigs with the exact reductions, gathers
    the result arrays into rank 0's memory and -- on rank 0 -- merges them into one PhaseResult in global ids."""

    def __init__(self, engine, vt: VariantTable, weights: List[int], params: PhaseParams, n_fragments: int, n_bams: int,
                 device=None, timers=None):
        self.engine = engine; self.vt = vt; self.params = params; self.n_fragments = n_fragments; self.n_bams = n_bams
        self.world = dist.get_world_size(); self.rank = dist.get_rank()
        self.device = torch.device(device if device is not None else engine.device)
        self.plan = plan_shards(weights, self.world)
        self.weights = list(weights)
        self.mine = self.plan[self.rank]
        self.timers = timers
        self.comm = DistComm(self.device, timers)
        self.svt, self.gid = sub_variant_table(vt, self.mine)
        self.gids = [sub_variant_table_ids(vt, p) for p in self.plan] if self.rank == 0 else None
        self.names = gather_names(params)
        self.frag_map = None          # local -> global fragment ids when the shard's reads were renumbered (sub_reads_tensors)
        self.host_cache = {}
        if self.rank == 0:          # the merge runs where the gathered arrays are
            self.gids_dev = [torch.from_numpy(g).to(self.device) for g in self.gids]
            nc = len(vt.contigs)
            self.contig_of_dev = torch.from_numpy(np.repeat(np.arange(nc, dtype=np.int64),
                                                            np.diff(np.asarray(vt.contig_var_off, np.int64)))).to(self.device)

    def loads(self):
        return [int(sum(self.weights[c] for c in p)) for p in self.plan]

    def step(self, batches, host_inputs=False, merge=True, to_host=True):
        from .vcfio import PhaserFatal
        err = None; res = None
        try:
            if self.mine:
                res = run_path(self.engine, self.svt, batches, self.params, self.n_fragments, comm=self.comm,
                               host_inputs=host_inputs, download=False)
            else:
                res = _idle_rank(self.params, self.n_bams, self.comm)
        except PhaserFatal as e:          # rank-local verdicts (phasing flags) must not leave the others in the gather
            err = e
        got = gather_results(self.engine, res if (self.mine and err is None) else None, self.names, self.n_bams, self.device,
                             self.timers, to_host=False, failed=err is not None, frag_map=self.frag_map)
        if err is not None:
            raise err
        if self.rank != 0 or not merge:
            return None
        if self.timers is not None:
            t0 = time.perf_counter()
        recv, lay, heads = got
        out = merge_results_device(recv, lay, heads, self.names, self.gids_dev, self.plan, self.vt, self.n_bams, res,
                                   self.contig_of_dev, want_read_ids=self.params.want_read_ids,
                                   want_kept_tuples=self.params.want_kept_tuples, host_cache=self.host_cache, engine=self.engine)
        if self.timers is not None:
            self.timers["merge_on_rank0_ms"] = self.timers.get("merge_on_rank0_ms", 0.0) + (time.perf_counter() - t0) * 1e3
        return out


def sub_variant_table_ids(vt: VariantTable, contigs: List[int]):
    return (np.concatenate([np.arange(vt.contig_var_off[c], vt.contig_var_off[c + 1]) for c in contigs]) if contigs
            else np.zeros(0, np.int64)).astype(np.int64)


def contig_weights(vt: VariantTable, batches) -> List[int]:
    """records per contig over all BAMs (+1 so that empty contigs still spread)"""
    def off(b):
        return np.asarray(b.contig_rec_off if isinstance(b, ReadBatch) else b["contig_rec_off"], np.int64)
    return [int(sum(off(b)[c + 1] - off(b)[c] for b in batches)) + 1 for c in range(len(vt.contigs))]


def run_sharded(engine, vt: VariantTable, batches: List[ReadBatch], params: PhaseParams, n_fragments: int, device=None):
    """Every rank calls this with the SAME host inputs (or at least its own contigs' part of them); rank 0
    gets the merged PhaseResult, the others None."""
    run = ShardedRun(engine, vt, contig_weights(vt, batches), params, n_fragments, len(batches), device=device)
    dev = [engine.upload_reads(sub_read_batch(b, run.mine)) for b in batches] if run.mine else []
    return run.step(dev)


class ThreadComm:
    """The same reductions between N logical shards that run as threads of ONE process (one engine each,
    possibly on one GPU): used to check that the merged output does not depend on the sharding."""

    def __init__(self, world):
        import threading
        self.world_size = world
        self._barrier = threading.Barrier(world)
        self._slots = [None] * world
        self._lock = threading.Lock()

    def view(self, rank):
        parent = self

        class _View:
            world_size = parent.world_size

            def __init__(self):
                self.rank = rank

            def _exchange(self, value):
                parent._slots[rank] = value
                parent._barrier.wait()
                vals = list(parent._slots)
                parent._barrier.wait()
                return vals

            def allreduce_sum(self, t):
                vals = self._exchange(t.detach().cpu().clone())
                out = vals[0].clone()
                for v in vals[1:]:
                    out += v
                return out.to(t.device)

            def allreduce_sum_ints(self, xs):
                vals = self._exchange(list(xs))
                return [sum(v[i] for v in vals) for i in range(len(xs))]

            def allreduce_max_int(self, x):
                return max(self._exchange(int(x)))

        return _View()


def run_logical_shards(make_engine, vt: VariantTable, batches: List[ReadBatch], params: PhaseParams, n_fragments: int,
                       n_shards: int) -> PhaseResult:
    """N logical shards (threads, one engine each) + merge: must equal the unsharded run."""
    import threading
    plan = plan_shards(contig_weights(vt, batches), n_shards)
    comm = ThreadComm(n_shards)
    parts = [None] * n_shards
    errors = []

    def work(r):
        try:
            if not plan[r]:
                _idle_rank(params, len(batches), comm.view(r))
                return
            engine = make_engine()
            svt, gid = sub_variant_table(vt, plan[r])
            dev = [engine.upload_reads(sub_read_batch(b, plan[r])) for b in batches]
            parts[r] = (run_path(engine, svt, dev, params, n_fragments, comm=comm.view(r)), gid, plan[r])
        except BaseException as e:          # noqa: BLE001 -- re-raised in the caller; never leave the others at the barrier
            errors.append(e)
            comm._barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(n_shards)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return merge_results(parts, vt, len(batches))
