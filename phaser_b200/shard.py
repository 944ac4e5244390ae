"""Multi-GPU: contigs are sharded over ranks, one process per GPU.

The reference itself maps, connects and builds blocks per contig (phaser/phaser.py:442, 533, 556, 650,
784; pairs are same-contig only, :1278-1280), so a contig never needs data of another one.  What IS
global, and therefore exchanged exactly between ranks, is tiny:
  * per BAM, the alignment-score histogram   -> all-reduce(sum)  -> same AS cutoff everywhere (phaser.py:545-553)
  * the two noise counters                   -> all-reduce(sum)  -> same noise_e / critical values (phaser.py:610-631)
  * the result arrays                        -> gather to rank 0, which renumbers blocks in the global
                                                output order (phaser.py:863-867) and writes the files.
No data-path collective: reads and tuples never leave their GPU.  Backend: NCCL on GPUs, gloo in the
CPU tests.
"""
from typing import List

import numpy as np
import torch
import torch.distributed as dist

from .layout import ReadBatch, VariantTable
from .pipeline import PhaseParams, PhaseResult, run_path

NONE32 = 0xFFFFFFFF


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
    """The exact cross-rank reductions of pipeline.run_path over torch.distributed."""

    def __init__(self, device):
        self.rank = dist.get_rank(); self.world_size = dist.get_world_size(); self.device = device

    def allreduce_sum(self, t):
        t = t.to(self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def allreduce_sum_ints(self, xs):
        t = torch.tensor(list(xs), dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
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


def merge_results(parts, vt: VariantTable, n_bams: int) -> PhaseResult:
    """parts: per rank (PhaseResult, global variant ids of the rank's table, global contig ids).  Returns one
    PhaseResult in global variant ids with blocks in the global output order."""
    V = vt.n_variants
    vfirst = np.full(V, np.iinfo(np.uint64).max, np.uint64)
    ncls = np.zeros(V * 3, np.uint32); setsize = np.zeros(V * 3, np.uint32); vb = np.zeros(V * n_bams * 2, np.uint32)
    v_final = np.full(V, NONE32, np.uint32); v_hap = np.zeros(V, np.uint8)
    contig_of = np.searchsorted(vt.contig_var_off, np.arange(V), side="right") - 1
    ed = {k: [] for k in ("ed_a", "ed_b", "ed_sup", "ed_tot", "ed_cfg", "ed_keep")}
    blocks = []          # (order key, rank index, local final block)
    rl = {k: [] for k in ("rl_frag", "rl_var", "rl_row")}
    sg = {k: [] for k in ("sg_var", "sg_cb", "sg_frag", "g_var", "g_cb", "g_frag")}
    counters = {}
    bb = max(1, int(np.ceil(np.log2(max(n_bams, 2)))))
    for ri, (res, gid, gcontigs) in enumerate(parts):
        if res is None:
            continue
        for k, v in res.counters.items():
            counters[k] = counters.get(k, 0) + v
        bam_start = np.concatenate([[0], np.cumsum(res.tuples_per_bam)]).astype(np.int64)
        lf = res.vfirst.astype(np.int64)
        seen = res.vfirst != NONE32
        bam_of = np.searchsorted(bam_start, lf, side="right") - 1
        # comparable across ranks: (BAM of the first tuple, contig, tuple order inside the rank)
        key = (bam_of.astype(np.uint64) << np.uint64(56)) | (contig_of[gid].astype(np.uint64) << np.uint64(40)) | lf.astype(np.uint64)
        vfirst[gid[seen]] = key[seen]
        for name, dst, w in (("ncls", ncls, 3), ("setsize", setsize, 3), ("vb_cnt", vb, n_bams * 2)):
            dst.reshape(V, w)[gid] = res.arrays[name].reshape(-1, w)
        v_hap[gid] = res.v_hap
        ed["ed_a"].append(gid[res.ed_a.astype(np.int64)].astype(np.uint32)); ed["ed_b"].append(gid[res.ed_b.astype(np.int64)].astype(np.uint32))
        for k in ("ed_sup", "ed_tot", "ed_cfg", "ed_keep"):
            ed[k].append(res.arrays[k])
        if "sg_var" in res.arrays:
            sg["sg_var"].append(gid[res.sg_var.astype(np.int64)].astype(np.uint32))
            sg["sg_cb"].append(res.sg_cb); sg["sg_frag"].append(res.sg_frag)
        if "g_var" in res.arrays:
            sg["g_var"].append(gid[res.g_var.astype(np.int64)].astype(np.uint32))
            sg["g_cb"].append(res.g_cb); sg["g_frag"].append(res.g_frag)
        # contig order of first appearance (phaser.py:573-574): first BAM with a tuple on the contig, then VCF order
        first_bam_of_contig = {}
        for v in np.nonzero(seen)[0].tolist():
            c = int(contig_of[gid[v]])
            b = int(bam_of[v])
            if c not in first_bam_of_contig or b < first_bam_of_contig[c]:
                first_bam_of_contig[c] = b
        for f in range(res.fb_first.shape[0]):
            first_member = int(gid[res.members[res.fb_first[f]]])
            c = int(contig_of[first_member])
            blocks.append(((first_bam_of_contig.get(c, 0), c, f), ri, f))
    blocks.sort()
    members, fb_first, fb_len, fb_sup, fb_tot, fb_cnt, fb_bcnt = [], [], [], [], [], [], []
    new_id = {}
    pos = 0
    for g, (_, ri, f) in enumerate(blocks):
        res, gid, _c = parts[ri]
        o = int(res.fb_first[f]); n = int(res.fb_len[f])
        m = gid[res.members[o:o + n].astype(np.int64)]
        members.append(m.astype(np.uint32)); fb_first.append(pos); fb_len.append(n); pos += n
        fb_sup.append(res.fb_sup[f]); fb_tot.append(res.fb_tot[f])
        fb_cnt.append(res.fb_cnt.reshape(-1, 2)[f]); fb_bcnt.append(res.fb_bcnt.reshape(-1, n_bams, 2)[f])
        v_final[m] = g
        new_id[(ri, f)] = g
    for ri, (res, gid, _c) in enumerate(parts):
        if res is None or "rl_row" not in res.arrays or res.rl_row.shape[0] == 0:
            continue
        row = res.rl_row.astype(np.int64)
        f_local = row >> (bb + 1)
        remap = np.array([new_id[(ri, f)] for f in range(res.fb_first.shape[0])], np.int64)
        rl["rl_row"].append(((remap[f_local] << (bb + 1)) | (row & ((1 << (bb + 1)) - 1))).astype(np.uint32))
        rl["rl_var"].append(gid[res.rl_var.astype(np.int64)].astype(np.uint32)); rl["rl_frag"].append(res.rl_frag)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    arrays = dict(vfirst=vfirst, ncls=ncls, setsize=setsize, vb_cnt=vb, v_final=v_final, v_hap=v_hap,
                  members=cat(members, np.uint32), fb_first=np.asarray(fb_first, np.uint32), fb_len=np.asarray(fb_len, np.uint32),
                  fb_sup=np.asarray(fb_sup, np.uint32), fb_tot=np.asarray(fb_tot, np.uint32),
                  fb_cnt=cat(fb_cnt, np.uint32).reshape(-1), fb_bcnt=cat([x.reshape(-1) for x in fb_bcnt], np.uint32))
    for k, dt in (("ed_a", np.uint32), ("ed_b", np.uint32), ("ed_sup", np.uint32), ("ed_tot", np.uint32), ("ed_cfg", np.uint8), ("ed_keep", np.uint8)):
        arrays[k] = cat(ed[k], dt)
    if rl["rl_row"]:
        # rows of one block live on one rank, so a stable sort by row keeps variant / tuple order
        row = np.concatenate(rl["rl_row"]); order = np.argsort(row, kind="stable")
        arrays["rl_row"] = row[order]; arrays["rl_var"] = np.concatenate(rl["rl_var"])[order]; arrays["rl_frag"] = np.concatenate(rl["rl_frag"])[order]
    else:
        arrays["rl_row"] = np.zeros(0, np.uint32); arrays["rl_var"] = np.zeros(0, np.uint32); arrays["rl_frag"] = np.zeros(0, np.uint32)
    for pre in ("sg_", "g_"):
        if sg[pre + "var"]:
            for k, dt in (("var", np.uint32), ("cb", np.uint8), ("frag", np.uint32)):
                arrays[pre + k] = cat(sg[pre + k], dt)
    first = next(p[0] for p in parts if p[0] is not None)
    return PhaseResult(n_bams, first.as_cutoff, [sum(p[0].tuples_per_bam[b] for p in parts if p[0] is not None) for b in range(n_bams)],
                       [sum(p[0].candidates_per_bam[b] for p in parts if p[0] is not None) for b in range(n_bams)],
                       first.noise_e, first.match, first.mismatch, counters, 0, arrays)


def run_sharded(engine, vt: VariantTable, batches: List[ReadBatch], params: PhaseParams, n_fragments: int, device=None):
    """Every rank calls this with the SAME host inputs (or at least its own contigs' part of them); rank 0
    gets the merged PhaseResult, the others None."""
    world = dist.get_world_size(); rank = dist.get_rank()
    weights = [int(sum(b.contig_rec_off[c + 1] - b.contig_rec_off[c] for b in batches)) + 1 for c in range(len(vt.contigs))]
    mine = plan_shards(weights, world)[rank]
    comm = DistComm(device if device is not None else engine.device)
    svt, gid = sub_variant_table(vt, mine)
    dev = [engine.upload_reads(sub_read_batch(b, mine)) for b in batches]
    res = run_path(engine, svt, dev, params, n_fragments, comm=comm)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((res, gid, mine), gathered, dst=0)
    if rank != 0:
        return None
    return merge_results(gathered, vt, len(batches))


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
    weights = [int(sum(b.contig_rec_off[c + 1] - b.contig_rec_off[c] for b in batches)) + 1 for c in range(len(vt.contigs))]
    plan = plan_shards(weights, n_shards)
    comm = ThreadComm(n_shards)
    parts = [None] * n_shards
    errors = []

    def work(r):
        try:
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
