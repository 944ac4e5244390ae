"""N > 1 path on CPU: two gloo ranks, contigs sharded by LPT, exact reductions, merge on rank 0."""
import os
import socket
import subprocess
import sys

import pytest

from phaser_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_plan_is_balanced_and_complete():
    w = [248, 242, 198, 190, 181, 170, 159, 145, 138, 133, 135, 133, 114, 107, 101, 90, 83, 80, 58, 64, 46, 50, 156, 57]
    plan = shard.plan_shards(w, 8)
    assert sorted(c for p in plan for c in p) == list(range(24))
    loads = [sum(w[c] for c in p) for p in plan]
    assert max(loads) <= 1.15 * sum(w) / 8


def _two_ranks(case, **env_extra):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), **env_extra)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "shard_worker.py"), case], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "SHARDED PARITY OK" in outs[0]
    return outs[0]


@pytest.mark.parametrize("case", ["rna_two_bams", "rna_conflict", "quirks", "opt_read_ids", "opt_network"])
def test_two_rank_sharded_run_matches_reference(case):
    assert "backend=hostsim comm=gloo" in _two_ranks(case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["rna_two_bams", "rna_conflict", "opt_read_ids"])
def test_two_rank_sharded_run_on_the_cuda_backend(case):
    """The contig-sharded path on the CUDA library, two ranks: over NCCL when the box shows two GPUs; on a one-GPU box
    both ranks drive cuda:0 and the exact reductions / the result gather go through gloo (NCCL refuses two ranks on one
    device), which still runs every sharded code path of the CUDA backend (sub-tables, phz_copy_array packing, merge)."""
    import torch
    comm = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    out = _two_ranks(case, PHZ_ENGINE="gpu", PHZ_COMM=comm)
    assert "backend=cuda-sm_100a comm=%s" % comm in out
    print("sharded CUDA run used", comm)


def _result_files(res, vt, sams, P, vcf, col):
    import gzip
    from phaser_b200 import writer
    from tests import util
    o = writer.Outputs(res, vt, util.bam_display_names(sams), P)
    got = dict(allelic_counts=o.allelic_counts(), variant_connections=o.variant_connections())
    got["haplotypes"], got["haplotypic_counts"], got["allele_config"] = o.block_tables()
    with gzip.open(vcf, "rt") as f:
        got["vcf"], _, _ = o.vcf_text(f.readlines(), col)
    return got


def test_output_does_not_depend_on_the_sharding(tmp_path):
    """1 shard vs 3 logical shards (threads, one host-simulation engine each) over 4 contigs and 2 BAMs:
    identical files after the merge."""
    from oracle import compare
    from phaser_b200 import pipeline
    from tests import util
    contigs = [("19", 120000), ("20", 90000), ("21", 70000), ("22", 50000)]
    vcf, sams = util.make_case(tmp_path, 81, 300, 2500, n_bams=2, contigs=contigs, switch_per_base=0.01)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams()
    e = util.hostsim_engine()
    one = pipeline.run_path(e, vt, [e.upload_reads(b) for b in batches], P, n_fragments=len(fd.names))
    many = shard.run_logical_shards(util.hostsim_engine, vt, batches, P, len(fd.names), 3)
    a = _result_files(one, vt, sams, P, vcf, col); b = _result_files(many, vt, sams, P, vcf, col)
    assert not compare.diff_outputs(a, b)
    assert a["haplotypic_counts"] == b["haplotypic_counts"] and a["vcf"] == b["vcf"]


def test_merged_arrays_equal_the_single_run_and_idle_shards_are_harmless(tmp_path):
    """results_equal (what bench.py asserts at full size): 1 shard == 3 == 6 logical shards over 4 contigs (two of the
    six own no contig and only take part in the reductions)."""
    from phaser_b200 import pipeline
    from tests import util
    contigs = [("19", 120000), ("20", 90000), ("21", 70000), ("22", 50000)]
    vcf, sams = util.make_case(tmp_path, 83, 300, 2500, n_bams=2, contigs=contigs, switch_per_base=0.01)
    vt, st, batches, col, fd = util.load_inputs(vcf, sams)
    P = pipeline.PhaseParams()
    e = util.hostsim_engine()
    one = pipeline.run_path(e, vt, [e.upload_reads(b) for b in batches], P, n_fragments=len(fd.names))
    for n in (3, 6):
        many = shard.run_logical_shards(util.hostsim_engine, vt, batches, P, len(fd.names), n)
        assert shard.results_equal(one, many) == []
    # and the comparison is not vacuous
    many.arrays["fb_cnt"] = many.arrays["fb_cnt"].copy(); many.arrays["fb_cnt"][0] += 1
    assert shard.results_equal(one, many) == ["fb_cnt"]


def test_tensor_form_of_a_shard_equals_the_array_form(tmp_path):
    """sub_reads_tensors (device-side slicing, used for resident shards) == sub_read_batch, including reads of odd
    length whose packed bases start in the middle of a byte."""
    import numpy as np
    import torch
    from tests import util
    contigs = [("19", 120000), ("20", 90000), ("21", 70000), ("22", 50000)]
    for read_len in (76, 75):
        vcf, sams = util.make_case(tmp_path, 90 + read_len, 200, 900, contigs=contigs, read_len=read_len)
        vt, st, batches, col, fd = util.load_inputs(vcf, sams)
        rb = batches[0]
        e = util.hostsim_engine()
        t = e.upload_reads(rb)
        for pick in ([1, 3], [0], [2, 3], [0, 1, 2, 3], []):
            a = shard.sub_read_batch(rb, pick)
            b = shard.sub_reads_tensors(t, pick)
            assert np.array_equal(np.asarray(b["contig_rec_off"]), a.contig_rec_off)
            nb = int(a.qual.shape[0])
            for k in ("pos", "tlen", "aln_score", "frag", "cigar_off", "cigar", "seq_off", "qual"):
                assert np.array_equal(b[k].numpy().view(getattr(a, k).dtype), getattr(a, k)), (read_len, pick, k)
            d = shard.sub_reads_tensors(t, pick, dense_frag=True)          # shard-local fragment ids + the table back
            fm = d["frag_map"].numpy().astype(np.int64) & 0xFFFFFFFF
            assert np.array_equal(fm[d["frag"].numpy().astype(np.int64)], a.frag.astype(np.int64))
            assert fm.shape[0] == np.unique(a.frag).shape[0] and (np.diff(fm) > 0).all()
            sa, sb = a.seq, b["seq"].numpy()
            assert np.array_equal(sa[:nb // 2], sb[:nb // 2]) and (nb % 2 == 0 or (sa[nb // 2] >> 4) == (sb[nb // 2] >> 4))
