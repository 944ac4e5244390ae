"""One rank of the world_size-2 gloo test (tests/test_shard_gloo.py): contig-sharded run on the
host-simulation backend, rank 0 merges, writes the files and diffs them against the reference fixture."""
import gzip
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist      # noqa: E402

from oracle import compare            # noqa: E402
from phaser_b200 import pipeline, shard, writer   # noqa: E402
from tests import util, golden_util as G           # noqa: E402


def main():
    case = sys.argv[1]
    on_gpu = os.environ.get("PHZ_ENGINE", "hostsim") == "gpu"
    # PHZ_COMM=gloo with PHZ_ENGINE=gpu: all ranks drive CUDA engines on ONE GPU and exchange through host memory
    # (NCCL refuses two ranks on one device) -- the sharded path on the CUDA backend where only one GPU is visible
    backend = os.environ.get("PHZ_COMM", "nccl" if on_gpu else "gloo")
    gpu_index = 0
    if on_gpu:
        import torch
        gpu_index = int(os.environ.get("LOCAL_RANK", os.environ["RANK"])) if backend == "nccl" else 0
        torch.cuda.set_device(gpu_index)
    dist.init_process_group(backend, init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"],
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    c = G.load_case(case)
    kw = G.args_to_kw(c["meta"]["args"])
    vt, st, batches, col, fd = util.load_inputs(c["vcf"], c["sams"])
    P = pipeline.PhaseParams(as_q_cutoff=kw.get("as_q_cutoff", 0.05), max_block_size=kw.get("max_block_size", 15),
                             haplo_count_bam_exclude=kw.get("exclude", []), isize=kw.get("isize", [0.0]),
                             want_read_ids=kw.get("output_read_ids", 0) == 1, want_kept_tuples=kw.get("output_network", "") != "")
    if on_gpu:
        from phaser_b200.engine import Engine
        e = Engine(device="cuda:%d" % gpu_index)
    else:
        e = util.hostsim_engine()
    res = shard.run_sharded(e, vt, batches, P, n_fragments=len(fd.names), device="cpu" if backend == "gloo" else None)
    rc = 0
    if dist.get_rank() == 0:
        o = writer.Outputs(res, vt, util.bam_display_names(c["sams"]), P,
                           read_names=fd.names if kw.get("output_read_ids", 0) == 1 else None,
                           output_network=kw.get("output_network", ""))
        got = dict(allelic_counts=o.allelic_counts(), variant_connections=o.variant_connections())
        got["haplotypes"], got["haplotypic_counts"], got["allele_config"] = o.block_tables()
        with gzip.open(c["vcf"], "rt") as f:
            got["vcf"], _, _ = o.vcf_text(f.readlines(), col)
        if o.network is not None:
            got["network_links"], got["network_nodes"] = o.network
        bad = compare.diff_outputs(c["ref"], got)
        if bad:
            print("\n".join(bad)); rc = 1
        else:
            print("SHARDED PARITY OK", case, "backend=%s comm=%s" % (e.backend, backend), res.counters)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(rc)


if __name__ == "__main__":
    main()
