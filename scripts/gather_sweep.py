#!/usr/bin/env python
"""NVLink gather sweep of configs[4]: the result gather of the contig-sharded path moves a few MB per
rank to rank 0 (shard.py: gather of the result arrays); this measures that collective from 1 KiB to 256 MiB per rank.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/gather_sweep.py
One JSON line per message size (rank 0): device-timed with CUDA events, max over ranks."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    size = 1 << 10
    while size <= (256 << 20):
        src = torch.full((size,), rank, dtype=torch.uint8, device=dev)
        dst = [torch.empty(size, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
        for _ in range(3):
            dist.gather(src, dst, dst=0)
        torch.cuda.synchronize(); dist.barrier()
        reps = 20 if size <= (16 << 20) else 5
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dist.gather(src, dst, dst=0)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ok = all(int(dst[r][0]) == r and int(dst[r][-1]) == r for r in range(world))
            ms = float(t.item())
            print(json.dumps({"collective": "gather to rank 0", "n_gpus": world, "bytes_per_rank": size, "ms": ms,
                              "gbs_into_rank0": (world - 1) * size / (ms * 1e-3) / 1e9, "correct": ok}))
        size <<= 2
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
