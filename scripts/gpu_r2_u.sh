#!/bin/bash
# Round 2, visit U (2 GPUs): merge on rank 0 with the native run expansion and the early shipment -- NCCL sharded parity
# tests, section times of the merge, sharded bench at N=2.
mkdir -p gpurun_out
echo "== sharded parity over NCCL (2 ranks)"; timeout 900 python -m pytest tests/test_shard_gloo.py tests/test_gpu_parity.py -m gpu -x -q -k "shard or rank" 2>&1 | tail -4 | tee gpurun_out/r2u_pytest.log
PHZ_MERGE_TRACE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29664 bench.py --gpus 2 --steps 10 --warmup 3 --no_replicas > gpurun_out/r2u_bench_n2.json 2> gpurun_out/r2u_bench_n2.err
grep "\[merge\]" gpurun_out/r2u_bench_n2.err | tail -3
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2u_bench_n2.json").read().strip().splitlines()[-1])
    print("N=2 value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["ms_per_step"], d["e2e"]["single_sample_ms"])
    print("  sharding", json.dumps(d.get("sharding", {}).get("collectives_ms_rank0_one_step_synchronised")), d.get("full_size_checks"))
except Exception as e:
    print("ERR", e)
PY
