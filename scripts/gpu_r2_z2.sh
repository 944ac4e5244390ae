#!/bin/bash
# Round 2, visit Z2 (2 GPUs): what the helper thread of run_path waits for at N = 2 (PHZ_TRACE).
mkdir -p gpurun_out
PHZ_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29666 bench.py --gpus 2 --steps 5 --warmup 3 --no_replicas --no_e2e > gpurun_out/r2z2_bench_n2.json 2> gpurun_out/r2z2_bench_n2.err
grep "run_path" gpurun_out/r2z2_bench_n2.err | tail -12
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2z2_bench_n2.json").read().strip().splitlines()[-1])
    print("N=2 value %.4g ms %.3f" % (d["value"], d["ms_per_step"]))
    print("  stages", d["stages_ms"])
    print("  sharding", json.dumps(d.get("sharding", {}).get("collectives_ms_rank0_one_step_synchronised")), d.get("full_size_checks"))
except Exception as e:
    print("ERR", e)
PY
echo "== sharded parity over NCCL (2 ranks)"; timeout 900 python -m pytest tests/test_shard_gloo.py -m gpu -x -q 2>&1 | tail -3
