#!/bin/bash
# Round 2, visit X2 (8 GPUs): one sample sharded over 8 GPUs with the device-side noise all-reduce (no replica leg).
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29758 bench.py --gpus 8 --steps 10 --warmup 3 --no_replicas > gpurun_out/r2x2_bench_n8.json 2> gpurun_out/r2x2_bench_n8.err; tail -2 gpurun_out/r2x2_bench_n8.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2x2_bench_n8.json").read().strip().splitlines()[-1])
    print("N 8 value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["single_sample_ms"])
    print("  stages", d["stages_ms"])
    print("  sharding", json.dumps(d.get("sharding", {}).get("collectives_ms_rank0_one_step_synchronised")), d.get("sharding", {}).get("resident_ms_per_rank"), d.get("full_size_checks", {}).get("merged_equals_single_gpu"))
except Exception as e:
    print("ERR", e)
PY
