#!/bin/bash
# Round 2, visit Z (2 GPUs): the noise all-reduce of the helper thread on a high-priority stream -- host gaps at N = 2.
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29665 bench.py --gpus 2 --steps 10 --warmup 3 --no_replicas > gpurun_out/r2z_bench_n2.json 2> gpurun_out/r2z_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2z_bench_n2.json").read().strip().splitlines()[-1])
    print("N=2 value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["ms_per_step"], d["e2e"]["single_sample_ms"])
    print("  stages", d["stages_ms"])
    print("  sharding", json.dumps(d.get("sharding", {}).get("collectives_ms_rank0_one_step_synchronised")), d.get("full_size_checks"))
except Exception as e:
    print("ERR", e)
PY
