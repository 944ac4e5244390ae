#!/bin/bash
# Round 2, visit X (8 GPUs): the bench exactly as the driver launches it at N = 8 and N = 4 (one sample sharded by contig +
# the replica second key).
mkdir -p gpurun_out
for n in 8 4; do
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2974$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2x_bench_n$n.json 2> gpurun_out/r2x_bench_n$n.err; tail -2 gpurun_out/r2x_bench_n$n.err | cut -c1-300
python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r2x_bench_n%s.json" % n).read().strip().splitlines()[-1])
    print("N", n, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["single_sample_ms"])
    print("  stages", d["stages_ms"])
    print("  sharding", json.dumps(d.get("sharding", {}).get("collectives_ms_rank0_one_step_synchronised")), d.get("sharding", {}).get("resident_ms_per_rank"), d.get("full_size_checks", {}).get("merged_equals_single_gpu"))
    print("  replicas", json.dumps(d.get("replicas"))[:600])
except Exception as e:
    print("N", n, "ERR", e)
PY
done
