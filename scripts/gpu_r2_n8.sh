#!/bin/bash
# Round 2, 8-GPU visit: the bench exactly as the driver launches it (one sample sharded by contig over 8 GPUs + the replica
# second key), the NCCL sharded parity test, the result-gather sweep at N=8.
mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2; nproc
nvidia-smi topo -m > gpurun_out/r2_n8_topo.txt 2>&1
echo "== sharded parity over NCCL (2 ranks)"; timeout 600 python -m pytest tests/test_shard_gloo.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_n8_pytest.log
for n in 8 4; do
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2964$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; tail -3 gpurun_out/r2_bench_n$n.err | cut -c1-300
python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r2_bench_n%s.json" % n).read().strip().splitlines()[-1])
    print("N", n, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["single_sample_ms"])
    print("  sharding", json.dumps(d.get("sharding", {}).get("collectives_ms_rank0_one_step_synchronised")), d.get("sharding", {}).get("resident_ms_per_rank"), d.get("full_size_checks", {}).get("merged_equals_single_gpu"))
    print("  replicas", json.dumps(d.get("replicas"))[:600])
except Exception as e:
    print("N", n, "ERR", e)
PY
done
echo "== gather sweep N=8"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29649 scripts/gather_sweep.py > gpurun_out/r2_gather_sweep_n8.jsonl 2> gpurun_out/r2_gather_sweep_n8.err; tail -4 gpurun_out/r2_gather_sweep_n8.jsonl
