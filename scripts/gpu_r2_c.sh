#!/bin/bash
# Round 2, visit C (2 GPUs): parity suite, sharded bench N=2 (device merge), N=1 bench (fragment-kernel timing).
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -8 | tee gpurun_out/r2c_pytest_gpu.log
echo "== bench N=1"
timeout 900 python bench.py --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -3 gpurun_out/r2c_bench_n1.err | cut -c1-300
echo "== bench N=2 sharded"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; tail -3 gpurun_out/r2c_bench_n2.err | cut -c1-300
python - <<'PY'
import json
for f in ("r2c_bench_n1", "r2c_bench_n2"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["ms_per_step"], d["e2e"]["single_sample_ms"])
        print("  stages", {k: v for k, v in d["stages_ms"].items() if k.startswith("graph")})
        print("  sharding", json.dumps(d.get("sharding", {}).get("collectives_ms_rank0_one_step_synchronised")), d.get("full_size_checks"))
    except Exception as e:
        print(f, "ERR", e)
PY
