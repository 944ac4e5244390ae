#!/bin/bash
# Round 2, visit D (1 GPU): graph A/B tests, bench N=1, ncu of the fragment kernel.
mkdir -p gpurun_out
echo "== pytest (graph / sharding / option tests)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2d_pytest_gpu.log
echo "== bench N=1"
timeout 900 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_e2e > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; tail -3 gpurun_out/r2d_bench_n1.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2d_bench_n1.json").read().strip().splitlines()[-1])
print("value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k: v for k, v in d["stages_ms"].items()})
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fragment_kernel' -c 1 -o gpurun_out/r2_frag_full -f python bench.py --steps 1 --warmup 0 --no_e2e --no_cpu_baseline > gpurun_out/r2b_ncu_frag.log 2>&1; tail -1 gpurun_out/r2b_ncu_frag.log | cut -c1-100
