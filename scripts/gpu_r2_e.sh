#!/bin/bash
# Round 2, visit E (1 GPU): graph A/B tests, bench N=1 with the WGS-shape leg.
mkdir -p gpurun_out
echo "== pytest (gpu parity file)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2e_pytest_gpu.log
echo "== bench N=1"
timeout 1200 python bench.py --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; tail -3 gpurun_out/r2e_bench_n1.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2e_bench_n1.json").read().strip().splitlines()[-1])
print("value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k: v for k, v in d["stages_ms"].items()})
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["single_sample_ms"], d["e2e"].get("host_placement_rank0"))
print("wgs", json.dumps(d.get("wgs_shape"))[:3000])
PY
