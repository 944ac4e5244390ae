#!/bin/bash
# Round 2, visit O (1 GPU): tile-uniform geometry from the pre-pass, shuffle scan, read lists without a flag array.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -8 | tee gpurun_out/r2o_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_e2e --no_wgs > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
python - <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2o_bench.json").read().strip().splitlines()[-1])
    s = d["stages_ms"]
    print("ms %.3f" % d["ms_per_step"], "K1", d["roofline"]["ms_parts"], s, d.get("full_size_checks"))
except Exception as e:
    print("ERR", e, open("gpurun_out/r2o_bench.err").read()[-600:])
PY
