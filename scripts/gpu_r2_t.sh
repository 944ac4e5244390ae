#!/bin/bash
# Round 2, visit T (1 GPU): adjacency by counting sort, merge changes (expand_runs entry point): full GPU suite, bench with
# e2e.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -8 | tee gpurun_out/r2t_pytest_gpu.log
timeout 900 python bench.py --no_cpu_baseline --no_wgs > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
python - <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2t_bench.json").read().strip().splitlines()[-1])
    s = d["stages_ms"]
    print("ms %.3f" % d["ms_per_step"], "K1", d["roofline"]["ms_parts"], s, d.get("full_size_checks"))
    print("e2e", json.dumps(d["e2e"])[:900])
except Exception as e:
    print("ERR", e, open("gpurun_out/r2t_bench.err").read()[-600:])
PY
