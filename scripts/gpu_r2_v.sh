#!/bin/bash
# Round 2, visit V (1 GPU): asynchronous noise read (no host wait between the commits and the graph stage), resident helper
# thread: full GPU suite, default bench line.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rs --timeout 600 2>&1 | tail -8 | tee gpurun_out/r2v_pytest_gpu.log
timeout 900 python bench.py --no_cpu_baseline --no_wgs > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
python - <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2v_bench.json").read().strip().splitlines()[-1])
    s = d["stages_ms"]
    print("ms %.3f" % d["ms_per_step"], "syncs", d.get("host_syncs_per_step"), "K1", d["roofline"]["ms_parts"], s, d.get("full_size_checks"))
    print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["single_sample_ms"])
except Exception as e:
    print("ERR", e, open("gpurun_out/r2v_bench.err").read()[-600:])
PY
