#!/bin/bash
# Round 2, visit P (1 GPU): read lists selected by warp tiles, compact edge-table sort key, implicit fragment ids in the
# transport form: full GPU suite, default bench line (with e2e).
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -8 | tee gpurun_out/r2p_pytest_gpu.log
timeout 900 python bench.py --no_cpu_baseline --no_wgs > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
python - <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2p_bench.json").read().strip().splitlines()[-1])
    s = d["stages_ms"]
    print("ms %.3f" % d["ms_per_step"], "K1", d["roofline"]["ms_parts"], s, d.get("full_size_checks"))
    print("e2e", json.dumps(d["e2e"])[:900])
except Exception as e:
    print("ERR", e, open("gpurun_out/r2p_bench.err").read()[-600:])
PY
