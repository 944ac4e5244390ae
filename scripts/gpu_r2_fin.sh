#!/bin/bash
# Round 2, final visit (1 GPU): full GPU suite and the default bench line of the final state
#
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rs --timeout 600 2>&1 | tail -8 | tee gpurun_out/r2fin_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2fin_bench.json 2> gpurun_out/r2fin_bench.err
python - <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2fin_bench.json").read().strip().splitlines()[-1])
    s = d["stages_ms"]
    print("ms %.3f" % d["ms_per_step"], "syncs", d.get("host_syncs_per_step"), "K1", d["roofline"]["ms_parts"], s, d.get("full_size_checks"))
    print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["single_sample_ms"])
except Exception as e:
    print("ERR", e, open("gpurun_out/r2fin_bench.err").read()[-600:])
PY
