#!/bin/bash
# Round 2, visit R (1 GPU): cursor atomic off the tile's critical path -- K1 mode equality + goldens, bench.
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k1 or files_match or mapper or seeded or larger or tile_order or config1" 2>&1 | tail -4 | tee gpurun_out/r2r_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_e2e --no_wgs > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
python - <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2r_bench.json").read().strip().splitlines()[-1])
    print("ms %.3f" % d["ms_per_step"], "K1", d["roofline"]["ms_parts"], d.get("full_size_checks"))
except Exception as e:
    print("ERR", e, open("gpurun_out/r2r_bench.err").read()[-600:])
PY
