#!/bin/bash
# Round 2, visit N (1 GPU): staged K1 emission + per-slot haplotypic counts + slot-chunk fragment kernel: full GPU suite,
# bench with the K1 emission both ways.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -8 | tee gpurun_out/r2n_pytest_gpu.log
for st in 1 0; do
  PHZ_OPTIONS=k1_staged_emit=$st timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_e2e --no_wgs > gpurun_out/r2n_bench_se$st.json 2> gpurun_out/r2n_bench_se$st.err
  python - $st <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2n_bench_se%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    s = d["stages_ms"]
    print("k1_staged_emit", sys.argv[1], "ms %.3f" % d["ms_per_step"], "K1", d["roofline"]["ms_parts"], s, d.get("full_size_checks"))
except Exception as e:
    print("ERR", e, open("gpurun_out/r2n_bench_se%s.err" % sys.argv[1]).read()[-600:])
PY
done
