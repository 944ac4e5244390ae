#!/bin/bash
# Round 2, visit M (1 GPU): slot-chunk fragment kernel -- A/B tests, graph-stage tests, bench with frag_stage 1 and 0.
mkdir -p gpurun_out
echo "== pytest (graph-stage tests)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -rs -k "slot_chunk or fragment or graph or seeded or files_match or larger or sharding or huge or window_agg" 2>&1 | tail -8 | tee gpurun_out/r2m_pytest_gpu.log
for st in 1 0; do
  PHZ_OPTIONS=frag_stage=$st timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_e2e --no_wgs > gpurun_out/r2m_bench_fs$st.json 2> gpurun_out/r2m_bench_fs$st.err
  python - $st <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2m_bench_fs%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    s = d["stages_ms"]
    print("frag_stage", sys.argv[1], "ms %.3f" % d["ms_per_step"], {k: s[k] for k in s if k.startswith("graph") or k.startswith("phase.hap")}, d.get("full_size_checks"))
except Exception as e:
    print("ERR", e, open("gpurun_out/r2m_bench_fs%s.err" % sys.argv[1]).read()[-600:])
PY
done
