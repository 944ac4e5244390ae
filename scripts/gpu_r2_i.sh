#!/bin/bash
# Round 2, visit I (1 GPU): full GPU suite, bench with both K1 register budgets, ncu captures for profiles/.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -6 | tee gpurun_out/r2i_pytest_gpu.log
for m in 8 6; do
timeout 900 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_e2e --no_wgs --k1_min_ctas $m > gpurun_out/r2i_bench_m$m.json 2> gpurun_out/r2i_bench_m$m.err; tail -2 gpurun_out/r2i_bench_m$m.err | cut -c1-300
python - $m <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2i_bench_m%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("min_ctas", sys.argv[1], "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "K1", d["roofline"]["ms_parts"], {k: v for k, v in d["stages_ms"].items() if k.startswith("phase")})
PY
done
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no_e2e --no_cpu_baseline --no_wgs --profiler_range > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-120
echo "== ncu full: K1 tile kernel + fragment kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_tile|fragment_kernel' -c 2 -o gpurun_out/r2_k1_frag_full -f python bench.py --steps 1 --warmup 0 --no_e2e --no_cpu_baseline --no_wgs > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-120
