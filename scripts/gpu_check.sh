#!/bin/bash
# One GPU-box visit: parity tests, smoke, both bench arms at full size, ncu launch list + full capture of K1.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-700 gpurun_out/bench_reference.json; tail -3 gpurun_out/bench_reference.err
echo "== bench full"; timeout 1500 python bench.py --profile > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 4000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
echo "== ncu launch list (our kernels + CUB passes of one step)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no_e2e --no_cpu_baseline --profiler_range > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
echo "== ncu full capture of the K1 tile kernel at full size"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_tile -c 1 -o gpurun_out/k1_tile_full -f python bench.py --steps 1 --warmup 0 --no_e2e --no_cpu_baseline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200
python - <<'PY'
import json
for ln in open('gpurun_out/ncu_full.log'):
    if ln.startswith('{'):
        json.dump({"records": json.loads(ln)["config"]["records"]}, open('gpurun_out/k1_tile_full.meta.json','w'))
PY
