#!/bin/bash
# One GPU-box visit: parity tests, smoke, a reduced and a full-size bench line, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench 1/10"; timeout 600 python bench.py --pairs 5000000 --variants 200000 --steps 3 --warmup 3 --cpu_pairs 10000 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -c 3000 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
echo "== bench full"; timeout 1200 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --pairs 2000000 --variants 100000 --steps 1 --warmup 1 --no_e2e --no_cpu_baseline > gpurun_out/ncu_bench.log 2>&1; tail -3 gpurun_out/ncu_bench.log
