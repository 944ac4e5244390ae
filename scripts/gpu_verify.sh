#!/bin/bash
# Light GPU-box visit: parity tests, smoke, our bench arm with stage profile (no ncu, no reference arm).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench full"; PHZ_TRACE=1 timeout 1500 python bench.py --profile ${BENCH_ARGS} > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 5000 gpurun_out/bench_full.json; tail -8 gpurun_out/bench_full.err
