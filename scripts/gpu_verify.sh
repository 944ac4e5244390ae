#!/bin/bash
# Light GPU-box visit: parity tests, smoke, both bench arms (no ncu).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-700 gpurun_out/bench_reference.json; tail -3 gpurun_out/bench_reference.err
echo "== bench full"; PHZ_TRACE=1 timeout 1500 python bench.py --profile > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 4000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
