#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --profile > gpurun_out/prof_full.json 2> gpurun_out/prof_full.err; tail -c 2800 gpurun_out/prof_full.json; tail -3 gpurun_out/prof_full.err
