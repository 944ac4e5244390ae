#!/bin/bash
mkdir -p gpurun_out
PHZ_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no_cpu_baseline --no_e2e > gpurun_out/quick.json 2> gpurun_out/quick.err; grep run_path gpurun_out/quick.err | tail -3
