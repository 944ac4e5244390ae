#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_tile -c 1 -o gpurun_out/k1_tile -f python bench.py --pairs 10000000 --variants 400000 --steps 1 --warmup 0 --no_e2e --no_cpu_baseline --k1_mode 3 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
