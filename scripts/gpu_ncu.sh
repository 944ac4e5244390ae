#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_window -c 2 -o gpurun_out/k1_window -f python bench.py --pairs 10000000 --variants 400000 --steps 1 --warmup 0 --no_e2e --no_cpu_baseline --k1_mode 1 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:for_each_kernel -s 1 -c 3 -o gpurun_out/k1_generic -f python bench.py --pairs 10000000 --variants 400000 --steps 1 --warmup 0 --no_e2e --no_cpu_baseline --k1_mode 0 > gpurun_out/ncu0.log 2>&1; tail -2 gpurun_out/ncu0.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
