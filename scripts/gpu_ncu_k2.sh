#!/bin/bash
# ncu --set full of the graph-stage kernels (first step) at full size
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'build_graph|variant_stats' -c 20 -o gpurun_out/k2_full -f python bench.py --steps 1 --warmup 0 --no_e2e --no_cpu_baseline > gpurun_out/ncu_k2.log 2>&1; tail -2 gpurun_out/ncu_k2.log | cut -c1-200
ls -la gpurun_out/k2_full.ncu-rep
