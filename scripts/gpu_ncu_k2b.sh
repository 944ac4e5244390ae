#!/bin/bash
# ncu --set full of the two shared-memory window kernels of the graph stage (first step, full size)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'variant_lists_kernel|entry_stats_kernel' -c 2 -o gpurun_out/k2_window -f python bench.py --steps 1 --warmup 0 --no_e2e --no_cpu_baseline > gpurun_out/ncu_k2b.log 2>&1; tail -2 gpurun_out/ncu_k2b.log | cut -c1-200
