#!/bin/bash
# Round 2, visit B (1 GPU): ncu --set full of the fragment kernel at full size (+ source lines), launch list of one step.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fragment_kernel' -c 1 -o gpurun_out/r2_frag_full -f python bench.py --steps 1 --warmup 0 --no_e2e --no_cpu_baseline > gpurun_out/r2b_ncu_frag.log 2>&1; tail -2 gpurun_out/r2b_ncu_frag.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 1 --warmup 1 --no_e2e --no_cpu_baseline --profiler_range > gpurun_out/r2b_ncu_launches.log 2>&1; tail -1 gpurun_out/r2b_ncu_launches.log | cut -c1-200
ls -la gpurun_out/r2_frag_full.ncu-rep
