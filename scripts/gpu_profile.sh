#!/bin/bash
# stage-level profile + ncu launch list restricted to our kernels and CUB passes
mkdir -p gpurun_out
timeout 900 python bench.py --pairs 5000000 --variants 200000 --steps 2 --warmup 2 --no_cpu_baseline --no_e2e --profile > gpurun_out/prof_small.json 2> gpurun_out/prof_small.err; tail -c 2500 gpurun_out/prof_small.json; tail -3 gpurun_out/prof_small.err
timeout 1200 python bench.py --steps 2 --warmup 2 --no_cpu_baseline --no_e2e --profile > gpurun_out/prof_full.json 2> gpurun_out/prof_full.err; tail -c 2500 gpurun_out/prof_full.json; tail -3 gpurun_out/prof_full.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'for_each_kernel|as_hist|DeviceRadixSort|DeviceScan' -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --pairs 5000000 --variants 200000 --steps 1 --warmup 1 --no_e2e --no_cpu_baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
