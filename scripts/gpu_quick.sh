#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "windowed or backend" 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for m in 3; do
timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --no_e2e --k1_mode $m --profile > gpurun_out/k1_mode$m.json 2> gpurun_out/k1_mode$m.err; python -c "
import json;d=json.load(open('gpurun_out/k1_mode$m.json'));print('k1_mode',$m,d['ms_per_step'],d['roofline']['ms_parts'],d['roofline']['frac']); print(d['stages_ms'])"; tail -3 gpurun_out/k1_mode$m.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_tile -c 1 -o gpurun_out/k1_tile -f python bench.py --pairs 10000000 --variants 400000 --steps 1 --warmup 0 --no_e2e --no_cpu_baseline --k1_mode 3 > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log | cut -c1-100
