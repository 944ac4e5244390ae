#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "windowed or backend" 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
for m in 1 2 3; do
timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --no_e2e --k1_mode $m > gpurun_out/k1_mode$m.json 2> gpurun_out/k1_mode$m.err; python -c "
import json;d=json.load(open('gpurun_out/k1_mode$m.json'));print('k1_mode',$m,d['ms_per_step'],d['roofline']['ms_count_scan_emit'],d['roofline']['frac'])"; tail -3 gpurun_out/k1_mode$m.err
done
