#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_e2e --profile > gpurun_out/quick.json 2> gpurun_out/quick.err; python -c "
import json;d=json.load(open('gpurun_out/quick.json'));print(d['ms_per_step'],d['roofline']['ms_parts'],d['roofline']['frac']); print(d['stages_ms'])"; tail -3 gpurun_out/quick.err
