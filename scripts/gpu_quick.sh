#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for c in 8; do
timeout 600 python bench.py --steps 3 --warmup 3 --no_e2e --k1_min_ctas $c > gpurun_out/quick$c.json 2> gpurun_out/quick$c.err; python -c "
import json;d=json.load(open('gpurun_out/quick$c.json'));print('min_ctas',$c,d['ms_per_step'],d['roofline']['ms_parts'],d['roofline']['frac'])"; tail -3 gpurun_out/quick$c.err
done
python -c "
import json;d=json.load(open('gpurun_out/quick8.json'));print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['seconds']);print('cli',d['cli_files_to_files'])"
