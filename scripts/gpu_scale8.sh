#!/bin/bash
# the bench at N = 8 exactly as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2; nproc
n=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; python -c "
import json;d=json.load(open('gpurun_out/bench_n$n.json'));print('N',$n,'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['single_sample_ms'])"; tail -3 gpurun_out/bench_n$n.err | cut -c1-300; free -g | head -2
