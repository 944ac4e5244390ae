#!/bin/bash
# 2-GPU visit: NCCL contig-sharded parity run + the bench at N=2 (weak scaling, one sample per GPU)
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
for case in rna_two_bams quirks; do
  PHZ_ENGINE=gpu timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/shard_worker.py $case 2>&1 | grep -E "SHARDED|Error|error|Traceback" | head -5
done
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
