#!/bin/bash
# Round 2, visit A (2 GPUs): parity suite, NCCL sharded parity, sharded bench small -> full at N=2, N=1 full for comparison.
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2a_gpus.txt; nproc >> gpurun_out/r2a_gpus.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -15 | tee gpurun_out/r2a_pytest_gpu.log
echo "== sharded bench, small, N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 3 --warmup 3 --pairs 2000000 --variants 80000 --no_cpu_baseline > gpurun_out/r2a_bench_small_n2.json 2> gpurun_out/r2a_bench_small_n2.err; tail -c 3000 gpurun_out/r2a_bench_small_n2.json; tail -5 gpurun_out/r2a_bench_small_n2.err | cut -c1-400
echo "== sharded bench, full, N=2"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2a_bench_n2.json 2> gpurun_out/r2a_bench_n2.err; tail -c 6000 gpurun_out/r2a_bench_n2.json; tail -5 gpurun_out/r2a_bench_n2.err | cut -c1-400
echo "== bench, full, N=1 (with CPU baseline + CLI parity)"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; tail -c 5000 gpurun_out/r2a_bench_n1.json; tail -5 gpurun_out/r2a_bench_n1.err | cut -c1-400
