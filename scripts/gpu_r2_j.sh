#!/bin/bash
# Round 2, visit J (2 GPUs): full GPU suite, bench N=1 / N=2, files-to-files at 20 M records with the ingest changes.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -6 | tee gpurun_out/r2j_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_wgs > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; tail -2 gpurun_out/r2j_bench_n1.err | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 10 --warmup 3 --no_replicas > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; tail -2 gpurun_out/r2j_bench_n2.err | cut -c1-300
python - <<'PY'
import json
for f in ("r2j_bench_n1", "r2j_bench_n2"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["ms_per_step"], d["e2e"]["single_sample_ms"], "syncs", d.get("host_syncs_per_step"), "launches", d["gpu_launches"], d["library_passes"])
        print("  stages", d["stages_ms"])
        print("  sharding", json.dumps(d.get("sharding", {}).get("collectives_ms_rank0_one_step_synchronised")), d.get("full_size_checks"))
    except Exception as e:
        print(f, "ERR", e)
PY
PHZ_IO_TIMING=1 timeout 1200 python scripts/files_to_files.py --pairs 10000000 --variants 400000 > gpurun_out/r2j_f2f_10m.json 2> gpurun_out/r2j_f2f_10m.err; tail -4 gpurun_out/r2j_f2f_10m.err | cut -c1-200; cut -c1-1500 gpurun_out/r2j_f2f_10m.json
