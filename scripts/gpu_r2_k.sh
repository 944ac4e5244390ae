#!/bin/bash
# Round 2, visit K (1 GPU): full GPU suite, default bench line, files-to-files at 20 M records after the commit-side
# ranking and the parallel ingest.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -6 | tee gpurun_out/r2k_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; tail -2 gpurun_out/r2k_bench_n1.err | cut -c1-300
python - <<'PY'
import json
for f in ("r2k_bench_n1",):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["ms_per_step"], d["e2e"]["single_sample_ms"], "syncs", d.get("host_syncs_per_step"), "launches", d["gpu_launches"], d["library_passes"])
        print("  stages", d["stages_ms"])
        print("  roofline", d["roofline"])
        print("  cli", json.dumps(d.get("cli_files_to_files"))[:1500])
    except Exception as e:
        print(f, "ERR", e)
PY
PHZ_IO_TIMING=1 timeout 1200 python scripts/files_to_files.py --pairs 10000000 --variants 400000 > gpurun_out/r2k_f2f_10m.json 2> gpurun_out/r2k_f2f_10m.err; tail -8 gpurun_out/r2k_f2f_10m.err | cut -c1-200; cut -c1-1500 gpurun_out/r2k_f2f_10m.json
