#!/bin/bash
# Round 2, visit G (1 GPU): full default bench (CPU baseline, CLI legs from SAM text and BGZF BAM, WGS leg) + drop-in surface tests.
mkdir -p gpurun_out
echo "== pytest -m gpu (cli / io / gene_ae / native vcf)"; timeout 900 python -m pytest tests/test_cli_and_io.py tests/test_gene_ae.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2g_pytest_gpu.log
echo "== bench default"
PHZ_IO_TIMING=1 timeout 1500 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -12 gpurun_out/r2g_bench.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms %.3f" % (d["value"], d["ms_per_step"]))
print("cpu", d["cpu_baseline"]["seconds"], d["cpu_baseline"]["value"])
c = d["cli_files_to_files"]
for k in ("from_sam_text", "from_bgzf_bam"):
    print(k, c[k]["seconds"], c[k]["cli_parity"], c[k]["stage_seconds"])
print("wgs", d["wgs_shape"].get("ms_per_step"), d["wgs_shape"].get("error"))
PY
