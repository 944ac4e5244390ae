#!/bin/bash
# Round 2, visit W (1 GPU): the default bench line as the driver runs it (cpu_baseline, WGS shape, command-line legs), the
# ncu launch list of one step and full captures of the two largest kernels, for profiles/.
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; tail -2 gpurun_out/r2w_bench_n1.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2w_bench_n1.json").read().strip().splitlines()[-1])
    print("value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["ms_per_step"], "syncs", d.get("host_syncs_per_step"), "launches", d["gpu_launches"], d["library_passes"])
    print("  roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "ms", "dram_frac")})
    print("  cpu_baseline", json.dumps(d["cpu_baseline"])[:500])
    print("  wgs", json.dumps({k: v for k, v in d["wgs_shape"].items() if k != "stages_ms"})[:900])
    c = d["cli_files_to_files"]; print("  cli", c["from_sam_text"]["seconds"], c["from_bgzf_bam"]["seconds"], c["cli_parity"])
except Exception as e:
    print("ERR", e)
PY
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 1 --warmup 1 --no_e2e --no_cpu_baseline --no_wgs --profiler_range > gpurun_out/r2w_ncu_launches.log 2>&1; tail -1 gpurun_out/r2w_ncu_launches.log | cut -c1-120
echo "== ncu full: K1 tile kernel + slot-chunk fragment kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_tile|fragment_slots' -c 2 -o gpurun_out/r2w_k1_frag_full -f python bench.py --steps 1 --warmup 0 --no_e2e --no_cpu_baseline --no_wgs > gpurun_out/r2w_ncu_full.log 2>&1; tail -1 gpurun_out/r2w_ncu_full.log | cut -c1-120
ls -la gpurun_out/r2w_*
