#!/bin/bash
# Round 2, visit Y (1 GPU): sweep of the slot-chunk fragment kernel (slots per CTA x resident CTAs per SM; tuning builds of
# the same sources: -DPHZ_FRAG_SLOTS / -DPHZ_FRAG_SLOT_CTAS).
mkdir -p gpurun_out
for lib in phaser_b200/_phz.so phaser_b200/_phz_s*.so; do
  PHZ_LIB=$PWD/$lib timeout 600 python bench.py --steps 5 --warmup 2 --no_cpu_baseline --no_e2e --no_wgs > gpurun_out/r2y_tmp.json 2> gpurun_out/r2y_tmp.err
  python - "$lib" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2y_tmp.json").read().strip().splitlines()[-1])
    s = d["stages_ms"]
    print(sys.argv[1], "ms %.3f" % d["ms_per_step"], {k: s[k] for k in s if k.startswith("graph")})
except Exception as e:
    print(sys.argv[1], "ERR", e, open("gpurun_out/r2y_tmp.err").read()[-300:])
PY
done 2>&1 | tee gpurun_out/r2y_sweep.txt
