#!/bin/bash
# Round 2, visit L (1 GPU): files-to-files at whole-genome scale (configs[1]: 50 M pairs x 2 M het SNVs) against the
# unmodified reference on the same records.
mkdir -p gpurun_out
df -h /tmp | tail -1; free -g | head -2
PHZ_IO_TIMING=1 timeout 1700 python scripts/files_to_files.py --pairs 50000000 --variants 2000000 --ref_timeout 1200 > gpurun_out/r2l_f2f_50m.json 2> gpurun_out/r2l_f2f_50m.err; echo rc $?; tail -12 gpurun_out/r2l_f2f_50m.err | cut -c1-200; cut -c1-1200 gpurun_out/r2l_f2f_50m.json
