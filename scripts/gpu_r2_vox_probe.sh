#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --e2e-sync --no-early --no-compact --no-latency"
timeout 900 ncu --set full --clock-control none -k regex:'k_vox_centroid|k_vox_heads|k_rs_scatter|k_vox_keys' -s 0 -c 8 -o gpurun_out/h_vox $B > gpurun_out/h_ncu.log 2>&1
ncu -i gpurun_out/h_vox.ncu-rep --page raw --csv > gpurun_out/h_vox_raw.csv 2>/dev/null
python scripts/ncu_table.py gpurun_out/h_vox_raw.csv
