#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --e2e-sync --no-early --no-compact --no-latency"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/g_launches_frames.csv $B > gpurun_out/g_ncu_launch.log 2>&1
python scripts/launch_summary.py gpurun_out/g_launches_frames.csv > gpurun_out/g_launches_frames.md; cat gpurun_out/g_launches_frames.md
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/g_launches_hdl64.csv python bench.py --workload stream_hdl64 --frames 12 --no-cpu --no-graph > gpurun_out/g_ncu_hdl64.log 2>&1
python scripts/launch_summary.py gpurun_out/g_launches_hdl64.csv > gpurun_out/g_launches_hdl64.md; cat gpurun_out/g_launches_hdl64.md
