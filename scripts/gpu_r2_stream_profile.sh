#!/bin/bash
# launch list (per-kernel device time) of the streaming workloads, eager launches so that ncu sees every kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in stream_hdl64 stream_vlp16; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_$w.csv \
      python bench.py --workload $w --frames 24 --no-cpu --no-graph > gpurun_out/r2_ncu_$w.log 2>&1
  python scripts/launch_summary.py gpurun_out/r2_launches_$w.csv > gpurun_out/r2_launches_$w.md 2>/dev/null
  python scripts/launch_summary.py gpurun_out/r2_launches_$w.csv median > gpurun_out/r2_launches_${w}_median.md 2>/dev/null
  tail -45 gpurun_out/r2_launches_${w}_median.md; tail -12 gpurun_out/r2_launches_$w.md
done
