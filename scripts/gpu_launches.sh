#!/bin/bash
mkdir -p gpurun_out
T=${1:-l1}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-sync --no-latency > gpurun_out/${T}_ncu_launch.log 2>&1
tail -3 gpurun_out/${T}_ncu_launch.log | cut -c1-300
