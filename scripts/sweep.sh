#!/bin/bash
# usage: scripts/sweep.sh  -> prints frames/s for cell sizes x register variants
run() { python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['avg_launch_ms'])"; }
cp lis_slam_b200/liblisreg.so /tmp/base.so
for h in 0.5 0.55 0.6 0.7 0.75; do LISREG_CELL=$h run "mb5 cell=$h"; done
for mb in 4 6 8; do cp lis_slam_b200/liblisreg_mb$mb.so lis_slam_b200/liblisreg.so; for h in 0.6 0.7; do LISREG_CELL=$h run "mb$mb cell=$h"; done; done
cp /tmp/base.so lis_slam_b200/liblisreg.so
