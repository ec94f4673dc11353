#!/bin/bash
run() { python bench.py --steps 6 --warmup 3 --no-cpu --no-latency --e2e-sync 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['roofline']['stage_ms_per_step'])"; }
cp lis_slam_b200/liblisreg.so /tmp/base.so
run base
for v in r6 r8 k10 k12; do cp lis_slam_b200/liblisreg_$v.so lis_slam_b200/liblisreg.so; run $v; done
cp /tmp/base.so lis_slam_b200/liblisreg.so
