#!/bin/bash
run() { python bench.py --steps 4 --warmup 3 --no-cpu $2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), d['roofline']['stage_ms_per_step'])"; }
run "split k12 frame" ""
run "split k12 lm" "--stage lm --batch 512"
LISREG_LM_FUSED=1 run "fused frame" ""
cp lis_slam_b200/liblisreg.so /tmp/base.so
for v in k8 k16; do cp lis_slam_b200/liblisreg_$v.so lis_slam_b200/liblisreg.so; run "split $v frame" ""; run "split $v lm" "--stage lm --batch 512"; done
cp /tmp/base.so lis_slam_b200/liblisreg.so
