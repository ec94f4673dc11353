#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
run() { env LISREG_CELL=$1 python bench.py --steps 6 --warmup 3 --no-cpu --no-latency --e2e-sync 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cell $1', round(d['value']), d['roofline']['stage_ms_per_step'])"; }
for h in 0.45 0.5 0.55 0.6 0.7; do run $h; done
