#!/bin/bash
run() { python bench.py --steps 4 --warmup 3 --no-cpu $2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), d['roofline']['stage_ms_per_step'])"; }
run "frame cell=default" ""
for h in 0.5 0.7 0.8; do LISREG_CELL=$h run "frame cell=$h" ""; done
run "lm cell=default" "--stage lm --batch 512"
