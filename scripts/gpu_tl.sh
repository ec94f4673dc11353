#!/bin/bash
mkdir -p gpurun_out
T=${1:-tl}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-sync --no-latency > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 600 python bench.py --cpu-sample 4 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.0f e2e %.0f ms %.2f e2e_ms %.2f stages %s frac %.4f lat %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], {k: round(v,2) for k,v in r['stage_ms_per_step'].items()}, r['frac'], d.get('single_frame_latency')))"
tail -2 gpurun_out/${T}_bench.err
