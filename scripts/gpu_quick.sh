#!/bin/bash
# quick GPU check: parity tests + bench (value/e2e) for a few settings.  usage: gpu_quick.sh TAG ["ENV=.. ENV=.." ...]
mkdir -p gpurun_out
T=${1:-q}; shift
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
summ() { python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(sys.argv[2], 'value %.0f e2e %.0f ms/step %.2f e2e_ms %.2f stages %s frac %.4f launch_ms %.4f bitid %s err %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], {k: round(v,2) for k,v in r['stage_ms_per_step'].items()}, r['frac'], r['avg_launch_ms'], d['e2e']['bit_identical_to_device_resident_run'], d.get('pose_err_vs_cpu')))
" "$1" "$2"; }
i=0
if [ $# -eq 0 ]; then set -- ""; fi
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 600 python bench.py --cpu-sample 4 > gpurun_out/${T}_bench$i.json 2> gpurun_out/${T}_bench$i.err
  summ gpurun_out/${T}_bench$i.json "[$envs]"; tail -2 gpurun_out/${T}_bench$i.err
done
