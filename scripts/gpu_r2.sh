#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -15 gpurun_out/${T}_tests.log
summ() { python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(sys.argv[2], 'value %.0f e2e %.0f ms/step %.2f e2e_ms %.2f stages %s frac %.4f launch_ms %.4f bitid %s err %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], {k: round(v,2) for k,v in r['stage_ms_per_step'].items()}, r['frac'], r['avg_launch_ms'], d['e2e']['bit_identical_to_device_resident_run'], d.get('pose_err_vs_cpu')))
" "$1" "$2"; }
run() { tag=$1; shift; envs=$1; shift; env $envs timeout 600 python bench.py --cpu-sample 4 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; summ gpurun_out/${T}_$tag.json "[$tag $envs $*]"; tail -2 gpurun_out/${T}_$tag.err; }
run b256 "LISREG_E2E_CHUNK=0"
run b256c128 "LISREG_E2E_CHUNK=128" --no-cpu
run noskip "LISREG_KNN_NOSKIP=1 LISREG_E2E_CHUNK=0" --no-cpu
run b128 "LISREG_E2E_CHUNK=0" --batch 128 --no-cpu
run b32 "LISREG_E2E_CHUNK=0" --batch 32 --no-cpu
run b8 "LISREG_E2E_CHUNK=0" --batch 8 --no-cpu
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_knn_search|k_knn_check|k_feat_segments|k_lm_resid' -s 0 -c 9 -o gpurun_out/${T}_prof python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${T}_ncu_full.log 2>&1
