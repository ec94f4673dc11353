#!/bin/bash
mkdir -p gpurun_out
T=${1:-r5}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -15 gpurun_out/${T}_tests.log
summ() { python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(sys.argv[2], 'value %.0f e2e %.0f ms/step %.2f e2e_ms %.2f stages %s frac %.4f launch_ms %.4f bitid %s err %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], {k: round(v,2) for k,v in r['stage_ms_per_step'].items()}, r['frac'], r['avg_launch_ms'], d['e2e']['bit_identical_to_device_resident_run'], d.get('pose_err_vs_cpu')))
" "$1" "$2"; }
run() { tag=$1; shift; envs=$1; shift; env $envs timeout 600 python bench.py "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; summ gpurun_out/${T}_$tag.json "[$tag $envs $*]"; tail -2 gpurun_out/${T}_$tag.err; }
run bench ""
run sync "" --e2e-sync --no-cpu
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; cat gpurun_out/${T}_ref.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_knn_search|k_knn_check|k_feat_segments|k_lm_resid|k_rs_scatter|k_vox_centroid' -s 0 -c 12 -o gpurun_out/${T}_prof python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${T}_ncu_full.log 2>&1
