#!/bin/bash
# profile capture for profiles/: launch list, DRAM traffic of the LM stage, ncu --set full of the top kernels
mkdir -p gpurun_out
T=${1:-p1}
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-sync > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_knn|k_lm_' -c 140 --csv --log-file gpurun_out/${T}_lm_dram.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-sync > gpurun_out/${T}_ncu_dram.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_knn_search|k_knn_check|k_knn_coop|k_feat_segments|k_lm_resid|k_rs_scatter|k_vox_centroid|k_feat_project|k_feat_compact' -s 0 -c 16 -o gpurun_out/${T}_prof python bench.py --steps 1 --warmup 1 --no-cpu --e2e-sync > gpurun_out/${T}_ncu_full.log 2>&1
cat gpurun_out/${T}_bench.json | cut -c1-400
