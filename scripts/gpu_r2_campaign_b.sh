#!/bin/bash
# round-2 measurement campaign, part B (1 GPU, under ncu: numbers printed by these runs are NOT bench values)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LISREG_DEV_SPLIT=0   # whole-batch launches: the unit the roofline and the traffic figures are quoted per
B="python bench.py --steps 1 --warmup 1 --no-cpu --e2e-sync --no-early --no-compact --no-latency"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_frames.csv $B > gpurun_out/r2_ncu_launch.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_frames.csv > gpurun_out/r2_launches_frames.md; tail -32 gpurun_out/r2_launches_frames.md
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'k_knn|k_lm_' -c 140 --csv --log-file gpurun_out/r2_lm_dram.csv $B > gpurun_out/r2_ncu_dram.log 2>&1
python scripts/make_traffic.py gpurun_out/r2_lm_dram.csv gpurun_out/r2_lm_iter_traffic.json | head -5
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'k_knn|k_lm_' -c 140 --csv --log-file gpurun_out/r2_lm_dram_distinct.csv $B --distinct-maps > gpurun_out/r2_ncu_dram_distinct.log 2>&1
python scripts/make_traffic.py gpurun_out/r2_lm_dram_distinct.csv gpurun_out/r2_lm_iter_traffic_distinct.json | head -5
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_feat|k_vox|k_rs' -c 40 --csv --log-file gpurun_out/r2_featvox_dram.csv $B > gpurun_out/r2_ncu_featvox.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_knn_search|k_knn_check|k_feat_segments|k_lm_resid|k_lm_solve|k_vox_block|k_feat_project|k_feat_compact|k_feat_gather' -s 0 -c 14 -o gpurun_out/r2_prof $B > gpurun_out/r2_ncu_full.log 2>&1
ncu -i gpurun_out/r2_prof.ncu-rep --page raw --csv > gpurun_out/r2_prof_raw.csv 2>/dev/null
python scripts/ncu_table.py gpurun_out/r2_prof_raw.csv > gpurun_out/r2_prof_table.md; cat gpurun_out/r2_prof_table.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_epsc_score' -s 1 -c 1 -o gpurun_out/r2_prof_epsc python bench.py --workload loop --loop-n 5000 --steps 1 --warmup 1 --no-cpu --loop-max-pairs 4 --loop-targets 1 > gpurun_out/r2_ncu_epsc.log 2>&1
ncu -i gpurun_out/r2_prof_epsc.ncu-rep --page raw --csv > gpurun_out/r2_prof_epsc_raw.csv 2>/dev/null
python scripts/ncu_table.py gpurun_out/r2_prof_epsc_raw.csv
ls -la gpurun_out/*.ncu-rep
