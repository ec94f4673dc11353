#!/bin/bash
# one GPU call: parity tests, both bench arms, ncu launch list, ncu --set full of the two LM kernels
mkdir -p gpurun_out
T=${1:-r01b}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lm_knn|k_lm_resid' -s 20 -c 2 -o gpurun_out/${T}_lm python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${T}_ncu_full.log 2>&1
tail -3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_bench.json; cat gpurun_out/${T}_ref.json
