#!/bin/bash
# loop-closure line only, N GPUs of one box (usage: gpu_r2_campaign_c_loop.sh N)
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 900 $TR bench.py --gpus $N --workload loop --steps 3 --warmup 1 > gpurun_out/r2_n${N}_loop.json 2> gpurun_out/r2_n${N}_loop.err; echo "loop rc=$?"; tail -c 300 gpurun_out/r2_n${N}_loop.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_n${N}_loop.json") if l.startswith("{")][-1])
print(d["n_gpus"], round(d["value"],1), d.get("stage_ms"), d["ms_per_step"])
PY
