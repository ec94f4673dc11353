#!/bin/bash
# round-2 measurement campaign, part C (N GPUs of one box): throughput mode + loop closure through torchrun
# usage: gpu_r2_campaign_c.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 900 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-latency > gpurun_out/r2_n${N}_frames.json 2> gpurun_out/r2_n${N}_frames.err; echo "frames rc=$?"; tail -c 400 gpurun_out/r2_n${N}_frames.err
timeout 900 $TR bench.py --gpus $N --workload loop --steps 3 --warmup 1 > gpurun_out/r2_n${N}_loop.json 2> gpurun_out/r2_n${N}_loop.err; echo "loop rc=$?"; tail -c 400 gpurun_out/r2_n${N}_loop.err
python - <<PY
import json
for f in ("r2_n${N}_frames","r2_n${N}_loop"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), (d.get("e2e_compact_input") or {}).get("value"), d.get("allgather"), d.get("stage_ms"), d["ms_per_step"], (d.get("clocks") or {}).get("reasons"))
    except Exception as e: print(f, "ERR", e)
PY
