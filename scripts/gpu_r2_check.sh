#!/bin/bash
# round-2 quick check: GPU tests + headline bench + short streams
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 6 --no-cpu --no-latency --no-early --no-compact > gpurun_out/c_frames.json 2> gpurun_out/c_frames.err; tail -c 400 gpurun_out/c_frames.err
python bench.py --workload stream_hdl64 --frames 120 --cpu-frames 16 > gpurun_out/c_hdl64.json 2> gpurun_out/c_hdl64.err; tail -c 400 gpurun_out/c_hdl64.err
python bench.py --workload stream_vlp16 --frames 200 --cpu-frames 16 > gpurun_out/c_vlp16.json 2> gpurun_out/c_vlp16.err; tail -c 400 gpurun_out/c_vlp16.err
python bench.py --workload loop --loop-n 5000 --steps 3 --loop-targets 4 > gpurun_out/c_loop.json 2> gpurun_out/c_loop.err; tail -c 400 gpurun_out/c_loop.err
python - <<'PY'
import json
for f in ("c_frames","c_hdl64","c_vlp16","c_loop"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, round(d["value"],1), round(d["e2e"]["value"],1), d.get("latency_ms",{}).get("e2e_host_call"), d["roofline"].get("stage_ms_per_step"), d.get("stage_ms"), round(d["roofline"]["frac"],4), (d.get("pose_err_vs_cpu") or {}).get("within_tolerance"), d.get("parity_check"))
    except Exception as e: print(f, "ERR", e)
PY
