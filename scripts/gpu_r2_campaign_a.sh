#!/bin/bash
# round-2 measurement campaign, part A (1 GPU): one bench line per BASELINE config + the config-exact / distinct-maps variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "== $name rc=$? $(wc -c < gpurun_out/$name.json) bytes"; tail -c 300 gpurun_out/$name.err; }
run r2_bench_frames
run r2_ref --impl reference --steps 2 --warmup 1
run r2_bench_frames_b512 --batch 512 --sweeps 64 --no-latency --cpu-check 64 --steps 6
run r2_bench_distinct --distinct-maps --no-latency --cpu-check 16 --steps 6
run r2_stream_hdl64 --workload stream_hdl64 --cpu-frames -1
run r2_stream_vlp16 --workload stream_vlp16 --cpu-frames 200
run r2_loop --workload loop --steps 3 --warmup 1
python - <<'PY'
import json
for f in ("r2_bench_frames","r2_ref","r2_bench_frames_b512","r2_bench_distinct","r2_stream_hdl64","r2_stream_vlp16","r2_loop"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, round(d["value"],1), round(d["e2e"]["value"],1), (d.get("e2e_compact_input") or {}).get("value"), d.get("latency_ms",{}).get("e2e_host_call"),
              (d.get("roofline") or {}).get("stage_ms_per_step"), d.get("stage_ms"), (d.get("roofline") or {}).get("frac"), (d.get("pose_err_vs_cpu") or {}), (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"))
    except Exception as e: print(f, "ERR", e)
PY
