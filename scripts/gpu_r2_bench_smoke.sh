#!/bin/bash
# round-2 bench smoke: every workload at a reduced size (debug run, numbers not for the record)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --workload stream_vlp16 --frames 80 --cpu-frames 24 > gpurun_out/s_vlp16.json 2> gpurun_out/s_vlp16.err
python bench.py --workload stream_hdl64 --frames 48 --cpu-frames 16 > gpurun_out/s_hdl64.json 2> gpurun_out/s_hdl64.err
python bench.py --workload loop --loop-n 1500 --steps 2 --loop-targets 2 > gpurun_out/s_loop.json 2> gpurun_out/s_loop.err
python bench.py --steps 4 --cpu-check 32 > gpurun_out/s_frames.json 2> gpurun_out/s_frames.err
tail -c 1500 gpurun_out/s_vlp16.err gpurun_out/s_hdl64.err gpurun_out/s_loop.err gpurun_out/s_frames.err
