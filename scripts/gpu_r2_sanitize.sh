#!/bin/bash
# round 2: compute-sanitizer memcheck over the tests that exercise the kernels added in round 2 (k_vox_block both forms and both
# sort paths, k_feat_front, k_feat_segments<true>, sub-batch streams, odometry with hints, ICP with the neighbour bound)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_memcheck.log; : > $L
for t in "tests/test_voxel_parity.py -k 'block_kernel or scan_features or duplicates'" "tests/test_features_parity.py -k 'fused_front'" "tests/test_frames_parity.py -k 'sub_batches'" "tests/test_stream_parity.py -k 'cloud_info and imu_only'" "tests/test_loop_parity.py"; do
  echo "== $t" >> $L
  eval timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest $t -m gpu -x -q >> $L 2>&1
  echo "rc=$?" >> $L
done
grep -E "^== |rc=|ERROR SUMMARY|passed|failed|Invalid|out of bounds" $L | head -40
