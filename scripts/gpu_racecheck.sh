#!/bin/bash
# compute-sanitizer memcheck over the parity tests that exercise the list / cooperative / chunked paths
mkdir -p gpurun_out
T=${1:-san}
for t in "tests/test_lm_parity.py -k 'knn5 or check_path or variant_a_matches or batch_matches or empty_and_tiny or map_distance'" "tests/test_features_parity.py" "tests/test_frames_parity.py -k 'chunked or submit'"; do
  echo "== $t" >> gpurun_out/${T}_racecheck.log
  eval timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest $t -m gpu -x -q >> gpurun_out/${T}_racecheck.log 2>&1
  echo "rc=$?" >> gpurun_out/${T}_racecheck.log
done
grep -E "^== |rc=|ERROR SUMMARY|passed|failed|hazard|Race" gpurun_out/${T}_racecheck.log | head -60
