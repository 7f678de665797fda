#!/bin/bash
# compute-sanitizer memcheck + racecheck over the kernels changed in the last session of round 2 (local BA, keyframe
# database query, undistortion), through their GPU parity tests and smoke().
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/sanitizer_r2s3.txt
: > $out
for tool in memcheck racecheck; do
  for t in "tests/test_lba_gpu.py tests/test_pose_gpu.py" "tests/test_kfdb_gpu.py -k not(two_processes)" "tests/test_undistort_gpu.py -k points"; do
    echo "=== compute-sanitizer --tool $tool: pytest $t" >> $out
    HFB_NO_GRAPH=1 timeout 900 compute-sanitizer --tool $tool python -m pytest $t -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" | sort | uniq -c | sort -rn | head -8 >> $out
  done
  echo "=== compute-sanitizer --tool $tool: __graft_entry__.py --smoke" >> $out
  HFB_NO_GRAPH=1 timeout 600 compute-sanitizer --tool $tool python __graft_entry__.py --smoke 2>&1 | grep -E "smoke ok|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" | sort | uniq -c | sort -rn | head -8 >> $out
done
cat $out
