#!/bin/bash
# A/B of the extract(+associate) step on the SAME box: tools/ab/lib_base.so (baseline build) against the in-tree library.
cd "$(dirname "$0")/.."
for rep in 1 2; do
  for lib in tools/ab/lib_base.so hfnet_slam_b200/libhfnet_b200.so; do
    echo "== $lib"
    HFB_LIB=$PWD/$lib PROFILE=${PROFILE:-0} python tools/step_time.py 2>&1 | grep -E "STEP|fused|global|sum"
    HFB_LIB=$PWD/$lib BATCH=1 PROFILE=0 python tools/step_time.py 2>&1 | grep STEP
  done
done
