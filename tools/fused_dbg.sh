#!/bin/bash
# Phase-stamp breakdown of the fused block kernel for the given layers (HFB_FUSED_DBG), one process per layer.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for l in ${LAYERS:-3 4 7 8 9 13 16}; do
  echo "=== layer $l"
  HFB_TRACE=1 HFB_FUSED_DBG=$l STEPS=2 timeout 120 python tools/profile_step.py 2>&1 | grep -E "fused layer_$l|launches"
done
