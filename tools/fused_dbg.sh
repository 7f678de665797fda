#!/bin/bash
# Phase-stamp breakdown of the fused block kernel for the given layers (HFB_FUSED_DBG), one process per layer.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for l in ${LAYERS:-3 4}; do
  echo "=== layer $l"
  HFB_NO_GRAPH=1 HFB_CPL_DBG=$l HFB_FUSED_DBG=$l STEPS=2 timeout 120 python tools/profile_step.py 2>&1 | grep -E "layer_$l |   \[|launches"
done
