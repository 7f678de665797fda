#!/bin/bash
# Runs every GPU parity test file in its own process (a trapped kernel poisons its CUDA context) and collects logs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for t in ${TESTS:-test_gemm_gpu test_postproc_gpu test_match_gpu test_projection_gpu test_kfdb_gpu test_lba_gpu test_pose_gpu test_distinctive_gpu test_extract_gpu test_golden_gpu test_shim_gpu}; do
  echo "=== $t" | tee -a gpurun_out/summary.txt
  timeout ${TEST_TIMEOUT:-600} python -m pytest tests/$t.py -m gpu -q --no-header -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/$t.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 40 gpurun_out/$t.log | tee -a gpurun_out/summary.txt
done
