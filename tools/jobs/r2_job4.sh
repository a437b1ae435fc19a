#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in 0 1 2; do
PSNW_FUSED_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_pyramid.py -x -q -k "fused" > gpurun_out/r2_j4_pytest$v.log 2>&1
tail -3 gpurun_out/r2_j4_pytest$v.log
done
S="768,576,4,1;768,576,4,4;768,576,4,16;768,576,4,64;768,576,3,4;1920,1080,4,1;1920,1080,4,8;1920,1080,4,32;3840,2160,5,1;3840,2160,5,4;3840,2160,5,16"
for v in 0 1 2; do
  PYR_SWEEP_ONLY="$S" PYR_SWEEP_OUT=gpurun_out/r2_j4_fused_v$v.json PSNW_PYR_MODE=fused PSNW_FUSED_VARIANT=$v timeout 300 python tools/pyr_sweep.py > gpurun_out/r2_j4_fused_v$v.log 2>&1
done
cat gpurun_out/r2_j4_fused_v*.log
