#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pyramid.py -x -q > gpurun_out/r2_j2_pytest.log 2>&1
tail -5 gpurun_out/r2_j2_pytest.log
S="768,576,4,1;768,576,4,4;768,576,4,16;768,576,4,64;768,576,3,4;1920,1080,4,1;1920,1080,4,8;1920,1080,4,32;3840,2160,5,1;3840,2160,5,4"
for t in 512 1024; do
  PYR_SWEEP_ONLY="$S" PYR_SWEEP_OUT=gpurun_out/r2_j2_fused$t.json PSNW_PYR_MODE=fused PSNW_FUSED_THREADS=$t timeout 300 python tools/pyr_sweep.py > gpurun_out/r2_j2_fused$t.log 2>&1
done
PYR_SWEEP_ONLY="$S" PYR_SWEEP_OUT=gpurun_out/r2_j2_auto.json timeout 300 python tools/pyr_sweep.py > gpurun_out/r2_j2_auto.log 2>&1
cat gpurun_out/r2_j2_fused512.log gpurun_out/r2_j2_fused1024.log
