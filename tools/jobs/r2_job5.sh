#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
PSNW_FUSED_WIDE=1 PSNW_FUSED_VARIANT=0 timeout 600 python -m pytest tests/test_gpu_pyramid.py -x -q -k "fused" 2>&1 | tail -2
S="768,576,4,1;768,576,4,4;768,576,3,4;768,576,4,16;1920,1080,4,1"
for w in 0 1; do
  echo wide=$w
  PYR_SWEEP_ONLY="$S" PYR_SWEEP_OUT=gpurun_out/r2_j5_w$w.json PSNW_PYR_MODE=fused PSNW_FUSED_VARIANT=0 PSNW_FUSED_WIDE=$w timeout 300 python tools/pyr_sweep.py 2>&1
done
