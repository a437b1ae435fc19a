#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pyramid.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -3
S="768,576,4,16;768,576,4,32;768,576,4,64;768,576,4,128"
for sp in 0 1; do
  echo split=$sp
  PYR_SWEEP_ONLY="$S" PSNW_PYR_SPLIT=$sp timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200
done
S="1920,1080,4,32;3840,2160,5,16"
PYR_SWEEP_ONLY="$S" PSNW_PYR_MODE=hybrid timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200
