#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pyramid.py -x -q 2>&1 | tail -3
S="768,576,4,1;768,576,4,4;768,576,4,16;768,576,4,64;768,576,3,4;1920,1080,4,1;1920,1080,4,8;1920,1080,4,32;3840,2160,5,1;3840,2160,5,4;3840,2160,5,16"
for m in level hybrid auto; do
  echo mode=$m
  PYR_SWEEP_ONLY="$S" PYR_SWEEP_OUT=gpurun_out/r2_j19_$m.json PSNW_PYR_MODE=$m timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200
done
