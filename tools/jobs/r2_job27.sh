#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pyramid.py -x -q -k "walk" 2>&1 | tail -1
PSNW_WALK_BASE=1 timeout 600 python -m pytest tests/test_gpu_pyramid.py -x -q -k "walk" 2>&1 | tail -1
S="768,576,4,16;768,576,4,64;1920,1080,4,8;1920,1080,4,32;3840,2160,5,4;3840,2160,5,16"
for b in 0 1; do echo "walk base=$b"; PSNW_WALK_BASE=$b PSNW_PYR_MODE=walk PYR_SWEEP_ONLY="$S" PYR_SWEEP_OUT=gpurun_out/r2_sweep_walkb$b.json timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200; done
