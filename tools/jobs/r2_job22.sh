#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pyramid.py -x -q 2>&1 | tail -2
for v in 1 2; do PSNW_FUSED_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_pyramid.py -x -q -k "fused or hybrid" 2>&1 | tail -1; done
S="768,576,4,1;768,576,4,4;768,576,4,16;768,576,4,64;768,576,3,4;1920,1080,4,1"
PYR_SWEEP_ONLY="$S" timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200
