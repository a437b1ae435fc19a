#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
PSNW_HYB_CHUNKS=2 timeout 900 python -m pytest tests/test_gpu_pyramid.py -x -q -k "hybrid or walker_many" 2>&1 | tail -1
S="768,576,4,16;768,576,4,64;1920,1080,4,8;1920,1080,4,32;3840,2160,5,16"
for c in 1 2 4 8; do echo "hybrid chunks=$c"; PSNW_HYB_CHUNKS=$c PSNW_PYR_MODE=hybrid PYR_SWEEP_ONLY="$S" timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200; done
