#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
PSNW_HYB_LEVELS=2 timeout 600 python -m pytest tests/test_gpu_pyramid.py -q -k "hybrid" 2>&1 | tail -1
S="768,576,4,16;768,576,4,32;768,576,4,64;1920,1080,4,8;1920,1080,4,32;3840,2160,5,4;3840,2160,5,16"
echo "level"; PSNW_PYR_MODE=level PYR_SWEEP_ONLY="$S" timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200
for v in 0 1 2; do echo "hybrid tail2 variant $v"; PSNW_FUSED_VARIANT=$v PSNW_HYB_LEVELS=2 PSNW_PYR_MODE=hybrid PYR_SWEEP_ONLY="$S" timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200; done
