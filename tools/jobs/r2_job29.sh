#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pyramid.py -x -q 2>&1 | tail -2
S="768,576,4,4;768,576,4,16;768,576,4,64;1920,1080,4,8;1920,1080,4,32;3840,2160,5,4;3840,2160,5,16"
for m in auto hybrid; do echo "mode=$m"; PSNW_PYR_MODE=$m PYR_SWEEP_ONLY="$S" PYR_SWEEP_OUT=gpurun_out/r2_sweep3_$m.json timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200; done
PSNW_PYR_MODE=hybrid PYR_SWEEP_ONLY=768,576,4,64 ncu --set full --clock-control none --import-source on -k regex:pyr_fused -s 6 -c 1 -o gpurun_out/r2_ncu_fused3_64c python tools/pyr_sweep.py > gpurun_out/r2_ncu_f3c.log 2>&1
