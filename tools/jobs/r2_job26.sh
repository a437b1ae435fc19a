#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pyramid.py -x -q -k "walk" 2>&1 | tail -1
S="768,576,4,64;1920,1080,4,32;3840,2160,5,16"
for c in 1 4 16; do echo "walk chunk=$c"; PSNW_WALK_CHUNK=$c PSNW_PYR_MODE=walk PYR_SWEEP_ONLY="$S" PYR_SWEEP_OUT=gpurun_out/r2_sweep_walkc$c.json timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200; done
PSNW_WALK_CHUNK=4 PSNW_PYR_MODE=walk PYR_SWEEP_ONLY=3840,2160,5,16 ncu --set full --clock-control none --import-source on -k regex:pyr_walk -s 6 -c 1 -o gpurun_out/r2_ncu_walk_4k16 python tools/pyr_sweep.py > gpurun_out/r2_ncu_walk.log 2>&1
tail -2 gpurun_out/r2_ncu_walk.log
