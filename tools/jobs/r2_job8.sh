#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
PSNW_FUSED_THREADS=256 timeout 600 python -m pytest tests/test_gpu_pyramid.py -x -q -k "fused or hybrid" 2>&1 | tail -1
S="768,576,4,4;768,576,4,16;768,576,4,64;1920,1080,4,8;1920,1080,4,32;3840,2160,5,4;3840,2160,5,16"
for m in hybrid fused; do for v in 0 1 2; do
  echo $m variant=$v threads=256
  PYR_SWEEP_ONLY="$S" PYR_SWEEP_OUT=gpurun_out/r2_j8_${m}_v$v.json PSNW_PYR_MODE=$m PSNW_FUSED_VARIANT=$v PSNW_FUSED_THREADS=256 timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200
done; done
