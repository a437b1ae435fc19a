#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PSNW_PYR_MODE=fused PSNW_FUSED_VARIANT=0
PYR_SWEEP_ONLY=768,576,4,4 python tools/pyr_sweep.py > gpurun_out/r2_p4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pyr_fused -s 6 -c 1 -o gpurun_out/r2_ncu_f3_4 env PYR_SWEEP_ONLY=768,576,4,4 python tools/pyr_sweep.py > gpurun_out/r2_ncu_f3_4.log 2>&1
