#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
PYR_SWEEP_ONLY=768,576,4,64 python tools/pyr_sweep.py > gpurun_out/r2_p64.log 2>&1 &&
PYR_SWEEP_ONLY=768,576,4,64 ncu --set full --clock-control none --import-source on -k regex:pyr_fused -s 6 -c 1 -o gpurun_out/r2_ncu_fused3_64 python tools/pyr_sweep.py > gpurun_out/r2_ncu_fused3_64.log 2>&1
tail -2 gpurun_out/r2_ncu_fused3_64.log
