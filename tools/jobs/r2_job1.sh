#!/bin/bash
# round-2 baseline: pyramid sweep in both builder modes, ncu of the fused kernel at 4 and 64 frames, detect + LK at c3
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
PSNW_PYR_MODE=fused python tools/pyr_sweep.py > gpurun_out/r2_sweep_fused.log 2>&1; cp gpurun_out/pyr_sweep.json gpurun_out/r2_sweep_fused.json
PSNW_PYR_MODE=level python tools/pyr_sweep.py > gpurun_out/r2_sweep_level.log 2>&1; cp gpurun_out/pyr_sweep.json gpurun_out/r2_sweep_level.json
export PSNW_PYR_MODE=fused
PYR_SWEEP_ONLY=768,576,4,64 python tools/pyr_sweep.py > gpurun_out/r2_p64.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pyr_fused -s 3 -c 1 -o gpurun_out/r2_ncu_fused64 env PYR_SWEEP_ONLY=768,576,4,64 python tools/pyr_sweep.py > gpurun_out/r2_ncu_fused64.log 2>&1
PYR_SWEEP_ONLY=768,576,4,4 python tools/pyr_sweep.py > gpurun_out/r2_p4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"pyr_fused|apron" -s 6 -c 2 -o gpurun_out/r2_ncu_fused4 env PYR_SWEEP_ONLY=768,576,4,4 python tools/pyr_sweep.py > gpurun_out/r2_ncu_fused4.log 2>&1
unset PSNW_PYR_MODE
python tools/step_parts.py c3 > gpurun_out/r2_parts_c3b.json 2>gpurun_out/r2_parts_c3b.err &&
ncu --set full --clock-control none --import-source on -k regex:"detect_select|lk_w21|min_eig|box_update" -s 12 -c 5 -o gpurun_out/r2_ncu_c3 python tools/step_parts.py c3 > gpurun_out/r2_ncu_c3.log 2>&1
ls -la gpurun_out
