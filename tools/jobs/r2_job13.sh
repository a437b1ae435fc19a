#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/step_parts.py c3 > gpurun_out/r2_parts_c3c.json 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"detect_select" -s 3 -c 1 -o gpurun_out/r2_ncu_det_c3 python tools/step_parts.py c3 > gpurun_out/r2_ncu_det_c3.log 2>&1
tail -2 gpurun_out/r2_ncu_det_c3.log
