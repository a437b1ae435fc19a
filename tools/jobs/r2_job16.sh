#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/bench_reference_pattern.py > gpurun_out/r2_refpat.json 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lk_generic -s 40 -c 4 -o gpurun_out/r2_ncu_generic python tools/bench_reference_pattern.py > gpurun_out/r2_ncu_generic.log 2>&1
tail -2 gpurun_out/r2_ncu_generic.log
