#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_refpat_launches.csv python tools/bench_reference_pattern.py > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_refpat_launches.csv
ncu --set full --clock-control none --import-source on -k regex:lk_generic -s 40 -c 4 -o gpurun_out/r2_ncu_generic2 python tools/bench_reference_pattern.py > gpurun_out/r2_ncu_generic2.log 2>&1
