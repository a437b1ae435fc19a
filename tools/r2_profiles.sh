#!/bin/bash
# round-2 evidence: launch list of the bench step, ncu --set full of the LK kernel, of the pyramid kernels (4 and 64 frames), of the c3 step
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs"
$B > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_c2_steps20.csv $B > gpurun_out/r02_launches.log 2>&1
python tools/quick_time.py > gpurun_out/r02_qt.json 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lk_w21 -s 12 -c 1 -o gpurun_out/r02_ncu_lk_w21 python tools/quick_time.py > gpurun_out/r02_ncu_lk.log 2>&1
PYR_SWEEP_ONLY=768,576,4,4 python tools/pyr_sweep.py > gpurun_out/r02_p4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pyr_ -s 6 -c 1 -o gpurun_out/r02_ncu_pyr_4 env PYR_SWEEP_ONLY=768,576,4,4 python tools/pyr_sweep.py > gpurun_out/r02_ncu_p4.log 2>&1
PYR_SWEEP_ONLY=768,576,4,64 python tools/pyr_sweep.py > gpurun_out/r02_p64.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pyr_ -s 8 -c 2 -o gpurun_out/r02_ncu_pyr_64 env PYR_SWEEP_ONLY=768,576,4,64 python tools/pyr_sweep.py > gpurun_out/r02_ncu_p64.log 2>&1
PYR_SWEEP_ONLY=3840,2160,5,16 python tools/pyr_sweep.py > gpurun_out/r02_p4k.log 2>&1 &&
ncu --set full --clock-control none -k regex:pyr_ -s 15 -c 5 -o gpurun_out/r02_ncu_pyr_4k16 env PYR_SWEEP_ONLY=3840,2160,5,16 python tools/pyr_sweep.py > gpurun_out/r02_ncu_p4k.log 2>&1
python tools/step_parts.py c3 > gpurun_out/r02_parts_c3.json 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"detect_select|lk_w21|box_update|min_eig" -s 12 -c 5 -o gpurun_out/r02_ncu_c3 python tools/step_parts.py c3 > gpurun_out/r02_ncu_c3.log 2>&1
python tools/step_parts.py c2 > gpurun_out/r02_parts_c2.json 2>&1
ls -la gpurun_out | grep r02
