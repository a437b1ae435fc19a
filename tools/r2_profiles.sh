#!/bin/bash
# round-2 evidence (final code): driver-shaped bench line, launch list of the bench step, ncu --set full of the LK kernel, of
# the pyramid kernels (4 / 64 frames of 768x576, 16 x 4K), of the c3 step's kernels, of the generic-window kernel; step parts
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_c2_reference_arm.json 2>> gpurun_out/r02_bench_c2.err
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
python tools/pyr_sweep.py > gpurun_out/r02_pyr_sweep.log 2>&1; cp gpurun_out/pyr_sweep.json gpurun_out/r02_pyr_sweep.json
python tools/step_parts.py c3 > gpurun_out/r02_parts_c3.json 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"detect_select|lk_w21|box_update|min_eig" -s 12 -c 5 -o gpurun_out/r02_ncu_c3 python tools/step_parts.py c3 > gpurun_out/r02_ncu_c3.log 2>&1
python tools/step_parts.py c2 > gpurun_out/r02_parts_c2.json 2>&1
python tools/bench_reference_pattern.py > gpurun_out/r02_refpattern_gpu.json 2>/dev/null &&
ncu --set full --clock-control none --import-source on -k regex:lk_generic -s 40 -c 2 -o gpurun_out/r02_ncu_lk_generic python tools/bench_reference_pattern.py > gpurun_out/r02_ncu_generic.log 2>&1
python tools/frames_in_flight.py 768x576 > gpurun_out/r02_frames_in_flight.log 2>&1; cp gpurun_out/frames_in_flight.json gpurun_out/r02_frames_in_flight.json
ls -la gpurun_out | grep r02 | wc -l
