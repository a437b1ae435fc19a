#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lk.py tests/test_gpu_e2e.py tests/test_gpu_chain.py tests/test_host_tracker2d.py -x -q 2>&1 | tail -2
timeout 600 python tools/bench_reference_pattern.py > gpurun_out/r2_refpat2.json 2>gpurun_out/r2_refpat2.err; tail -1 gpurun_out/r2_refpat2.json
PSNW_KLT_LIB=$PWD/build/variants/a_base.so timeout 600 python tools/bench_reference_pattern.py 2>/dev/null | tail -1
