#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lk.py tests/test_gpu_configs.py tests/test_gpu_e2e.py tests/test_gpu_pipeline.py tests/test_ref_pin.py -m gpu -x -q 2>&1 | tail -4
python tools/quick_time.py
QT_WORKLOAD=c3 python tools/quick_time.py
QT_WORKLOAD=c1 python tools/quick_time.py
