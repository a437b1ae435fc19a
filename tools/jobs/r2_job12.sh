#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_box_detect.py tests/test_gpu_configs.py tests/test_gpu_e2e.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -4
python tools/step_parts.py c3
python tools/step_parts.py c2
