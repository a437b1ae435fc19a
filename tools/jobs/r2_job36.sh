#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_box_detect.py tests/test_gpu_configs.py tests/test_gpu_pipeline.py tests/test_gpu_e2e.py -x -q 2>&1 | tail -2
PSNW_DETECT_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_box_detect.py tests/test_gpu_configs.py tests/test_gpu_pipeline.py tests/test_gpu_e2e.py tests/test_gpu_chain.py -x -q 2>&1 | tail -2
for sp in 0 1; do echo "split=$sp"; PSNW_DETECT_SPLIT=$sp python tools/step_parts.py c3; PSNW_DETECT_SPLIT=$sp python tools/step_parts.py c2; done
