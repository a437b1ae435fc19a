#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_box_detect.py tests/test_gpu_configs.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -3
python tools/step_parts.py c2; python tools/step_parts.py c3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'lk', d['lk']['ms'])"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:min_eig -c 3 python tools/step_parts.py c3 2>&1 | grep -E "min_eig|duration|inst_executed" | head -8
