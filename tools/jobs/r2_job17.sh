#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_lk.py tests/test_gpu_box_detect.py tests/test_gpu_chain.py tests/test_host_tracker2d.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1])
print({k:(d[k]['ms_per_step'] if isinstance(d[k],dict) else d[k]) for k in ('ms_per_step','e2e','e2e_host_abi')})"
