#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pyramid.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_b.json').read().strip().splitlines()[-1])
r=d['roofline']; print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'])
print('pyr stream ms',r['ms'],'frac',r['frac'],'isolated',r['isolated']['ms'],r['isolated']['frac'])
print(json.dumps(d.get('frames_in_flight_sweep', d.get('pyramid_sweep')))[:1500])
P
PSNW_PYR_OVERLAP=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('no-overlap: pyr stream ms',r['ms'],'frac',r['frac'])"
