#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pyramid.py -x -q 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_c.json').read().strip().splitlines()[-1])
r=d['roofline']; print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'], 'abi', d['e2e_host_abi']['ms_per_step'])
print('pyr stream ms',r['ms'],'frac',r['frac'],'isolated',r['isolated']['ms'],r['isolated']['frac'])
for m in r.get('frames_in_flight_sweep',[]): print(m['frame'],m['frames_per_launch'],'iso %.4f %.3f  stream %.4f %.3f'%(m['isolated_ms'],m['frac_isolated'],m['stream_ms'],m['frac_stream']), m['kernels_per_launch_group'], m['ring_slots'])
print('lk', d['lk']['ms'])
for k,v in d.get('configs',{}).items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if not isinstance(b,(dict,list,str))})
P
