#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo rc=$?
tail -5 gpurun_out/r2_bench_a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_a.json'))
print(json.dumps({k:d[k] for k in d if k not in ('configs','roofline','lk','cpu_baseline','config')}, indent=0)[:1500])
print(json.dumps(d['roofline'], indent=0)[:3000])
print(json.dumps(d['lk'], indent=0)[:1500])
print(json.dumps(d.get('configs'), indent=0)[:3000])
PY
