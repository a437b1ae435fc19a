#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo rc=$?
tail -5 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().splitlines()[-1])
print(json.dumps({k:d[k] for k in d if k in ('value','ms_per_step','e2e','e2e_host_abi','e2e_gather','per_rank_ms_per_step','n_gpus')}, indent=0)[:3000])
PY
