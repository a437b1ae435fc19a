#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for m in level hybrid; do
PSNW_PYR_MODE=$m PYR_SWEEP_ONLY=768,576,4,64 python tools/pyr_sweep.py > gpurun_out/r2_p64_$m.log 2>&1 &&
PSNW_PYR_MODE=$m PYR_SWEEP_ONLY=768,576,4,64 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pyr_ -s 12 -c 8 --csv --log-file gpurun_out/r2_l64_$m.csv python tools/pyr_sweep.py > /dev/null 2>&1
done
python - <<'PY'
import csv
for m in ('level','hybrid'):
    rows=[l for l in open('gpurun_out/r2_l64_%s.csv'%m) if not l.startswith('==')]
    for r in csv.DictReader(rows):
        print(m, r['ID'], r['Kernel Name'][:40], r['Grid Size'] if 'Grid Size' in r else '', r['Metric Name'], r['Metric Value'], r['Metric Unit'])
PY
