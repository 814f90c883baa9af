#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2o
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -6
timeout 300 python scripts/exp_gs_fuse.py --m 64 --its 100 --modes 4,5 > gpurun_out/${T}_gs_fuse.json 2> gpurun_out/${T}_gs_fuse.err
tail -3 gpurun_out/${T}_gs_fuse.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_gs_fuse.json'))
print(d['bit_identical_small'])
for k,v in d['runs'].items():
    for r in v: print('mode',k, round(r['gdofs'],2), {a:round(b,3) for a,b in r['kernel_ms'].items()})
"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/${T}_launches.csv python bench.py --steps 1 --warmup 3 --maxit 50 --no-cpu --no-e2e --no-general > gpurun_out/${T}_bench_under_ncu.log 2>&1
tail -n 300 /tmp/${T}_launches.csv > gpurun_out/${T}_launches_tail.csv
