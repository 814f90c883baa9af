#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2E
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -5
NEKB_CRS_AMG=1 timeout 200 python scripts/bench_hsmg.py --m 48 --calls 10 --no-gmres > gpurun_out/${T}_hsmg_m48_amg.json 2> gpurun_out/${T}_hsmg_m48_amg.err
tail -2 gpurun_out/${T}_hsmg_m48_amg.err; cat gpurun_out/${T}_hsmg_m48_amg.json
NEKB_H1MG_GRAPH=0 NEKB_CRS_AMG=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/${T}_launches.csv python scripts/bench_hsmg.py --m 48 --calls 1 --no-gmres > gpurun_out/${T}_hsmg_under_ncu.log 2>&1
grep "mg_fdm_kernel" /tmp/${T}_launches.csv | tail -2 | cut -c1-60,200-
tail -n 40 /tmp/${T}_launches.csv > gpurun_out/${T}_hsmg_launches_tail.csv
