#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2u
timeout 700 python -m pytest tests -q -m gpu -x tests/test_gpu_hsmg.py tests/test_gpu_golden.py tests/test_zz_gpu_configs.py tests/test_gpu_dropin.py 2>&1 | tee gpurun_out/${T}_pytest_gpu_hsmg.log | tail -6
NEKB_CRS_AMG=1 timeout 300 python scripts/bench_hsmg.py --m 48 --calls 10 > gpurun_out/${T}_hsmg_m48_amg.json 2> gpurun_out/${T}_hsmg_m48_amg.err
tail -3 gpurun_out/${T}_hsmg_m48_amg.err; cat gpurun_out/${T}_hsmg_m48_amg.json
NEKB_MG_TENSOR3_GENERIC=1 NEKB_CRS_AMG=1 timeout 300 python scripts/bench_hsmg.py --m 48 --calls 10 > gpurun_out/${T}_hsmg_m48_amg_generic_t3.json 2> gpurun_out/${T}_hsmg_m48_amg_generic_t3.err
cat gpurun_out/${T}_hsmg_m48_amg_generic_t3.json
timeout 300 python tests/_mgpu_channel_worker.py 2>&1 | grep -E "CHANNEL" | tee gpurun_out/${T}_channel_n1_graph.log
NEKB_H1MG_GRAPH=0 NEKB_CRS_AMG=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/${T}_launches.csv python scripts/bench_hsmg.py --m 48 --calls 2 > gpurun_out/${T}_hsmg_under_ncu.log 2>&1
tail -n 120 /tmp/${T}_launches.csv > gpurun_out/${T}_hsmg_launches_tail.csv
