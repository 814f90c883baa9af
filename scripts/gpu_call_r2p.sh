#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2p
scripts/ubench_dmma.bin > gpurun_out/${T}_ubench_dmma.txt 2>&1
cat gpurun_out/${T}_ubench_dmma.txt
timeout 400 python bench.py --steps 2 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -3 gpurun_out/${T}_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','answer_check','clocks')})
PY
