#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2F
timeout 150 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -2 gpurun_out/${T}_bench_n1.err; cut -c1-200 gpurun_out/${T}_bench_n1.json
