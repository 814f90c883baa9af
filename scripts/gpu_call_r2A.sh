#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2A
timeout 700 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -6
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -3 gpurun_out/${T}_bench_n1.err; cut -c1-200 gpurun_out/${T}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
cut -c1-200 gpurun_out/${T}_bench_reference.json
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 700 python scripts/bench_sweep.py --dims 8,16,32,48,64,96,128,160x128x128 --its 50 > gpurun_out/${T}_sweep.json 2> gpurun_out/${T}_sweep.err
tail -2 gpurun_out/${T}_sweep.err; cat gpurun_out/${T}_sweep.json | cut -c1-1500
