#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2y
timeout 700 python -m pytest tests -q -m gpu -x 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -8
timeout 300 python scripts/bench_ophinv.py > gpurun_out/${T}_ophinv.json 2> gpurun_out/${T}_ophinv.err
tail -2 gpurun_out/${T}_ophinv.err; cat gpurun_out/${T}_ophinv.json | cut -c1-700
NEKB_HCG_STRUCT=0 timeout 300 python scripts/bench_ophinv.py > gpurun_out/${T}_ophinv_nostruct.json 2> gpurun_out/${T}_ophinv_nostruct.err
cat gpurun_out/${T}_ophinv_nostruct.json | cut -c1-700
