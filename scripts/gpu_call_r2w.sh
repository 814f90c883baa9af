#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2w
timeout 700 python -m pytest tests -q -m gpu -x 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -8
timeout 300 python scripts/bench_ophinv.py > gpurun_out/${T}_ophinv.json 2> gpurun_out/${T}_ophinv.err
tail -2 gpurun_out/${T}_ophinv.err; cat gpurun_out/${T}_ophinv.json | cut -c1-900
NEKB_HCG_RHO_KERNEL=1 timeout 300 python scripts/bench_ophinv.py > gpurun_out/${T}_ophinv_rho_kernel.json 2> gpurun_out/${T}_ophinv_rho_kernel.err
cat gpurun_out/${T}_ophinv_rho_kernel.json | cut -c1-900
