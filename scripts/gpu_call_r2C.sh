#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2C
timeout 700 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -6
python __graft_entry__.py smoke 2>&1 | tail -2
