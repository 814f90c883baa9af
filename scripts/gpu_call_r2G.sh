#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2G
timeout 120 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -5
