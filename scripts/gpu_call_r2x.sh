#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2x
timeout 700 python -m pytest tests -q -m gpu -x 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -8
