#!/bin/bash
# First GPU call of the next round, one B200, about 8 minutes of box time:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/next_round_first_call.sh'
# 1. the whole GPU suite INCLUDING what was written after round 1's GPU budget was spent (crs_* facade);
# 2. BASELINE config 3: the E sweep extended to 128^3 and a non-cubic fill;
# 3. the headline bench line;
# 4. ncu --set full of the three kernels of the open gather-scatter question: the stock pair
#    gs_local_kernel + cggos_update2_kernel and the gather-fused cggos_update2_gs_kernel (slower: where does the traffic go?).
set -x
mkdir -p gpurun_out
NEKB_TEST_UNVALIDATED=1 timeout 300 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/r2a_pytest_gpu.log | tail -8
timeout 150 python scripts/bench_sweep.py --dims 64,128,160x128x128 --its 60 > gpurun_out/r2a_sweep.json 2> gpurun_out/r2a_sweep.err
timeout 240 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
for k in gs_local_kernel cggos_update2_kernel cggos_update2_gs_kernel; do
  NEKB_GS_FUSE_UPDATE=1 timeout 240 ncu --set full --clock-control none --import-source on -k regex:"^$k" --launch-skip 12 -c 1 \
      -f -o gpurun_out/r2a_$k python scripts/exp_gs_fuse.py --skip-small --m 64 --its 4 > gpurun_out/r2a_ncu_$k.log 2>&1
done
ls -la gpurun_out
