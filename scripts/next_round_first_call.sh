#!/bin/bash
# First GPU call of the next round, one B200, about 12 minutes of box time:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/next_round_first_call.sh'
# 1. the whole GPU suite INCLUDING what was written after round 1's GPU budget was spent (crs_* facade);
# 2. BASELINE config 3: the E sweep extended to 128^3 and a non-cubic fill;
# 3. the headline bench line;
# 4. ncu --set full of the three kernels of the open gather-scatter question: the stock pair
#    gs_local_kernel + cggos_update2_kernel and the gather-fused cggos_update2_gs_kernel (slower: where does the traffic go?);
# 5. h1mg_solve at 48^3 elements with the Jacobi-PCG coarse solve and with NEKB_CRS_AMG=1.
set -x
mkdir -p gpurun_out
NEKB_TEST_UNVALIDATED=1 timeout 300 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/r2a_pytest_gpu.log | tail -8
timeout 150 python scripts/bench_sweep.py --dims 64,128,160x128x128 --its 60 > gpurun_out/r2a_sweep.json 2> gpurun_out/r2a_sweep.err
timeout 240 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
# exp_gs_fuse.py --skip-small --its 4 launches, per mode 0 / 1 / 2 / 0 / 1 / 2: 13 x (gs_local + update2) | 13 x update2_gs |
# 13 x (gs_local on the edge/corner groups + update2_gs) | ...
for spec in gs_local_kernel:12:stock cggos_update2_kernel:12:stock cggos_update2_gs_kernel:12:mode1 cggos_update2_gs_kernel:25:mode2 gs_local_kernel:25:mode2; do
  k=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; tag=${rest#*:}
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:"^$k" --launch-skip $skip -c 1 \
      -f -o gpurun_out/r2a_${k}_$tag python scripts/exp_gs_fuse.py --skip-small --m 64 --its 4 > gpurun_out/r2a_ncu_${k}_$tag.log 2>&1
done
timeout 120 python scripts/exp_gs_fuse.py --m 64 --its 100 > gpurun_out/r2a_gs_fuse.json 2> gpurun_out/r2a_gs_fuse.err
# 5. coarse solve beyond the dense limit at 48^3 elements: Jacobi-PCG (today's path) against CG over the aggregation hierarchy
timeout 200 python scripts/bench_hsmg.py --m 48 --calls 10 > gpurun_out/r2a_hsmg_m48_pcg.json 2> gpurun_out/r2a_hsmg_m48_pcg.err
NEKB_CRS_AMG=1 timeout 200 python scripts/bench_hsmg.py --m 48 --calls 10 > gpurun_out/r2a_hsmg_m48_amg.json 2> gpurun_out/r2a_hsmg_m48_amg.err
ls -la gpurun_out
