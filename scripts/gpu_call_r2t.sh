#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2t
timeout 700 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -6
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -3 gpurun_out/${T}_bench_n1.err; cut -c1-300 gpurun_out/${T}_bench_n1.json
for spec in ax_cg_affine_mma_kernel:6 cggos_update6_kernel:6 gs_gval_kernel:6; do
  k=${spec%%:*}; skip=${spec#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip $skip -c 1 \
      -f -o /tmp/${T}_$k python scripts/bench_sweep.py --dims 64 --its 10 > gpurun_out/${T}_ncu_$k.log 2>&1
  ncu -i /tmp/${T}_$k.ncu-rep --page raw --csv > gpurun_out/${T}_$k.raw.csv 2>/dev/null
done
NEKB_AX_AFFINE=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ax_cg_mma_kernel --launch-skip 6 -c 1 \
      -f -o /tmp/${T}_axgen python scripts/bench_sweep.py --dims 64 --its 10 > gpurun_out/${T}_ncu_ax_cg_mma_kernel.log 2>&1
ncu -i /tmp/${T}_axgen.ncu-rep --page raw --csv > gpurun_out/${T}_ax_cg_mma_kernel.raw.csv 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/${T}_launches.csv python bench.py --steps 1 --warmup 3 --maxit 50 --no-cpu --no-e2e --no-general --no-check > gpurun_out/${T}_bench_under_ncu.log 2>&1
tail -n 300 /tmp/${T}_launches.csv > gpurun_out/${T}_launches_tail.csv
python __graft_entry__.py smoke 2>&1 | tail -2
du -sh gpurun_out
