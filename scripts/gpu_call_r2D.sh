#!/bin/bash
# ncu --set full of the multigrid smoother and transfer kernels at 32^3 elements
set -x
mkdir -p gpurun_out
T=r2D
for spec in "mg_fdm_kernel<10:fdm10:0" "mg_tensor3_t_kernel<4, 8:t3_4_8:0" "mg_tensor3_t_kernel<8, 4:t3_8_4:0"; do
  k=${spec%%:*}; rest=${spec#*:}; name=${rest%%:*}; skip=${rest#*:}
  NEKB_H1MG_GRAPH=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip $skip -c 1 \
      -f -o /tmp/${T}_$name python scripts/bench_hsmg.py --m 32 --calls 2 --no-gmres > gpurun_out/${T}_ncu_$name.log 2>&1
  ncu -i /tmp/${T}_$name.ncu-rep --page raw --csv > gpurun_out/${T}_$name.raw.csv 2>/dev/null
  tail -2 gpurun_out/${T}_ncu_$name.log
done
ls -la gpurun_out/${T}_*
