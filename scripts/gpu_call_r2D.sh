#!/bin/bash
# ncu --set full of the multigrid smoother and transfer kernels at 32^3 elements.  ncu's -k matches the base
# name; the instance is chosen by position in the V-cycle: mg_fdm_kernel launch 0 = <10,2> (top level), 1 = <6,6>;
# mg_tensor3_t_kernel launch 0 = <4,8> (restriction 8 -> 4), 3 = <8,4> (prolongation 4 -> 8).
set -x
mkdir -p gpurun_out
T=r2D
for spec in "mg_fdm_kernel:fdm10:0" "mg_fdm_kernel:fdm6:1" "mg_tensor3_t_kernel:t3_4_8:0" "mg_tensor3_t_kernel:t3_8_4:3"; do
  k=${spec%%:*}; rest=${spec#*:}; name=${rest%%:*}; skip=${rest#*:}
  NEKB_H1MG_GRAPH=0 NEKB_CRS_AMG=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip $skip -c 1 \
      -f -o /tmp/${T}_$name python scripts/bench_hsmg.py --m 32 --calls 1 --no-gmres > gpurun_out/${T}_ncu_$name.log 2>&1
  ncu -i /tmp/${T}_$name.ncu-rep --page raw --csv > gpurun_out/${T}_$name.raw.csv 2>/dev/null
  tail -1 gpurun_out/${T}_ncu_$name.log | cut -c1-200
done
ls -la gpurun_out/${T}_*csv
