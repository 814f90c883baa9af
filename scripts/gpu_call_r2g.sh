#!/bin/bash
# One B200: ncu --set full of the kernels of the BP5 iteration (structured gather, mode 4) and of the h1mg V-cycle kernels at
# 32^3 elements; exported to CSV on the box (reports are ~15 MB each).
set -x
mkdir -p gpurun_out
T=r2g
export NEKB_GS_FUSE_UPDATE=4
for spec in ax_cg_kernel:6 cggos_update4_kernel:6 gs_gval_kernel:6; do
  k=${spec%%:*}; skip=${spec#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip $skip -c 1 \
      -f -o /tmp/${T}_$k python scripts/bench_sweep.py --dims 64 --its 10 > gpurun_out/${T}_ncu_$k.log 2>&1
  ncu -i /tmp/${T}_$k.ncu-rep --page raw --csv > gpurun_out/${T}_$k.raw.csv 2>/dev/null
done
# launch list of one bench step (shares)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/${T}_launches.csv python bench.py --steps 1 --warmup 3 --maxit 50 --no-cpu --no-e2e > gpurun_out/${T}_bench_under_ncu.log 2>&1
tail -n 400 /tmp/${T}_launches.csv > gpurun_out/${T}_launches_tail.csv
unset NEKB_GS_FUSE_UPDATE
export NEKB_H1MG_GRAPH=0
for spec in "mg_fdm_kernel<10:1" "mg_fdm_kernel<6:1" mg_tensor3_kernel:4 mg_add_overlap_kernel:2 mg_mask_faces_kernel:2; do
  k=${spec%%:*}; skip=${spec#*:}; kn=$(echo $k | tr -d '<')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip $skip -c 1 \
      -f -o /tmp/${T}_$kn python scripts/bench_hsmg.py --m 32 --calls 2 --no-gmres > gpurun_out/${T}_ncu_$kn.log 2>&1
  ncu -i /tmp/${T}_$kn.ncu-rep --page raw --csv > gpurun_out/${T}_$kn.raw.csv 2>/dev/null
done
du -sh gpurun_out
