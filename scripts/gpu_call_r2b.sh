#!/bin/bash
# Second GPU call of round 2 (one B200).  Everything brought back is small: ncu reports are exported to CSV on the box
# (scripts/ncu_rawcsv.sh) and the .ncu-rep files removed, because gpurun_out/ is capped at 64 MiB.
set -x
mkdir -p gpurun_out
T=r2b
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -5
timeout 200 python scripts/bench_sweep.py --dims 64,128,160x128x128 --its 60 > gpurun_out/${T}_sweep.json 2> gpurun_out/${T}_sweep.err
tail -3 gpurun_out/${T}_sweep.err
timeout 240 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
for spec in gs_local_kernel:12:stock cggos_update2_kernel:12:stock cggos_update2_gs_kernel:25:mode2 gs_local_kernel:25:mode2; do
  k=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; tag=${rest#*:}
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:"^$k" --launch-skip $skip -c 1 \
      -f -o /tmp/${T}_${k}_$tag python scripts/exp_gs_fuse.py --skip-small --m 64 --its 4 > gpurun_out/${T}_ncu_${k}_$tag.log 2>&1
  ncu -i /tmp/${T}_${k}_$tag.ncu-rep --page raw --csv > gpurun_out/${T}_${k}_$tag.raw.csv 2>/dev/null
done
timeout 120 python scripts/exp_gs_fuse.py --m 64 --its 100 > gpurun_out/${T}_gs_fuse.json 2> gpurun_out/${T}_gs_fuse.err
timeout 200 python scripts/bench_hsmg.py --m 48 --calls 10 > gpurun_out/${T}_hsmg_m48_pcg.json 2> gpurun_out/${T}_hsmg_m48_pcg.err
NEKB_CRS_AMG=1 timeout 200 python scripts/bench_hsmg.py --m 48 --calls 10 > gpurun_out/${T}_hsmg_m48_amg.json 2> gpurun_out/${T}_hsmg_m48_amg.err
tail -5 gpurun_out/${T}_hsmg_m48_amg.err
du -sh gpurun_out
