#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2n
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -6
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -3 gpurun_out/${T}_bench_n1.err; cut -c1-200 gpurun_out/${T}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
cut -c1-300 gpurun_out/${T}_bench_reference.json
timeout 300 python scripts/bench_ophinv.py > gpurun_out/${T}_ophinv.json 2> gpurun_out/${T}_ophinv.err
tail -2 gpurun_out/${T}_ophinv.err; cat gpurun_out/${T}_ophinv.json
for V in 5 6; do
  NEKB_AXCG_VARIANT=$V timeout 200 python scripts/exp_gs_fuse.py --skip-small --m 64 --its 100 --modes 4 > gpurun_out/${T}_affine_v$V.json 2> gpurun_out/${T}_affine_v$V.err
  python -c "
import json
d=json.load(open('gpurun_out/${T}_affine_v$V.json'))
for k,v in d['runs'].items():
    for r in v: print('affine variant $V', round(r['gdofs'],2), {a:round(b,3) for a,b in r['kernel_ms'].items()})
"
done
for spec in ax_cg_affine_kernel:6 cggos_update4_kernel:6; do
  k=${spec%%:*}; skip=${spec#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip $skip -c 1 \
      -f -o /tmp/${T}_$k python scripts/bench_sweep.py --dims 64 --its 10 > gpurun_out/${T}_ncu_$k.log 2>&1
  ncu -i /tmp/${T}_$k.ncu-rep --page raw --csv > gpurun_out/${T}_$k.raw.csv 2>/dev/null
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/${T}_launches.csv python bench.py --steps 1 --warmup 3 --maxit 50 --no-cpu --no-e2e > gpurun_out/${T}_bench_under_ncu.log 2>&1
tail -n 500 /tmp/${T}_launches.csv > gpurun_out/${T}_launches_tail.csv
python __graft_entry__.py smoke 2>&1 | tail -2
du -sh gpurun_out
