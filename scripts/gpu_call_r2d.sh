#!/bin/bash
# Fourth GPU call of round 2 (one B200): the structured-gather cggos update (default) and the one-launch aggregation-CG coarse
# solve -- tests first, then timings, then the launch list of one h1mg_solve at 48^3 elements (HSMG kernel shares).
set -x
mkdir -p gpurun_out
T=r2d
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -15
timeout 200 python scripts/exp_gs_fuse.py --m 64 --its 100 > gpurun_out/${T}_gs_fuse.json 2> gpurun_out/${T}_gs_fuse.err
tail -3 gpurun_out/${T}_gs_fuse.err; cat gpurun_out/${T}_gs_fuse.json
timeout 240 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -3 gpurun_out/${T}_bench_n1.err; cut -c1-400 gpurun_out/${T}_bench_n1.json
for m in 48 64; do
NEKB_CRS_AMG=1 timeout 300 python scripts/bench_hsmg.py --m $m --calls 10 > gpurun_out/${T}_hsmg_m${m}_amg.json 2> gpurun_out/${T}_hsmg_m${m}_amg.err
tail -3 gpurun_out/${T}_hsmg_m${m}_amg.err; cat gpurun_out/${T}_hsmg_m${m}_amg.json
done
NEKB_CRS_AMG=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_hsmg_m48_launches.csv \
   python scripts/bench_hsmg.py --m 48 --calls 2 --no-gmres > gpurun_out/${T}_hsmg_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2d_hsmg_m48_launches.csv')) if len(r)>10]
hdr=rows[0]; ki,vi=hdr.index("Kernel Name"),hdr.index("Metric Value")
data=rows[1:]
# the last h1mg_solve call: from the last mg_mask_faces_kernel of the top level backwards ... simply aggregate the last 1/6 of launches
agg=collections.OrderedDict(); cnt=collections.Counter()
last=[i for i,r in enumerate(data) if 'mg_mask_faces' in r[ki]]
start=last[-2] if len(last)>=2 else 0
for r in data[start:]:
    k=r[ki].split('(')[0].replace('nekb::','').replace('void ','')
    agg[k]=agg.get(k,0)+float(r[vi].replace(',','')); cnt[k]+=1
tot=sum(agg.values())
print("h1mg_solve launch list (last call), total us:", tot/1e3)
for k,v in agg.items(): print(f"{k:60s} {cnt[k]:4d} {v/1e3:10.1f} us {100*v/tot:5.1f}%")
PY
# keep only the tail of the launch csv (size)
tail -n 700 gpurun_out/${T}_hsmg_m48_launches.csv > gpurun_out/${T}_hsmg_m48_launches_tail.csv; rm gpurun_out/${T}_hsmg_m48_launches.csv
du -sh gpurun_out
