#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2i
for V in 0 1 2 3 4; do
  NEKB_UPD4_VARIANT=$V timeout 200 python scripts/exp_gs_fuse.py --skip-small --m 64 --its 100 --modes 4 > gpurun_out/${T}_upd4_v$V.json 2> gpurun_out/${T}_upd4_v$V.err
  tail -2 gpurun_out/${T}_upd4_v$V.err
  python -c "
import json
d=json.load(open('gpurun_out/${T}_upd4_v$V.json'))
for k,v in d['runs'].items():
    for r in v: print('variant $V mode',k, round(r['gdofs'],2), {a:round(b,3) for a,b in r['kernel_ms'].items()})
"
done
timeout 200 python scripts/exp_gs_fuse.py --skip-small --m 64 --its 100 --modes 0 > gpurun_out/${T}_stock.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/${T}_stock.json'))
for k,v in d['runs'].items():
    for r in v: print('stock mode',k, round(r['gdofs'],2), {a:round(b,3) for a,b in r['kernel_ms'].items()})
"
