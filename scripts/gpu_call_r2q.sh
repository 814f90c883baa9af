#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2q
for V in 0 10 11 12 13 14; do
  SK=""; [ $V -ge 11 ] && SK="--skip-small"
  NEKB_AXCG_VARIANT=$V timeout 300 python scripts/exp_axcg.py --m 64 --its 100 $SK > gpurun_out/${T}_axcg_v$V.json 2> gpurun_out/${T}_axcg_v$V.err
  tail -2 gpurun_out/${T}_axcg_v$V.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${T}_axcg_v$V.json'))
    for s in d.get('small',[]): print('small',s)
    for r in d['runs']: print('variant $V', round(r['gdofs'],2), {a:round(b,3) for a,b in r['kernel_ms'].items()}, r['relerr'])
except Exception as e: print('variant $V failed', e)
PY
done
