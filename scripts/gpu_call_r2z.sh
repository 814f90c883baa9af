#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2z
show() { python - <<PY
import json
try:
    d=json.load(open('$1'))
    for s in d.get('small',[]): print('small',s)
    for r in d['runs']: print('$2', round(r['gdofs'],2), {a:round(b,3) for a,b in r['kernel_ms'].items()}, r['relerr'])
except Exception as e: print('$2 failed', e)
PY
}
for V in 20 0; do
  SK="--skip-small"; [ $V -eq 20 ] && SK=""
  NEKB_AXCG_VARIANT=$V timeout 300 python scripts/exp_axcg.py --m 64 --its 100 $SK > gpurun_out/${T}_affine_v$V.json 2> gpurun_out/${T}_affine_v$V.err
  tail -2 gpurun_out/${T}_affine_v$V.err
  show gpurun_out/${T}_affine_v$V.json affine_v$V
  NEKB_AXCG_VARIANT=$V timeout 300 python scripts/exp_axcg.py --m 64 --its 100 --general --skip-small > gpurun_out/${T}_general_v$V.json 2> gpurun_out/${T}_general_v$V.err
  tail -2 gpurun_out/${T}_general_v$V.err
  show gpurun_out/${T}_general_v$V.json general_v$V
done
