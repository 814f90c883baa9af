#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2s
show() { python - <<PY
import json
try:
    d=json.load(open('$1'))
    for s in d.get('small',[]): print('small',s)
    for r in d['runs']: print('$2', round(r['gdofs'],2), {a:round(b,3) for a,b in r['kernel_ms'].items()}, r['relerr'])
except Exception as e: print('$2 failed', e)
PY
}
# affine default (mma + erec + update6) vs erec off
timeout 300 python scripts/exp_axcg.py --m 64 --its 100 > gpurun_out/${T}_affine_default.json 2> gpurun_out/${T}_affine_default.err; tail -2 gpurun_out/${T}_affine_default.err
show gpurun_out/${T}_affine_default.json affine_default
NEKB_GS_EREC=0 timeout 300 python scripts/exp_axcg.py --m 64 --its 100 --skip-small > gpurun_out/${T}_affine_noerec.json 2> gpurun_out/${T}_affine_noerec.err
show gpurun_out/${T}_affine_noerec.json affine_noerec
# general geometry: old kernel (0) vs mma variants
for V in 0 10 11 12; do
  SK="--skip-small"; [ $V -eq 10 ] && SK=""
  NEKB_AXCG_VARIANT=$V timeout 300 python scripts/exp_axcg.py --m 64 --its 100 --general $SK > gpurun_out/${T}_general_v$V.json 2> gpurun_out/${T}_general_v$V.err
  tail -2 gpurun_out/${T}_general_v$V.err
  show gpurun_out/${T}_general_v$V.json general_v$V
done
