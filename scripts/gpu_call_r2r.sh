#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2r
for V in 0 1 2 3; do
  SK=""; [ $V -ge 1 ] && SK="--skip-small"
  NEKB_UPD6_VARIANT=$V timeout 300 python scripts/exp_gs_fuse.py --m 64 --its 100 --modes 4,6 $SK > gpurun_out/${T}_upd6_v$V.json 2> gpurun_out/${T}_upd6_v$V.err
  tail -2 gpurun_out/${T}_upd6_v$V.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${T}_upd6_v$V.json'))
    print(d['bit_identical_small'])
    for k,v in d['runs'].items():
        for r in v: print('upd6 variant $V mode',k, round(r['gdofs'],2), {a:round(b,3) for a,b in r['kernel_ms'].items()}, r['relerr'])
except Exception as e: print('variant $V failed', e)
PY
done
timeout 600 python -m pytest tests -q -m gpu -x -k "affine or cggos or bp5 or fused" 2>&1 | tail -5
