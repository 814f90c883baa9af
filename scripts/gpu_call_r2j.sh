#!/bin/bash
set -x
mkdir -p gpurun_out
T=r2k
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -8
for V in 0 1 2 3 4; do
  NEKB_UPD4_VARIANT=3 NEKB_AXCG_VARIANT=$V timeout 200 python scripts/exp_gs_fuse.py --skip-small --m 64 --its 100 --modes 4 > gpurun_out/${T}_affine_v$V.json 2> gpurun_out/${T}_affine_v$V.err
  tail -2 gpurun_out/${T}_affine_v$V.err
  python -c "
import json
d=json.load(open('gpurun_out/${T}_affine_v$V.json'))
for k,v in d['runs'].items():
    for r in v: print('affine variant $V', round(r['gdofs'],2), {a:round(b,3) for a,b in r['kernel_ms'].items()}, r['relerr'])
"
done
NEKB_UPD4_VARIANT=3 timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -3 gpurun_out/${T}_bench_n1.err; python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_n1.json'))
print(d['value'], d['e2e']['value'], d['operator_kernel'], d['general_geometry'], d['roofline']['frac'], d['roofline']['kernel_ms_per_iteration'], d['relerr'])
"
