#!/bin/bash
# One B200: validates the element-organised update kernel (NEKB_GS_FUSE_UPDATE=4), the reworked one-launch AMG CG and the
# captured V-cycle; times them.
set -x
mkdir -p gpurun_out
T=r2f
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -8
timeout 300 python scripts/exp_gs_fuse.py --m 64 --its 100 > gpurun_out/${T}_gs_fuse.json 2> gpurun_out/${T}_gs_fuse.err
tail -3 gpurun_out/${T}_gs_fuse.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_gs_fuse.json'))
print(d['bit_identical_small'])
for k,v in d['runs'].items():
    for r in v: print(k, round(r['gdofs'],2), {a:round(b,3) for a,b in r['kernel_ms'].items()})
PY
for m in 48 64; do
NEKB_CRS_AMG=1 timeout 300 python scripts/bench_hsmg.py --m $m --calls 10 > gpurun_out/${T}_hsmg_m${m}_amg.json 2> gpurun_out/${T}_hsmg_m${m}_amg.err
tail -3 gpurun_out/${T}_hsmg_m${m}_amg.err; cat gpurun_out/${T}_hsmg_m${m}_amg.json
done
NEKB_H1MG_GRAPH=0 NEKB_CRS_AMG=1 timeout 300 python scripts/bench_hsmg.py --m 48 --calls 10 > gpurun_out/${T}_hsmg_m48_amg_nograph.json 2> gpurun_out/${T}_hsmg_m48_nograph.err
cat gpurun_out/${T}_hsmg_m48_amg_nograph.json
timeout 300 python tests/_mgpu_channel_worker.py 2>&1 | grep -E "CHANNEL" | tee gpurun_out/${T}_channel_n1_graph.log
NEKB_H1MG_GRAPH=0 timeout 300 python tests/_mgpu_channel_worker.py 2>&1 | grep -E "CHANNEL" | tee gpurun_out/${T}_channel_n1_nograph.log
du -sh gpurun_out
