#!/bin/bash
# Third GPU call of round 2 (one B200): BASELINE config 3 (E sweep to the HBM fill, lean set-up above 10^6 elements) and
# the coarse solve beyond the dense limit (h1mg_solve at 48^3 and 64^3 elements: Jacobi-PCG vs CG over the aggregation hierarchy).
set -x
mkdir -p gpurun_out
T=r2c
timeout 900 python scripts/bench_sweep.py --dims 8,16,32,48,64,96,128,144x128x128,160x128x128 --its 60 > gpurun_out/${T}_sweep.json 2> gpurun_out/${T}_sweep.err
tail -12 gpurun_out/${T}_sweep.err
NEKB_CRS_AMG=1 timeout 300 python scripts/bench_hsmg.py --m 48 --calls 10 > gpurun_out/${T}_hsmg_m48_amg.json 2> gpurun_out/${T}_hsmg_m48_amg.err
tail -3 gpurun_out/${T}_hsmg_m48_amg.err; cat gpurun_out/${T}_hsmg_m48_amg.json
NEKB_CRS_AMG=1 timeout 400 python scripts/bench_hsmg.py --m 64 --calls 10 > gpurun_out/${T}_hsmg_m64_amg.json 2> gpurun_out/${T}_hsmg_m64_amg.err
tail -3 gpurun_out/${T}_hsmg_m64_amg.err; cat gpurun_out/${T}_hsmg_m64_amg.json
timeout 200 python -m pytest tests/test_gpu_hsmg.py tests/test_zz_gpu_configs.py -q -m gpu -x 2>&1 | tail -4
du -sh gpurun_out
