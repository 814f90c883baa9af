#!/bin/bash
# One 8 x B200 box (charged 8x), lean: parity at 8 ranks with the round-2 kernels, weak scaling at 8, strong scaling at 4 / 8.
T=r2B
set -x
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $RUN --nproc-per-node=8 --master-port 29703 tests/_mgpu_worker.py 2>&1 | grep -E "MGPU-OK|rror|assert" | head -12 | sed "s/^/[bp5 np=8 p2p=1] /" | tee -a gpurun_out/${T}_mgpu_workers.log
timeout 300 $RUN --nproc-per-node=8 --master-port 29768 bench.py --gpus 8 --steps 2 --warmup 3 --scaling weak --no-general > gpurun_out/${T}_bench_n8_weak.json 2> gpurun_out/${T}_bench_n8_weak.err
tail -2 gpurun_out/${T}_bench_n8_weak.err; cut -c1-160 gpurun_out/${T}_bench_n8_weak.json
for W in 8 4; do
  timeout 300 $RUN --nproc-per-node=$W --master-port $((29770+W)) bench.py --gpus $W --steps 2 --warmup 3 --scaling strong --no-general --no-check --no-e2e > gpurun_out/${T}_bench_n${W}_strong.json 2> gpurun_out/${T}_bench_n${W}_strong.err
  tail -2 gpurun_out/${T}_bench_n${W}_strong.err; cut -c1-160 gpurun_out/${T}_bench_n${W}_strong.json
done
timeout 300 $RUN --nproc-per-node=4 --master-port 29764 bench.py --gpus 4 --steps 2 --warmup 3 --scaling weak --no-general --no-check --no-e2e > gpurun_out/${T}_bench_n4_weak.json 2> gpurun_out/${T}_bench_n4_weak.err
tail -2 gpurun_out/${T}_bench_n4_weak.err; cut -c1-160 gpurun_out/${T}_bench_n4_weak.json
