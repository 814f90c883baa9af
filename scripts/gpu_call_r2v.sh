#!/bin/bash
# 2 x B200: multi-GPU parity workers with the new default kernels (DMMA operator kernels, TMA-ring update), bench at N = 2.
T=r2v
set -x
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for P2P in 1 0; do
  NEKB_GS_P2P=$P2P timeout 300 $RUN --nproc-per-node=2 --master-port 29701 tests/_mgpu_worker.py 2>&1 | grep -E "MGPU-OK|rror|assert" | head -8 | sed "s/^/[bp5 np=2 p2p=$P2P] /" | tee -a gpurun_out/${T}_mgpu_workers.log
done
timeout 300 $RUN --nproc-per-node=2 --master-port 29704 tests/_mgpu_hsmg_worker.py 2>&1 | grep -E "MGPU-HSMG-OK|rror|assert" | head -8 | sed "s/^/[hsmg np=2] /" | tee -a gpurun_out/${T}_mgpu_workers.log
timeout 300 $RUN --nproc-per-node=2 --master-port 29705 tests/_mgpu_channel_worker.py 2>&1 | grep -E "MGPU-CHANNEL-OK|CHANNEL-JSON|rror|assert" | head -8 | sed "s/^/[channel np=2] /" | tee -a gpurun_out/${T}_mgpu_workers.log
timeout 400 $RUN --nproc-per-node=2 --master-port 29768 bench.py --gpus 2 --steps 2 --warmup 3 --scaling weak > gpurun_out/${T}_bench_n2_weak.json 2> gpurun_out/${T}_bench_n2_weak.err
tail -2 gpurun_out/${T}_bench_n2_weak.err; cut -c1-160 gpurun_out/${T}_bench_n2_weak.json
timeout 400 $RUN --nproc-per-node=2 --master-port 29772 bench.py --gpus 2 --steps 2 --warmup 3 --scaling strong --no-general --no-check > gpurun_out/${T}_bench_n2_strong.json 2> gpurun_out/${T}_bench_n2_strong.err
tail -2 gpurun_out/${T}_bench_n2_strong.err; cut -c1-160 gpurun_out/${T}_bench_n2_strong.json
