#!/bin/bash
# One 8 x B200 box (charged 8x): multi-GPU parity on record at 4 and 8 ranks, BASELINE config 5 at 4 / 8 ranks, weak scaling
# at 1 / 8 and STRONG scaling (E = 262,144 in total) at 2 / 4 / 8 with the parity field.
T=$1
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# 4-rank parity runs side by side on the two halves of the box (not timed)
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 $RUN --nproc-per-node=4 --master-port 29701 tests/_mgpu_worker.py 2>&1 | grep -E "MGPU-OK|rror|assert" | head -8 | sed "s/^/[bp5 np=4 p2p=1] /" > gpurun_out/${T}_w4a.log ) &
( CUDA_VISIBLE_DEVICES=4,5,6,7 NEKB_CHANNEL_CALLS=20 timeout 300 $RUN --nproc-per-node=4 --master-port 29702 tests/_mgpu_channel_worker.py 2>&1 | grep -E "MGPU-CHANNEL-OK|CHANNEL-JSON|rror|assert" | head -8 | sed "s/^/[channel np=4] /" > gpurun_out/${T}_w4b.log ) &
wait
cat gpurun_out/${T}_w4a.log gpurun_out/${T}_w4b.log | tee -a gpurun_out/${T}_mgpu_workers.log
for P2P in 1 0; do
  NEKB_GS_P2P=$P2P timeout 300 $RUN --nproc-per-node=8 --master-port 29703 tests/_mgpu_worker.py 2>&1 | grep -E "MGPU-OK|rror|assert" | head -12 | sed "s/^/[bp5 np=8 p2p=$P2P] /" | tee -a gpurun_out/${T}_mgpu_workers.log
done
timeout 300 $RUN --nproc-per-node=8 --master-port 29704 tests/_mgpu_hsmg_worker.py 2>&1 | grep -E "MGPU-HSMG-OK|rror|assert" | head -12 | sed "s/^/[hsmg np=8] /" | tee -a gpurun_out/${T}_mgpu_workers.log
NEKB_CHANNEL_CALLS=20 timeout 300 $RUN --nproc-per-node=8 --master-port 29705 tests/_mgpu_channel_worker.py 2>&1 | grep -E "MGPU-CHANNEL-OK|CHANNEL-JSON|rror|assert" | head -12 | sed "s/^/[channel np=8] /" | tee -a gpurun_out/${T}_mgpu_workers.log
rm -f gpurun_out/${T}_w4a.log gpurun_out/${T}_w4b.log
timeout 300 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_bench_n1_weak.json 2> gpurun_out/${T}_bench_n1_weak.err
cut -c1-160 gpurun_out/${T}_bench_n1_weak.json
timeout 400 $RUN --nproc-per-node=8 --master-port 29768 bench.py --gpus 8 --steps 2 --warmup 3 --scaling weak > gpurun_out/${T}_bench_n8_weak.json 2> gpurun_out/${T}_bench_n8_weak.err
tail -2 gpurun_out/${T}_bench_n8_weak.err; cut -c1-160 gpurun_out/${T}_bench_n8_weak.json
for W in 2 4 8; do
  timeout 400 $RUN --nproc-per-node=$W --master-port $((29770+W)) bench.py --gpus $W --steps 2 --warmup 3 --scaling strong > gpurun_out/${T}_bench_n${W}_strong.json 2> gpurun_out/${T}_bench_n${W}_strong.err
  tail -2 gpurun_out/${T}_bench_n${W}_strong.err; cut -c1-160 gpurun_out/${T}_bench_n${W}_strong.json
done
du -sh gpurun_out
