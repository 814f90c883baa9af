#!/bin/bash
# Multi-GPU call (gpurun --gpus N): the torchrun parity tests with their logs kept, the headline bench at N (weak and strong,
# with the parity field), and BASELINE config 5 (channel pressure solve, .ma2 partition) at 1..N ranks.
# usage: bash scripts/gpu_call_mgpu.sh <N> <tag> [full]     ("full" also runs the whole single-GPU suite first)
N=$1; T=$2; FULL=$3
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$FULL" = "full" ]; then
  timeout 900 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/${T}_pytest_gpu.log | tail -12
else
  timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_gpu" 2>&1 | tee gpurun_out/${T}_pytest_mgpu.log | tail -12
fi
# the workers once more by hand so that their MGPU-OK lines are on record
for W in 2 4 8; do
  if [ $W -le $N ]; then
    for P2P in 1 0; do
      NEKB_GS_P2P=$P2P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port $((29700+W)) \
        tests/_mgpu_worker.py 2>&1 | grep -E "MGPU-OK|Error|error|assert" | head -12 | sed "s/^/[bp5 np=$W p2p=$P2P] /" | tee -a gpurun_out/${T}_mgpu_workers.log
    done
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port $((29720+W)) \
      tests/_mgpu_hsmg_worker.py 2>&1 | grep -E "MGPU-HSMG-OK|Error|error|assert" | head -12 | sed "s/^/[hsmg np=$W] /" | tee -a gpurun_out/${T}_mgpu_workers.log
  fi
done
for W in 1 2 4 8; do
  if [ $W -le $N ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port $((29740+W)) \
      tests/_mgpu_channel_worker.py 2>&1 | grep -E "MGPU-CHANNEL-OK|CHANNEL-JSON|Error|error|assert" | head -12 | sed "s/^/[channel np=$W] /" | tee -a gpurun_out/${T}_channel.log
  fi
done
for W in 1 2 4 8; do
  if [ $W -le $N ]; then
    for SC in weak strong; do
      if [ $W -eq 1 ] && [ $SC = strong ]; then continue; fi
      if [ $W -eq 1 ]; then
        timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_bench_n1_$SC.json 2> gpurun_out/${T}_bench_n1_$SC.err
      else
        timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port $((29760+W)) \
          bench.py --gpus $W --steps 3 --warmup 3 --scaling $SC > gpurun_out/${T}_bench_n${W}_$SC.json 2> gpurun_out/${T}_bench_n${W}_$SC.err
      fi
      tail -2 gpurun_out/${T}_bench_n${W}_$SC.err; cut -c1-300 gpurun_out/${T}_bench_n${W}_$SC.json
    done
  fi
done
du -sh gpurun_out
