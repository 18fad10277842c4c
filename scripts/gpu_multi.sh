#!/usr/bin/env bash
# usage: gpu_multi.sh N  - LP replicas and B&B (windows, continuous) on N GPUs of one box
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
{
  echo "== LP replicas x$N"; timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-block 2> gpurun_out/multi_$N.err | grep '^{' | tee gpurun_out/lp_${N}gpu.json | cut -c1-200
  echo "== bnb windows x$N"; timeout 600 $TR bench.py --gpus $N --workload bnb --slots 32 --steps 20 --warmup 3 2>> gpurun_out/multi_$N.err | grep '^{' | tee gpurun_out/bnb_${N}gpu.json | cut -c1-200
  echo "== bnb stream x$N"; timeout 600 $TR bench.py --gpus $N --workload bnb --slots 32 --steps 5 --warmup 3 --stream-factor 4 2>> gpurun_out/multi_$N.err | grep '^{' | tee gpurun_out/bnb_stream_${N}gpu.json | cut -c1-200
  tail -3 gpurun_out/multi_$N.err
} > gpurun_out/multi_$N.log 2>&1
cat gpurun_out/multi_$N.log
