#!/usr/bin/env bash
mkdir -p gpurun_out
{
  for f in 1 2 4 8; do echo "== bnb slots 32 share factor $f"; SB200_SHARE_FACTOR=$f timeout 300 python bench.py --workload bnb --slots 32 --steps 20 --warmup 3 2>> gpurun_out/bnb.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['nodes'], round(d['lp_device_ms_per_node'],2), d['lp_iterations'])"; done
  echo "== bnb slots 16 share factor 4"; SB200_SHARE_FACTOR=4 timeout 300 python bench.py --workload bnb --slots 16 --steps 20 --warmup 3 2>> gpurun_out/bnb.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['nodes'], round(d['lp_device_ms_per_node'],2), d['lp_iterations'])"
} > gpurun_out/round41.log 2>&1
cat gpurun_out/round41.log
