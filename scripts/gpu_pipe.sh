#!/usr/bin/env bash
# windows in flight (BatchedBnb pipeline = 2) against one window at a time
mkdir -p gpurun_out
{
  echo "== pytest"; timeout 600 python -m pytest tests/test_gpu_bnb.py tests/test_gpu_cta.py -x -q 2>&1 | tail -8
  for i in scpnre1 scpnrg1; do for pl in 2 1; do
    echo "== $i pipeline $pl"
    python bench.py --workload bnb --bnb-instance $i --steps 12 --warmup 3 --bnb-pipeline $pl 2>>gpurun_out/pipe.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); b=d['bnb']
print(round(b['value'],1), b['incumbent'], b['nodes'], round(b['lp_device_ms_per_node'],2), b['batching'])
print('  round_ms', b['rank0']['round_ms']); print('  wait_ms', b['rank0']['solve_ms']); print('  window_device_ms', b['rank0']['window_device_ms'])"
  done; done
  tail -5 gpurun_out/pipe.err
} > gpurun_out/pipe.log 2>&1
cat gpurun_out/pipe.log
