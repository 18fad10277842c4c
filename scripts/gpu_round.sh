#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== pytest bnb"; timeout 900 python -m pytest tests/test_gpu_bnb.py -x -q 2>&1 | tail -4
  for s in 8 16 32; do echo "== bnb slots $s"; timeout 600 python bench.py --workload bnb --steps 12 --slots $s 2>> gpurun_out/bench_bnb.err; done
} > gpurun_out/round18.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/round18.log'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','nodes','lp_iterations','lp_device_ms_per_node','incumbent','root_bound')}, d['config']['slots_per_gpu'])
    else: print(l.rstrip())
PY
