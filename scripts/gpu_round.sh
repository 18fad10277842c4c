#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-solve 2> gpurun_out/bench_o.err | tee gpurun_out/bench_o.json
} > gpurun_out/round21.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/round21.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['gpu_launches'], {k:round(v['ms']*1e3,1) for k,v in d['phases'].items()}, d['pcg_50kx1M'].get('cg_iteration_us'))
    else: print(l.rstrip())
PY
