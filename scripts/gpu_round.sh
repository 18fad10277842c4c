#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pcg-block 2> gpurun_out/bench_j.err | tee gpurun_out/bench_j.json
} > gpurun_out/round15.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/round15.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], {k:round(v['ms']*1e3,1) for k,v in d['phases'].items()})
    else: print(l.rstrip())
PY
