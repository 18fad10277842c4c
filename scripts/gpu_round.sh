#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== timeline"; timeout 60 scripts/bin/df_timeline 1024 > gpurun_out/tl11.log; grep "^rep\|residual" gpurun_out/tl11.log; grep "^trsv" gpurun_out/tl11.log | cut -c1-40
  echo "== timeline 4096"; timeout 60 scripts/bin/df_timeline 4096 | grep "^rep [23]\|residual"
  echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pcg-block 2> gpurun_out/bench_k.err | tee gpurun_out/bench_k.json
} > gpurun_out/round16.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/round16.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], {k:round(v['ms']*1e3,1) for k,v in d['phases'].items()})
    else: print(l.rstrip())
PY
