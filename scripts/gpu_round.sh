#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== timeline"; timeout 60 scripts/bin/df_timeline 1024 > gpurun_out/tl13.log; grep "^rep\|residual" gpurun_out/tl13.log; grep -A2 "^chain   [67] " gpurun_out/tl13.log
  echo "== timeline 4096"; timeout 60 scripts/bin/df_timeline 4096 | grep "^rep [23]\|residual"
  echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-block 2> gpurun_out/bench_n.err | tee gpurun_out/bench_n.json
} > gpurun_out/round20.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/round20.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['gpu_launches'], {k:round(v['ms']*1e3,1) for k,v in d['phases'].items()})
    else: print(l.rstrip())
PY
