#!/usr/bin/env bash
# the one-block solver after a kernel change: its tests, the per-phase times, a short headline line
tag=${1:-cta2}; shift || true
mkdir -p gpurun_out
{
  echo "== pytest cta + bnb"; timeout 900 python -m pytest tests/test_gpu_cta.py tests/test_gpu_bnb.py -x -q --durations=4 -p timeout --timeout 120 2>&1 | tail -12
  echo "== phases"; timeout 600 python scripts/cta_phases.py 148 2>&1 | tail -8
  echo "== headline"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pcg-block --quick-single --no-phases "$@" 2>gpurun_out/${tag}.err | tee gpurun_out/${tag}_bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'ms_per_step',round(d['ms_per_step'],2),'frac',round(d['roofline']['frac'],4))
print(d['roofline'].get('us_per_iteration_inside_one_block'))
for k,v in (d.get('bnb') or {}).items(): print('bnb',k,{a:v.get(a) for a in ('value','nodes','lp_iterations_per_node','incumbent','error')})
"
  tail -3 gpurun_out/${tag}.err
} > gpurun_out/${tag}.log 2>&1
cat gpurun_out/${tag}.log
