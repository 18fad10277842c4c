#!/usr/bin/env bash
# One GPU pass: the GPU suite, smoke(), the reference arm (short) and one default bench line.
#   bash scripts/gpu_round.sh <tag> [bench args...]
tag=${1:-round}; shift || true
mkdir -p gpurun_out
{
  echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -16
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/${tag}_ref.err | tee gpurun_out/${tag}_ref.json | cut -c1-600
  echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 "$@" 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms_per_step',round(d['ms_per_step'],3), d['config']['workload'], d['config'].get('form'))
print('window_ms',d.get('window_ms'),'e2e',d['e2e'])
print('roofline',{k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k in ('kernel','achieved','peak','frac','traffic','ms_per_launch','factorisation_phase_per_sm','us_per_iteration_inside_one_block')})
sl=d.get('single_lp') or d
print('single_lp value',round(sl['value'],1),'e2e',round(sl['e2e']['value'],1),'loop_ms',round(sl['loop_ms_per_lp'],3))
print({k:round(v['ms']*1e3,1) for k,v in sl['phases'].items()})
print('single roofline',{k:(round(v,4) if isinstance(v,float) else v) for k,v in sl['roofline'].items() if k in ('kernel','achieved','peak','frac')})
for k,v in (d.get('bnb') or {}).items(): print('bnb',k,{a:v.get(a) for a in ('value','nodes','lp_iterations_per_node','lp_device_ms_per_node','incumbent','root_bound','model','error')}, v.get('rank0'))
print('pcg',{k:d.get('pcg_50kx1M',{}).get(k) for k in ('cg_iteration_us','lp_solve','error')})
print('cpu',d.get('cpu_baseline'))
"
  tail -5 gpurun_out/${tag}_bench.err
} > gpurun_out/${tag}.log 2>&1
cat gpurun_out/${tag}.log
