#!/usr/bin/env bash
# N-GPU pass: the default bench line (LP replicas + B&B block) and the bnb workload under torchrun
#   bash scripts/gpu_multi.sh <tag> <gpus>
tag=${1:-multi}; n=${2:-2}
mkdir -p gpurun_out
tr() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@"; }
{
  echo "== default line x$n"; tr --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'n_gpus',d['n_gpus'])
for k,v in (d.get('bnb') or {}).items(): print('bnb',k,{a:v.get(a) for a in ('value','nodes','lp_iterations_per_node','incumbent','ms_per_round','error')}); print('   ', v.get('exchange'))
"
  tail -3 gpurun_out/${tag}_bench.err
  echo "== reference arm x$n"; tr --impl reference --steps 2 --warmup 1 2>> gpurun_out/${tag}_bench.err | cut -c1-300
  for inst in scpnre1 scpnrg1; do
    echo "== bnb $inst x$n"; tr --workload bnb --bnb-instance $inst --steps 10 --warmup 3 2>> gpurun_out/${tag}_bnb.err | tee -a gpurun_out/${tag}_bnb.jsonl | python -c "
import sys,json
d=json.loads(sys.stdin.read()); b=d['bnb']
print({a:b.get(a) for a in ('value','nodes','lp_iterations_per_node','incumbent','ms_per_round')}); print('   ', b.get('exchange'))"
  done
  tail -3 gpurun_out/${tag}_bnb.err
} > gpurun_out/${tag}.log 2>&1
cat gpurun_out/${tag}.log
