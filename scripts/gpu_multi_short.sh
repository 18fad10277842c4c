#!/usr/bin/env bash
# N-GPU check of the driver's own invocation: default line at N (LP replicas + throughput block on rank 0 + B&B block on all ranks)
tag=${1:-multi}; n=${2:-8}
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n --steps 10 --warmup 3 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'n_gpus',d['n_gpus'], 'single_lp', (d.get('single_lp') or {}).get('value'))
for k,v in (d.get('bnb') or {}).items(): print('bnb',k,{a:v.get(a) for a in ('value','nodes','lp_iterations_per_node','incumbent','ms_per_round','error')}); print('   ', (v.get('exchange') or {}).get('per_rank_[nodes, busy_ms, exchange_wait_ms, open_nodes]'))
" > gpurun_out/${tag}.log 2>&1
tail -4 gpurun_out/${tag}_bench.err >> gpurun_out/${tag}.log
cat gpurun_out/${tag}.log
