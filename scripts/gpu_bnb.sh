#!/usr/bin/env bash
# B&B pass: the B&B GPU tests and the bnb workload of the bench on the two configs[4] instances.
#   bash scripts/gpu_bnb.sh <tag> [gpus]
tag=${1:-bnb}; n=${2:-1}
mkdir -p gpurun_out
run() { if [ "$n" -gt 1 ]; then timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; else timeout 600 python bench.py "$@"; fi; }
{
  if [ "$n" -eq 1 ]; then echo "== pytest gpu bnb"; timeout 900 python -m pytest tests/test_gpu_bnb.py -x -q --durations=5 2>&1 | tail -12; fi
  for inst in scpnre1 scpnrg1; do
    for nlp in reference converged; do
      echo "== bnb $inst $nlp x$n"; run --workload bnb --bnb-instance $inst --node-lp $nlp --steps 8 --warmup 3 2>> gpurun_out/${tag}.err | tee -a gpurun_out/${tag}.jsonl | python -c "
import sys,json
d=json.loads(sys.stdin.read()); b=d['bnb']
print({a:b.get(a) for a in ('value','nodes','lp_iterations_per_node','lp_device_ms_per_node','incumbent','root_bound','model','ms_per_round')})
print(b['rank0']); print(b.get('exchange'))"
    done
  done
  tail -5 gpurun_out/${tag}.err
} > gpurun_out/${tag}.log 2>&1
cat gpurun_out/${tag}.log
