#!/usr/bin/env bash
tag=${1:-cta}
mkdir -p gpurun_out
{
  echo "== pytest cta"; timeout 900 python -m pytest tests/test_gpu_cta.py -x -q --durations=6 2>&1 | tail -30
  echo "== pytest bnb"; timeout 900 python -m pytest tests/test_gpu_bnb.py -x -q 2>&1 | tail -6
  for inst in scpnre1 scpnrg1; do for sl in 128; do
    echo "== bnb $inst slots $sl (one CTA per LP)"; timeout 600 python bench.py --workload bnb --bnb-instance $inst --slots $sl --steps 6 --warmup 2 2>>gpurun_out/${tag}.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); b=d['bnb']
print({a:(round(b[a],2) if isinstance(b[a],float) else b[a]) for a in ('value','nodes','lp_iterations_per_node','lp_device_ms_per_node','ms_per_round','incumbent')}); print(b['rank0'])"
  done; done
  tail -5 gpurun_out/${tag}.err
} > gpurun_out/${tag}.log 2>&1
cat gpurun_out/${tag}.log
