#!/usr/bin/env bash
tag=${1:-warm2}
mkdir -p gpurun_out
{
  echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_cta.py tests/test_gpu_bnb.py -x -q 2>&1 | tail -4
  for inst in scpnre1 scpnrg1; do for extra in "" "--warm-start"; do
    echo "== bnb $inst $extra"; timeout 600 python bench.py --workload bnb --bnb-instance $inst --steps 12 --warmup 3 $extra 2>>gpurun_out/${tag}.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); b=d['bnb']
print({a:(round(b[a],2) if isinstance(b[a],float) else b[a]) for a in ('value','nodes','lp_iterations_per_node','lp_device_ms_per_node','ms_per_round','incumbent')}); print('  ', b['rank0']['round_ms'], b['rank0'])"
  done; done
  tail -3 gpurun_out/${tag}.err
} > gpurun_out/${tag}.log 2>&1
cat gpurun_out/${tag}.log
