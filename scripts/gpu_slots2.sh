#!/usr/bin/env bash
# B&B windows larger than the SM count: ONE launch of K > 148 thread blocks, the hardware block scheduler starts the next
# node's block on whichever SM frees up first (no straggler wait inside the window)
mkdir -p gpurun_out
{
for inst in scpnre1 scpnrg1; do for sl in 148 296 444 592; do
  echo "== bnb $inst slots $sl"; timeout 600 python bench.py --workload bnb --bnb-instance $inst --slots $sl --steps 6 --warmup 3 2>>gpurun_out/slots2.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); b=d['bnb']
print({a:(round(b[a],2) if isinstance(b[a],float) else b[a]) for a in ('value','nodes','lp_iterations_per_node','lp_device_ms_per_node','ms_per_round','incumbent')}); print(b['rank0'])"
done; done
tail -5 gpurun_out/slots2.err
} > gpurun_out/slots2.log 2>&1
cat gpurun_out/slots2.log
