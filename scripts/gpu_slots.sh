#!/usr/bin/env bash
# B&B throughput against the number of LP slots and the CTAs each factorisation may take
mkdir -p gpurun_out
out=gpurun_out/${1:-slots}.log; : > $out
for inst in scpnre1 scpnrg1; do
  for cfg in "32 2 4" "64 2 1" "64 1 1" "128 2 1" "128 1 1" "256 1 1"; do
    set -- $cfg
    echo "== $inst slots=$1 share_factor=$2 min_ctas=$3" >> $out
    SB200_SHARE_FACTOR=$2 SB200_MIN_CTAS=$3 timeout 300 python bench.py --workload bnb --bnb-instance $inst --slots $1 --steps 6 --warmup 2 2>>gpurun_out/slots.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); b=d['bnb']
print({a:(round(b[a],2) if isinstance(b[a],float) else b[a]) for a in ('value','nodes','lp_iterations_per_node','lp_device_ms_per_node','ms_per_round','incumbent')})" >> $out 2>&1
  done
done
cat $out; tail -3 gpurun_out/slots.err
