#!/usr/bin/env bash
for i in scpnre1 scpnrg1; do for pl in 3 2; do
  python bench.py --workload bnb --bnb-instance $i --steps 16 --warmup 4 --bnb-pipeline $pl 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); b=d['bnb']
print('$i pipeline $pl', round(b['value'],1), b['incumbent'], b['nodes'], 'round_ms', b['rank0']['round_ms'])"
done; done
