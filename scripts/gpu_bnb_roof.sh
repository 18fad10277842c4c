#!/usr/bin/env bash
python bench.py --no-pcg-block --no-cpu-baseline --quick-single --no-e2e --no-phases --steps 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'frac',round(d['roofline']['frac'],4))
for k,v in d['bnb'].items(): print(k, round(v['value'],1), v.get('roofline'), v['batching'][:60])"
