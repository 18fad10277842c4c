#!/usr/bin/env bash
# final lines of the round with the final library
mkdir -p gpurun_out
{
  echo "== pytest gpu all"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  echo "== bench full"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; cut -c1-200 gpurun_out/bench_l.json
  rm -f gpurun_out/bench_l_configs.jsonl
  for w in scp4 scpnrf scpclr13; do echo "== bench $w"; timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-pcg-block 2>> gpurun_out/bench_l.err | tee -a gpurun_out/bench_l_configs.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), {k:round(v['ms']*1e3,1) for k,v in d['phases'].items()})"; done
} > gpurun_out/round36.log 2>&1
cat gpurun_out/round36.log
