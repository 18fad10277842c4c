#!/usr/bin/env bash
# last check of the round: the GPU suite, smoke() and one bench line with the library as committed
mkdir -p gpurun_out
{
  echo "== pytest gpu all"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-block 2> gpurun_out/bench_o.err | tee gpurun_out/bench_o.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['loop_ms_per_lp'],3), {k:round(v['ms']*1e3,1) for k,v in d['phases'].items()}, d['concurrent_lps']['value'])"
} > gpurun_out/round48.log 2>&1
cat gpurun_out/round48.log
