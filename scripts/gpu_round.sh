#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== unfused tool sanity"; timeout 120 scripts/bin/df_timeline_la1 1000 | grep -E "^rep|^info"
  echo "== pytest gpu all"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
  echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-block 2> gpurun_out/bench_o.err | tee gpurun_out/bench_o.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k:round(v['ms']*1e3,1) for k,v in d['phases'].items()}, d['concurrent_lps']['value'])"
  tail -3 gpurun_out/bench_o.err
} > gpurun_out/round38.log 2>&1
cat gpurun_out/round38.log
