#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== graph check"; timeout 300 python scripts/check_graph.py 2>&1 | tail -2
  echo "== pytest gpu all"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
  echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-block 2> gpurun_out/bench_o.err | tee gpurun_out/bench_o.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], round(d['loop_ms_per_lp'],3), {k:round(v['ms']*1e3,1) for k,v in d['phases'].items()}, d['concurrent_lps']['value'])"
  echo "== bnb slots 32"; timeout 300 python bench.py --workload bnb --slots 32 --steps 20 --warmup 3 2>> gpurun_out/bnb.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['nodes'], round(d['lp_device_ms_per_node'],2), d['lp_iterations'])"
  tail -2 gpurun_out/bench_o.err
} > gpurun_out/round40.log 2>&1
cat gpurun_out/round40.log
