#!/usr/bin/env bash
# last check of the round: the GPU suite, smoke() and one bench line with the library as committed
mkdir -p gpurun_out
{
  echo "== pytest gpu all"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  echo "== bench (default flags)"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python -c "import json; d=json.load(open('gpurun_out/bench_final.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], d['roofline']['frac'], d['cpu_baseline']['value'], d['pcg_50kx1M']['lp_solve']['time_to_lp_opt_s'], d['clocks'])"
} > gpurun_out/round44.log 2>&1
cat gpurun_out/round44.log
