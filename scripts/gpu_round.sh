#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== pytest gpu bnb"; timeout 900 python -m pytest tests/test_gpu_bnb.py -x -q -m gpu 2>&1 | tail -4
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  for s in 16 32; do echo "== bnb slots $s"; timeout 300 python bench.py --workload bnb --slots $s --steps 20 --warmup 3 2>> gpurun_out/bnb.err | tee gpurun_out/bnb_n_s$s.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['nodes'], round(d['lp_device_ms_per_node'],2), d['lp_iterations'], d['rank0_rounds']['ms'][:5])"; done
  echo "== bnb slots 32 stream"; timeout 300 python bench.py --workload bnb --slots 32 --steps 5 --warmup 3 --stream-factor 4 2>> gpurun_out/bnb.err | tee gpurun_out/bnb_n_stream_s32.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['nodes'], round(d['lp_device_ms_per_node'],2), d['lp_iterations'])"
  echo "== heuristics kernel time (ncu, 3 launches)"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_node_heuristics" -s 20 -c 3 --csv python bench.py --workload bnb --slots 8 --steps 6 --warmup 1 2>/dev/null | grep k_node_heur | cut -d, -f 5,12- | head -3
} > gpurun_out/round42.log 2>&1
cat gpurun_out/round42.log
