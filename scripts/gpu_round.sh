#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== pytest gpu all"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
  echo "== bench scpnrf"; timeout 300 python bench.py --workload scpnrf --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-block 2>> gpurun_out/bench_o.err | tee gpurun_out/bench_scpnrf.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v['ms']*1e3,1) for k,v in d['phases'].items()})"
  for s in 16 32; do echo "== bnb slots $s"; timeout 300 python bench.py --workload bnb --slots $s --steps 20 --warmup 3 2>> gpurun_out/bnb.err | tee gpurun_out/bnb_l_s$s.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['nodes'], round(d['lp_device_ms_per_node'],2), d['lp_iterations'])"; done
  echo "== bnb slots 32 stream"; timeout 300 python bench.py --workload bnb --slots 32 --steps 5 --warmup 3 --stream-factor 4 2>> gpurun_out/bnb.err | tee gpurun_out/bnb_l_stream_s32.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['nodes'], round(d['lp_device_ms_per_node'],2), d['lp_iterations'])"
} > gpurun_out/round32.log 2>&1
cat gpurun_out/round32.log
